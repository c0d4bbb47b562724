#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference's own test data (run in the build container).

Inputs (read-only, NOT available on the GPU box, hence the committed fixtures):
  /root/reference/tests/common/data.rs          FIXED_LOW_RES_ATOMS -- golden per-atom SASA produced by
                                                the real crate (tests/units.rs:17-43 configuration)
  /root/reference/tests/data/pdbs/*             example.cif, 151L_H3.pdb, bad_seqadv_1A06.pdb
  /root/reference/tests/data/freesasa_pdbs/*    88 PDB files of the quality regression (tests/quality.rs)
  /root/reference/tests/data/freesasa_reference FreeSASA chain totals for the same files

Outputs:
  tests/golden/example_cif_vdw.npz   atoms of example.cif in pdb.atoms() order with pdbtbx vdW radii,
                                     the golden SASA vector and the integer counts recovered from it
  tests/golden/structures.npz        every structure above after row A0 (default options: ProtOr radii,
                                     no hydrogens, no HETATM), stored compactly: coordinates as integer
                                     milli-Angstrom (exact for 3-decimal files), radii as table indices,
                                     residue ranges, polar flags, chain labels, FreeSASA chain totals

Only extracted numeric data is stored -- no reference source.
"""
import glob
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rustsasa_b200.structure import (SASACalcError, VDW_RADII, build_atoms_and_mapping,  # noqa: E402
                                     read_structure)

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def golden_vector():
    src = open(f"{REF}/tests/common/data.rs").read()
    body = src.split("FIXED_LOW_RES_ATOMS: [f32; 2622] = [")[1].split("];")[0]
    return np.array([np.float32(t) for t in re.findall(r"[-+0-9.eE]+", body)], dtype=np.float32)


def make_example_vdw():
    st = read_structure(f"{REF}/tests/data/pdbs/example.cif")
    atoms = list(st.atoms())
    xyzr = np.array([[a.x, a.y, a.z, VDW_RADII[a.element]] for a in atoms], dtype=np.float32)
    gold = golden_vector()
    assert gold.shape[0] == xyzr.shape[0] == 2622
    r = xyzr[:, 3] + np.float32(1.4)
    unit = (np.float32(4.0) * np.float32(np.pi)) * (r * r) / np.float32(100)
    q = gold.astype(np.float64) / unit.astype(np.float64)
    counts = np.rint(q).astype(np.uint32)
    assert np.abs(q - counts).max() < 2e-3
    np.savez_compressed(os.path.join(OUT, "example_cif_vdw.npz"), xyzr=xyzr, gold_sasa=gold, gold_counts=counts)
    print("example_cif_vdw: atoms", xyzr.shape[0], "sum counts", counts.sum(), "sum sasa", gold.sum(dtype=np.float32))


def pack(path, name, store):
    st = read_structure(path)
    try:
        p = build_atoms_and_mapping(st, "residue")
    except SASACalcError as e:
        print(f"  skip {name}: {e.kind}: {e}")
        return False
    milli = np.rint(p.xyzr[:, :3].astype(np.float64) * 1000.0).astype(np.int32)
    back = (milli.astype(np.float64) / 1000.0).astype(np.float32)
    assert np.array_equal(back, p.xyzr[:, :3]), name
    chains = []
    res_chain = np.zeros(len(p.seg_meta), np.uint16)
    for k, m in enumerate(p.seg_meta):
        if m["chain_id"] not in chains:
            chains.append(m["chain_id"])
        res_chain[k] = chains.index(m["chain_id"])
    store[name] = dict(milli=milli, radius=p.xyzr[:, 3].copy(), seg_be=p.seg_be, polar=p.seg_polar,
                       res_chain=res_chain, chains=chains,
                       res_names=[m["name"] for m in p.seg_meta])
    return True


def main():
    os.makedirs(OUT, exist_ok=True)
    make_example_vdw()
    store = {}
    for f in ("example.cif", "151L_H3.pdb", "bad_seqadv_1A06.pdb"):
        pack(f"{REF}/tests/data/pdbs/{f}", f, store)
    fs_totals = {}
    for path in sorted(glob.glob(f"{REF}/tests/data/freesasa_pdbs/*.pdb")):
        stem = os.path.splitext(os.path.basename(path))[0]
        if pack(path, stem, store):
            ref = json.load(open(f"{REF}/tests/data/freesasa_reference/{stem}.json"))
            tot = {}
            for result in ref["results"]:
                for s in result["structure"]:
                    for c in s["chains"]:
                        tot[c["label"]] = c["area"]["total"]
            fs_totals[stem] = tot
    names = list(store)
    radii_table = np.unique(np.concatenate([store[n]["radius"] for n in names]))
    arrays = dict(names=np.array(names), radii_table=radii_table.astype(np.float32),
                  atom_off=np.cumsum([0] + [store[n]["milli"].shape[0] for n in names]).astype(np.int64),
                  seg_off=np.cumsum([0] + [store[n]["seg_be"].shape[0] for n in names]).astype(np.int64),
                  milli=np.concatenate([store[n]["milli"] for n in names]),
                  radius_idx=np.concatenate([np.searchsorted(radii_table, store[n]["radius"]) for n in names]
                                            ).astype(np.uint8),
                  seg_be=np.concatenate([store[n]["seg_be"] for n in names]).astype(np.uint32),
                  polar=np.concatenate([store[n]["polar"] for n in names]).astype(np.uint8),
                  res_chain=np.concatenate([store[n]["res_chain"] for n in names]).astype(np.uint16),
                  res_names=np.array(sum((store[n]["res_names"] for n in names), [])),
                  chains_json=np.array(json.dumps({n: store[n]["chains"] for n in names})),
                  freesasa_json=np.array(json.dumps(fs_totals)))
    path = os.path.join(OUT, "structures.npz")
    np.savez_compressed(path, **arrays)
    print("structures:", len(names), "atoms", arrays["milli"].shape[0], "segments", arrays["seg_be"].shape[0],
          "bytes", os.path.getsize(path))


if __name__ == "__main__":
    main()
