#!/usr/bin/env python
"""Oracle fingerprints of the full-size single-structure configurations (BASELINE cfg4: 150k atoms x 100 points, cfg5: 1M atoms
x 960 points): sha256 of the per-atom exposed-point counts (uint32, little endian, atom order) and their sum, computed by the
CPU oracle on the seeded workloads of rustsasa_b200/workloads.py.  The -m gpu tests and bench.py's `secondary` block recompute
the counts on the GPU and compare against tests/golden/cfg_hashes.json (running the oracle at 1M x 960 takes 5-20 s of CPU,
too long to repeat on every rank of every bench run).   usage: python tools/make_cfg_hashes.py"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import load  # noqa: E402
from rustsasa_b200 import workloads as W  # noqa: E402


def fingerprint(counts):
    c = np.ascontiguousarray(counts, dtype="<u4")
    return hashlib.sha256(c.tobytes()).hexdigest(), int(c.astype(np.int64).sum())


def main():
    fast = load(fast=True)
    out = {}
    for key, data, n_points in (("cfg4", W.large_assembly(150000), 100), ("cfg5", W.capsid_shell(1000000), 960)):
        t0 = time.perf_counter()
        o = fast.calculate_sasa_internal(data.xyzr, 1.4, n_points, threads=-1)
        h, s = fingerprint(o["counts"])
        out[key] = dict(atoms=data.n_atoms, n_points=n_points, probe=1.4, lanes=8, sha256_counts=h, sum_counts=s,
                        xyzr_sha256=hashlib.sha256(np.ascontiguousarray(data.xyzr).tobytes()).hexdigest(),
                        oracle_seconds=round(time.perf_counter() - t0, 1))
        print(key, out[key], flush=True)
    # cfg3: protein totals (global, polar, non-polar) of the frames a rank of an N = 1, 2, 4 or 8 job starts with
    md = W.md_trajectory(n_frames=10000, n_atoms=5000)
    frames = sorted({10000 * r // w for w in (1, 2, 4, 8) for r in range(w)})
    tot = {}
    for f in frames:
        xyzr = np.concatenate([md.xyz[f], md.radii[:, None]], axis=1).astype(np.float32)
        o = fast.calculate_sasa_internal(xyzr, 1.4, 100, threads=-1)
        t = fast.protein_totals(o["sasa"], md.seg_be, md.seg_polar)
        tot[str(f)] = [float(x) for x in np.asarray(t, np.float32)]
    out["cfg3"] = dict(frames=10000, atoms_per_frame=int(md.xyz.shape[1]), n_points=100, protein_totals=tot,
                       xyz_sha256_first_frame=hashlib.sha256(np.ascontiguousarray(md.xyz[0]).tobytes()).hexdigest())
    print("cfg3", out["cfg3"], flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "cfg_hashes.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
