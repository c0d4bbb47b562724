#!/usr/bin/env python
"""End-to-end directory mode of the C++ CLI (row f-1): N synthetic PDB files -> sasa_b200_cli <in> <out> --format json.
Files are poly-alanine re-labellings of the committed coordinate templates (tests/golden/structures.npz): every five
consecutive atoms become one ALA residue (N, CA, C, O, CB), so that parsing, ProtOr radius lookup, packing, the engine
and the JSON writers all do real work.  The geometry is the template's; only the labels are synthetic.
usage (GPU box): cli_dir_bench.py [n_files] [level]"""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rustsasa_b200 import workloads as W  # noqa: E402
from rustsasa_b200 import host_lib  # noqa: E402

NAMES = [(" N  ", " N"), (" CA ", " C"), (" C  ", " C"), (" O  ", " O"), (" CB ", " C")]


def write_pdb(path, xyz):
    out = []
    for i, (x, y, z) in enumerate(xyz):
        nm, el = NAMES[i % 5]
        out.append("ATOM  %5d %s ALA A%4d    %8.3f%8.3f%8.3f  1.00  0.00          %s\n"
                   % ((i + 1) % 100000, nm, (i // 5 + 1) % 10000, x, y, z, el))
    out.append("END\n")
    with open(path, "w") as fh:
        fh.writelines(out)


def cpu_port_arm(ind, files, level, n_atoms):
    """The CPU port driven the same way: the same C++ reader + extraction on every host core (threads; ctypes releases the GIL),
    then the oracle's -O3 build with the reference's directory-mode threading (one structure per task on all cores), then one
    JSON text per file.  This is test / benchmark infrastructure: the product never links the oracle."""
    import json
    from concurrent.futures import ThreadPoolExecutor
    from oracle import load
    fast = load(fast=True)
    cores = len(os.sched_getaffinity(0))
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        packs = list(ex.map(lambda f: host_lib.pack(os.path.join(ind, f), level), files))
    t_parse = time.perf_counter() - t0
    off = np.cumsum([0] + [p["xyzr"].shape[0] for p in packs]).astype(np.uint64)
    soff = np.cumsum([0] + [p["seg_be"].shape[0] for p in packs]).astype(np.uint64)
    xyzr = np.concatenate([p["xyzr"] for p in packs])
    seg = np.concatenate([p["seg_be"] for p in packs])
    t1 = time.perf_counter()
    out = fast.run_batch(xyzr, off, 1.4, 100, 8, cores, seg_be=seg, struct_seg_off=soff, want_counts=False)
    t_engine = time.perf_counter() - t1
    t2 = time.perf_counter()
    texts = [json.dumps({"Residue": out["seg"][int(soff[i]):int(soff[i + 1])].tolist()}) for i in range(len(packs))]
    t_write = time.perf_counter() - t2
    dt = time.perf_counter() - t0
    print(f"cpu port arm: wall {dt:.2f} s  {n_atoms / dt / 1e6:.2f} M atoms/s on {cores} cores (parse+extract {t_parse:.2f} s, oracle "
          f"{t_engine:.2f} s, serialise {t_write:.2f} s; {sum(len(t) for t in texts) / 1e6:.0f} MB of JSON)", flush=True)
    return dt


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n_files = int(args[0]) if len(args) > 0 else 2000
    level = args[1] if len(args) > 1 else "residue"
    extra = []
    for k in ("--devices", "--tile"):
        if k in sys.argv:
            extra += [k, sys.argv[sys.argv.index(k) + 1]]
    data = W.proteome_batch(n_files, seed=W.SEED)
    with tempfile.TemporaryDirectory() as tmp:
        ind, outd = os.path.join(tmp, "in"), os.path.join(tmp, "out")
        os.mkdir(ind)
        t0 = time.perf_counter()
        for s in range(n_files):
            a0, a1 = int(data.struct_off[s]), int(data.struct_off[s + 1])
            write_pdb(os.path.join(ind, "s%05d.pdb" % s), data.xyzr[a0:a1, :3])
        nbytes = sum(os.path.getsize(os.path.join(ind, f)) for f in os.listdir(ind))
        print(f"wrote {n_files} files, {data.n_atoms} atoms, {nbytes / 1e6:.0f} MB in {time.perf_counter() - t0:.1f} s", flush=True)
        best = None
        for rep in range(3):   # later runs: page cache warm, context creation still inside
            env = dict(os.environ)
            if rep == 2:
                env["SASA_B200_TRACE"] = "1"
            t0 = time.perf_counter()
            r = subprocess.run([host_lib.CLI_PATH, ind, outd, "--format", "json", "-o", level] + extra, capture_output=True, text=True, env=env)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            tail = (r.stdout.strip().splitlines() or [""])[-1]
            print(f"run {rep}: rc={r.returncode} wall {dt:.2f} s  {data.n_atoms / dt / 1e6:.1f} M atoms/s incl. process start | cli: {tail}",
                  flush=True)
            if r.returncode != 0 or rep == 2:
                print(r.stderr[-600:])
        print("outputs:", len(os.listdir(outd)), "nproc", os.cpu_count())
        if "--cpu" in sys.argv:
            cdt = cpu_port_arm(ind, sorted(os.listdir(ind)), level, data.n_atoms)
            print(f"GPU CLI best wall {best:.2f} s vs CPU port {cdt:.2f} s: {cdt / best:.1f}x", flush=True)


if __name__ == "__main__":
    main()
