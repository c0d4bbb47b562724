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


def main():
    n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    level = sys.argv[2] if len(sys.argv) > 2 else "residue"
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
        for rep in range(2):   # second run: page cache warm, context creation still inside
            t0 = time.perf_counter()
            r = subprocess.run([host_lib.CLI_PATH, ind, outd, "--format", "json", "-o", level], capture_output=True, text=True)
            dt = time.perf_counter() - t0
            tail = (r.stdout.strip().splitlines() or [""])[-1]
            print(f"run {rep}: rc={r.returncode} wall {dt:.2f} s  {data.n_atoms / dt / 1e6:.1f} M atoms/s incl. process start | cli: {tail}",
                  flush=True)
            if r.returncode != 0:
                print(r.stderr[-400:])
        print("outputs:", len(os.listdir(outd)), "nproc", os.cpu_count())


if __name__ == "__main__":
    main()
