#!/bin/bash
set -u
TAG=${1:-r02s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q --timeout=300 2>&1 | tail -15 ) > $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
python - <<PY
import time, numpy as np, sys
sys.path.insert(0, ".")
from rustsasa_b200 import Engine, workloads as W
eng = Engine(0)
d = W.proteome_batch(400, seed=5)
b = eng.batch(d.struct_off, d.seg_be, d.struct_seg_off, d.seg_polar)
for n in (100, 960):
    b.run_host(d.xyzr, n_points=n, want=("seg",))
    t0 = time.perf_counter()
    for _ in range(3):
        r = b.run_host(d.xyzr, n_points=n, want=("seg",))
    dt = (time.perf_counter() - t0) / 3
    print(f"400 structures ({d.n_atoms} atoms) at {n} points: {dt*1e3:.2f} ms  {d.n_atoms/dt/1e6:.0f} M atoms/s (host buffers)")
PY
