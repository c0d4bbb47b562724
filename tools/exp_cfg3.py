#!/usr/bin/env python
"""Where does cfg3's end-to-end time go?  frames (12 B/atom) vs float4 (16 B/atom) host calls, the device span reported by the
library, and the plain pinned H2D bandwidth of the box.   usage (GPU box): python tools/exp_cfg3.py [frames]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rustsasa_b200 import Engine, workloads as W  # noqa: E402
from rustsasa_b200.engine import BatchResult  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
eng = Engine(0)
md = W.md_trajectory(n_frames=F, n_atoms=5000)
N = md.xyz.shape[1]
G = len(md.seg_be)
off = np.arange(F + 1, dtype=np.uint64) * N
b = eng.batch(off, np.tile(md.seg_be, (F, 1)), np.arange(F + 1, dtype=np.uint64) * G, np.tile(md.seg_polar, F))
h3 = eng.pinned_empty((F * N, 3), np.float32)
h3[...] = md.xyz.reshape(-1, 3)
h4 = eng.pinned_empty((F * N, 4), np.float32)
h4[:, :3] = md.xyz.reshape(-1, 3)
h4[:, 3] = np.tile(md.radii, F)
res = BatchResult(protein=eng.pinned_empty((F, 3), np.float32))


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        st = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, st


for name, fn in (("frames 12 B/atom", lambda: b.run_frames_host(h3, md.radii, want=("protein",), result=res).stats),
                 ("float4 16 B/atom", lambda: b.run_host(h4, want=("protein",), result=res).stats)):
    ms, st = timed(fn)
    print(f"{name}: wall {ms:.2f} ms  device span {st['kernel_ms']:.2f} ms  launches {st['gpu_launches']}  "
          f"{F * N / ms / 1e3:.0f} M atoms/s", flush=True)
d = torch.empty(F * N * 3, dtype=torch.float32, device="cuda")
src = torch.from_numpy(np.asarray(h3).reshape(-1))
for nbytes in (16 << 20, 64 << 20, d.numel() * 4):
    n = nbytes // 4
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        d[:n].copy_(src[:n], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"H2D {nbytes / 1e6:.0f} MB from the pinned buffer: {nbytes / dt / 1e9:.1f} GB/s", flush=True)
dd = torch.from_numpy(np.asarray(h4)).cuda()
dp = torch.zeros((F, 3), dtype=torch.float32, device="cuda")
for _ in range(2):
    b.run_device(dd, protein=dp)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    b.run_device(dd, protein=dp)
torch.cuda.synchronize()
print(f"device-resident float4: {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms", flush=True)
for env in ("2000000", "4000000"):
    pass
