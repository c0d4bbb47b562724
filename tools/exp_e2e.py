#!/usr/bin/env python
"""Where the end-to-end call spends its time: device-side span (first enqueue to last event) and host wall clock of
sasa_b200_batch_run_indexed_host on the bench workload, next to the device-resident kernel time.
usage (GPU box): [SASA_B200_CHUNK_ATOMS=..] python tools/exp_e2e.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rustsasa_b200 import Engine, workloads as W  # noqa: E402
from rustsasa_b200.engine import BatchResult, index_radii  # noqa: E402

data = W.proteome_batch(4400, seed=W.SEED)
eng = Engine(0)
b = eng.batch(data.struct_off, data.seg_be, data.struct_seg_off, data.seg_polar)
N, G = data.n_atoms, int(data.seg_be.shape[0])
pal, idx = index_radii(data.xyzr[:, 3])
h_xyz = eng.pinned_empty((N, 3), np.float32)
h_xyz[...] = data.xyzr[:, :3]
h_idx = eng.pinned_empty((N,), np.uint8)
h_idx[...] = idx
res = BatchResult(seg_sasa=eng.pinned_empty((G,), np.float32))
for _ in range(3):
    b.run_indexed_host(h_xyz, h_idx, pal, want=("seg",), result=res)
ks, ts, ws = [], [], []
for _ in range(10):
    t0 = time.perf_counter()
    r = b.run_indexed_host(h_xyz, h_idx, pal, want=("seg",), result=res)
    ws.append((time.perf_counter() - t0) * 1e3)
    ks.append(r.stats["kernel_ms"])
    ts.append(r.stats["total_ms"])
h_x4 = eng.pinned_empty((N, 4), np.float32)
h_x4[...] = data.xyzr
for _ in range(3):
    b.run_host(h_x4, want=("seg",), result=res)
k4 = []
for _ in range(10):
    k4.append(b.run_host(h_x4, want=("seg",), result=res).stats["kernel_ms"])
print(f"float4 wire format: device span {np.median(k4):.3f} ms")
d_xyzr = torch.from_numpy(data.xyzr).cuda()
d_seg = torch.zeros(G, dtype=torch.float32, device="cuda")
for _ in range(3):
    b.run_device(d_xyzr, seg_sasa=d_seg)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    b.run_device(d_xyzr, seg_sasa=d_seg)
e1.record()
torch.cuda.synchronize()
print(f"chunk={os.environ.get('SASA_B200_CHUNK_ATOMS', 'default')}: device-resident {e0.elapsed_time(e1) / 10:.3f} ms | host call: device span "
      f"{np.median(ks):.3f} ms, library wall {np.median(ts):.3f} ms, python wall {np.median(ws):.3f} ms, launches {r.stats['gpu_launches']}")
