#!/usr/bin/env python
"""Experiment: the same proteome batch, clipped so that every structure fits the 512-thread x 2-CTA/SM configuration,
run device-resident under different SASA_B200_CFGS settings (which shared-memory configuration the structures land in).

usage (GPU box): [EXP_MEAN_ATOMS=800 EXP_STRUCTURES=12000] exp_cfg.py [hi_atoms] [cfgs ...]        e.g. exp_cfg.py 2590 12 2 1
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rustsasa_b200 import Engine  # noqa: E402
from rustsasa_b200 import workloads as W  # noqa: E402

hi = int(sys.argv[1]) if len(sys.argv) > 1 else 2590
cfgs = sys.argv[2:] or ["12", "2"]
mean = float(os.environ.get("EXP_MEAN_ATOMS", "2400"))
nstruct = int(os.environ.get("EXP_STRUCTURES", "4400"))
data = W.proteome_batch(nstruct, seed=W.SEED, hi=hi, mean_atoms=mean, sd_atoms=mean / 6.0, lo=min(400, int(mean / 2)))
N, G = data.n_atoms, int(data.seg_be.shape[0])
d_xyzr = torch.from_numpy(data.xyzr).cuda()
ref = None
for c in cfgs:
    os.environ["SASA_B200_CFGS"] = c
    eng = Engine(0)
    batch = eng.batch(data.struct_off, data.seg_be, data.struct_seg_off, data.seg_polar)
    d_seg = torch.zeros(G, dtype=torch.float32, device="cuda")
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):
            batch.run_device(d_xyzr, seg_sasa=d_seg, probe_radius=1.4, n_points=100, stream=st.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        steps = 8
        for _ in range(steps):
            batch.run_device(d_xyzr, seg_sasa=d_seg, probe_radius=1.4, n_points=100, stream=st.cuda_stream)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    info = batch.sync()
    out = d_seg.cpu().numpy()
    if ref is None:
        ref = out
    print(f"hi={hi} CFGS={c:4s} {N / ms / 1e3:8.1f} M atoms/s  {ms:7.3f} ms/step  launches/step={info['gpu_launches']}  "
          f"same_as_first={bool(np.array_equal(ref, out))}", flush=True)
    batch.close()
    eng.close()
