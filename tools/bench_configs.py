#!/usr/bin/env python
"""Secondary measurements for BASELINE.json configs 3-5 (bench.py owns config 2, the headline).

    python tools/bench_configs.py [cfg3] [cfg4] [cfg5] [--frames N]        # one GPU
    torchrun --nproc-per-node N tools/bench_configs.py cfg5                # atom ranges split over N GPUs (NCCL all-reduce)

Each line is a JSON object: atoms/s through the host C-ABI call (pinned buffers, copies inside the timed region) and,
where a device-resident form exists, the kernel-only figure; a bounded oracle sample gives parity and the CPU rate."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rustsasa_b200 import Engine, workloads as W  # noqa: E402
from rustsasa_b200.engine import BatchResult  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def cfg3(eng, n_frames):
    """MD trajectory: frames of one ~5,000-atom protein, ProteinLevel, 12 B/atom/frame on the wire."""
    from oracle import load
    md = W.md_trajectory(n_frames=n_frames, n_atoms=5000)
    F, N = md.xyz.shape[:2]
    off = (np.arange(F + 1, dtype=np.uint64) * N)
    G = len(md.seg_be)
    b = eng.batch(off, np.tile(md.seg_be, (F, 1)), (np.arange(F + 1, dtype=np.uint64) * G), np.tile(md.seg_polar, F))
    h_xyz = eng.pinned_empty((F * N, 3), np.float32)
    h_xyz[...] = md.xyz.reshape(-1, 3)
    res = BatchResult(protein=eng.pinned_empty((F, 3), np.float32))
    dt = timed(lambda: b.run_frames_host(h_xyz, md.radii, want=("protein",), result=res), reps=3, warm=1)
    # kernels alone, frames resident in HBM as float4
    d_xyzr = torch.from_numpy(np.concatenate([md.xyz.reshape(-1, 3), np.tile(md.radii, F)[:, None]], axis=1).astype(np.float32)).cuda()
    d_prot = torch.zeros((F, 3), dtype=torch.float32, device="cuda")
    kdt = timed(lambda: b.run_device(d_xyzr, protein=d_prot), reps=3, warm=1)
    del d_xyzr
    fast = load(fast=True)
    m = min(F, 64)
    xyzr = np.concatenate([md.xyz[:m].reshape(-1, 3), np.tile(md.radii, m)[:, None]], axis=1).astype(np.float32)
    t0 = time.perf_counter()
    o = fast.run_batch(xyzr, off[:m + 1], seg_be=np.tile(md.seg_be, (m, 1)), struct_seg_off=(np.arange(m + 1, dtype=np.uint64) * G))
    cpu_dt = time.perf_counter() - t0
    want = np.stack([fast.protein_totals(o["sasa"][f * N:(f + 1) * N], md.seg_be, md.seg_polar) for f in range(m)])
    b.close()
    return dict(config="cfg3 MD trajectory", frames=F, atoms_per_frame=N, level="protein", n_points=100,
                e2e_atoms_per_s=F * N / dt, e2e_frames_per_s=F / dt, ms_per_call=dt * 1e3, h2d_bytes=F * N * 12, d2h_bytes=F * 12,
                device_atoms_per_s=F * N / kdt, device_ms=kdt * 1e3,
                parity_first_frames=bool(np.array_equal(res.protein[:m], want)), cpu_atoms_per_s=m * N / cpu_dt,
                cpu_cores=fast.max_threads())


def single(eng, data, n_points, label, sample_cpu=True):
    from oracle import load
    b = eng.batch(data.struct_off)
    N = data.n_atoms
    h = eng.pinned_empty((N, 4), np.float32)
    h[...] = data.xyzr
    res = BatchResult(atom_sasa=eng.pinned_empty(N, np.float32), counts=eng.pinned_empty(N, np.uint32))
    dt = timed(lambda: b.run_host(h, n_points=n_points, want=("counts", "atom"), result=res))
    d_xyzr = torch.from_numpy(data.xyzr).cuda()
    d_atom = torch.zeros(N, dtype=torch.float32, device="cuda")
    kdt = timed(lambda: b.run_device(d_xyzr, n_points=n_points, atom_sasa=d_atom))
    out = dict(config=label, atoms=N, level="atom", n_points=n_points, e2e_atoms_per_s=N / dt, ms_per_call=dt * 1e3,
               device_atoms_per_s=N / kdt, device_ms=kdt * 1e3, launches=b.sync()["gpu_launches"],
               sum_counts=int(np.asarray(res.counts, dtype=np.int64).sum()))
    if sample_cpu:
        fast = load(fast=True)
        t0 = time.perf_counter()
        o = fast.calculate_sasa_internal(data.xyzr, 1.4, n_points, threads=-1)
        cdt = time.perf_counter() - t0
        out.update(parity=bool(np.array_equal(np.asarray(res.counts), o["counts"])), cpu_atoms_per_s=N / cdt,
                   cpu_cores=fast.max_threads(), cpu_note="serial neighbour build + atoms over all cores, like src/lib.rs:278-290")
    b.close()
    return out


def cfg5_split(eng, data, n_points, rank, world):
    import torch.distributed as dist
    from rustsasa_b200.shard import run_atom_range
    b = eng.batch(data.struct_off)
    N = data.n_atoms
    d_xyzr = torch.from_numpy(data.xyzr).cuda()
    counts = torch.empty(N, dtype=torch.int32, device="cuda")
    atom = torch.empty(N, dtype=torch.float32, device="cuda")

    def step():
        run_atom_range(lambda r, w: (b.run_atom_range_device(d_xyzr, r, w, n_points=n_points, counts=counts, atom_sasa=atom),
                                     (counts, atom))[1], rank, world)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 5], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total = int(counts.sum().item())
    b.close()
    return dict(config="cfg5 capsid, atom ranges split + NCCL all-reduce", atoms=N, n_points=n_points, n_gpus=world,
                device_ms=float(t[0]), device_atoms_per_s=N / (float(t[0]) * 1e-3), sum_counts=total)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    frames = int(sys.argv[sys.argv.index("--frames") + 1]) if "--frames" in sys.argv else 10000
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    eng = Engine(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        r = cfg5_split(eng, W.capsid_shell(1000000), 960, rank, world)
        if rank == 0:
            print(json.dumps(r), flush=True)
        dist.destroy_process_group()
        return
    want = args or ["cfg3", "cfg4", "cfg5"]
    if "cfg3" in want:
        print(json.dumps(cfg3(eng, frames)), flush=True)
    if "cfg4" in want:
        print(json.dumps(single(eng, W.large_assembly(150000), 100, "cfg4 150k-atom assembly")), flush=True)
    if "cfg5" in want:
        print(json.dumps(single(eng, W.capsid_shell(1000000), 960, "cfg5 1M-atom capsid, 960 points (one GPU)", sample_cpu="--cpu5" in sys.argv)),
              flush=True)


if __name__ == "__main__":
    main()
