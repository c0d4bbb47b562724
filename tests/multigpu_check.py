#!/usr/bin/env python
"""Multi-GPU parity check, one process per GPU (not collected by pytest; run on a multi-GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_check.py

(1) a proteome batch sharded over the ranks (no data-path collective; results gathered to rank 0 over a gloo
    side group) equals the oracle;  (2) the atom-range split of one large structure, per-rank partial vectors
    summed with an NCCL all-reduce over NVLink, equals the single-GPU result and the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rustsasa_b200 import Engine, workloads as W  # noqa: E402
from rustsasa_b200.shard import run_atom_range, run_sharded  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    host_group = dist.new_group(backend="gloo")
    eng = Engine(local)

    # (1) structures sharded over ranks
    d = W.proteome_batch(64)

    def compute(sh):
        b = eng.batch(sh.struct_off, sh.seg_be, sh.struct_seg_off, sh.seg_polar)
        try:
            return b.run_host(sh.xyzr, sh.id_class)
        finally:
            b.close()
    res = run_sharded(compute, d.xyzr, d.struct_off, d.seg_be, d.struct_seg_off, d.seg_polar, host_group=host_group)
    if rank == 0:
        from oracle import load
        o = load(fast=True).run_batch(d.xyzr, d.struct_off, seg_be=d.seg_be, struct_seg_off=d.struct_seg_off)
        assert np.array_equal(res.counts, o["counts"]) and np.array_equal(res.seg_sasa, o["seg"])
        print(f"sharded batch over {world} GPUs: parity OK ({d.n_atoms} atoms, bounds {res.bounds.tolist()})", flush=True)

    # (2) atom-range split of one large structure + NCCL all-reduce
    a = W.capsid_shell(120000)
    b = eng.batch(a.struct_off, a.seg_be, a.struct_seg_off, a.seg_polar)
    d_xyzr = torch.from_numpy(a.xyzr).cuda()

    def compute_range(r, w):
        counts = torch.empty(a.n_atoms, dtype=torch.int32, device="cuda")
        atom = torch.empty(a.n_atoms, dtype=torch.float32, device="cuda")
        b.run_atom_range_device(d_xyzr, r, w, n_points=960, counts=counts, atom_sasa=atom)
        return counts, atom
    counts, atom = run_atom_range(compute_range)
    d_seg = torch.zeros(len(a.seg_be), dtype=torch.float32, device="cuda")
    b.reduce_device(atom, d_seg, None)
    torch.cuda.synchronize()
    b.sync()
    single = b.run_host(a.xyzr, n_points=960, want=("counts", "atom", "seg"))
    assert np.array_equal(counts.cpu().numpy().view(np.uint32), single.counts)
    assert np.array_equal(atom.cpu().numpy(), single.atom_sasa)
    assert np.array_equal(d_seg.cpu().numpy(), single.seg_sasa)
    if rank == 0:
        from oracle import load
        o = load(fast=True).calculate_sasa_internal(a.xyzr, 1.4, 960, threads=-1)
        assert np.array_equal(single.counts, o["counts"])
        print(f"atom-range split over {world} GPUs + all-reduce: parity OK ({a.n_atoms} atoms, 960 points)", flush=True)
    b.close()

    # (3) BASELINE cfg5 at full size: 1 M atoms x 960 points, atom ranges split over all ranks, against the oracle's
    # fingerprint (tests/golden/cfg_hashes.json, written by tools/make_cfg_hashes.py) and the single-GPU run.  Work blocks are
    # cut at cell boundaries, so the per-rank partitions agree although the order of atoms inside a cell differs per GPU.
    import hashlib
    import json
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "cfg_hashes.json")))["cfg5"]
    a = W.capsid_shell(1000000)
    assert a.n_atoms == gold["atoms"]
    b = eng.batch(a.struct_off)
    d_xyzr = torch.from_numpy(a.xyzr).cuda()

    def compute_range_big(r, w):
        counts = torch.empty(a.n_atoms, dtype=torch.int32, device="cuda")
        atom = torch.empty(a.n_atoms, dtype=torch.float32, device="cuda")
        b.run_atom_range_device(d_xyzr, r, w, n_points=960, counts=counts, atom_sasa=atom)
        return counts, atom
    single = b.run_host(a.xyzr, n_points=960, want=("counts", "atom"))
    assert hashlib.sha256(np.ascontiguousarray(single.counts, dtype="<u4").tobytes()).hexdigest() == gold["sha256_counts"]
    owned = []
    for _ in range(3):
        counts, atom = run_atom_range(compute_range_big)
        torch.cuda.synchronize()
        assert np.array_equal(counts.cpu().numpy().view(np.uint32), single.counts)
        assert np.array_equal(atom.cpu().numpy(), single.atom_sasa)
    # the same with the exchange fused into the kernel: every rank writes its atoms' values straight into all ranks' vectors
    # (torch symmetric memory: NVLink peer pointers), no zero-fill and no all-reduce
    from rustsasa_b200.shard import PeerVectors, run_atom_range_peers
    peers = PeerVectors(a.n_atoms)
    for _ in range(3):
        pc, pa = run_atom_range_peers(
            lambda r, w, pv: b.run_atom_range_peers_device(d_xyzr, r, w, pv.count_ptrs, pv.atom_ptrs, n_points=960), peers)
        torch.cuda.synchronize()
        assert np.array_equal(pc.cpu().numpy().view(np.uint32), single.counts)
        assert np.array_equal(pa.cpu().numpy(), single.atom_sasa)
        pc.zero_()
        pa.zero_()
        torch.cuda.synchronize()
    if rank == 0:
        print(f"atom-range split with peer writes over {world} GPUs: equals the single-GPU result (3 runs)", flush=True)
    # balance of the interleaved ownership: atoms evaluated by this rank
    mine, _ = compute_range_big(rank, world)
    part = b.run_atom_range_host(a.xyzr, rank, world, n_points=960)
    n_mine = torch.tensor([float(((part.counts != 0) | (part.atom_sasa != 0)).sum())], device="cuda")
    lo, hi = n_mine.clone(), n_mine.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"atom-range split of {a.n_atoms} atoms x 960 points over {world} GPUs: equals the oracle fingerprint and the single-GPU "
              f"result (3 runs); exposed atoms per rank {int(lo)}..{int(hi)}", flush=True)
    b.close()
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
