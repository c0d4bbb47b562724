"""Host-side multi-GPU logic on CPU: partitioning, shard slicing and the world_size-2 gloo gather.

The compute step is injected: here it is the oracle (the checker), on the GPU box it is `Batch.run_host`
(see tests/test_gpu_parity.py::test_sharded_run_matches_single and bench.py)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_is_contiguous_balanced_and_complete():
    from rustsasa_b200.shard import partition_structures, structure_cost
    rng = np.random.default_rng(1)
    sizes = rng.integers(300, 6000, size=997)
    off = np.concatenate([[0], np.cumsum(sizes)])
    for parts in (1, 2, 3, 4, 8):
        b = partition_structures(off, parts)
        assert b[0] == 0 and b[-1] == len(sizes) and np.all(np.diff(b) >= 0)
        cost = structure_cost(off)
        per = np.array([cost[b[r]:b[r + 1]].sum() for r in range(parts)])
        assert per.max() / per.mean() < 1.02           # within one structure of the ideal split
    # fewer structures than ranks: empty ranges, nothing lost
    b = partition_structures([0, 10, 30], 8)
    assert b[0] == 0 and b[-1] == 2 and np.all(np.diff(b) >= 0)
    assert partition_structures([0], 4).tolist() == [0, 0, 0, 0, 0]
    with pytest.raises(ValueError):
        partition_structures([0, 1], 0)


def test_take_shard_rebases_offsets():
    from rustsasa_b200 import workloads as W
    from rustsasa_b200.shard import partition_structures, take_shard
    d = W.proteome_batch(12)
    b = partition_structures(d.struct_off, 3)
    seen_atoms = seen_segs = 0
    for r in range(3):
        sh = take_shard(b, r, d.xyzr, d.struct_off, d.seg_be, d.struct_seg_off, d.seg_polar)
        assert sh.struct_off[0] == 0 and sh.struct_off[-1] == sh.xyzr.shape[0] == sh.a1 - sh.a0
        assert sh.struct_seg_off[0] == 0 and sh.struct_seg_off[-1] == sh.seg_be.shape[0]
        assert np.array_equal(sh.xyzr, d.xyzr[sh.a0:sh.a1])
        assert np.array_equal(sh.seg_be, d.seg_be[sh.g0:sh.g1])          # ranges are structure-relative: unchanged
        seen_atoms += sh.xyzr.shape[0]
        seen_segs += sh.seg_be.shape[0]
    assert seen_atoms == d.n_atoms and seen_segs == d.seg_be.shape[0]


WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["SASA_ROOT"])
from types import SimpleNamespace
from oracle import load
from rustsasa_b200 import workloads as W
from rustsasa_b200.shard import run_sharded

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
d = W.proteome_batch(int(os.environ["SASA_NSTRUCT"]))
orc = load()

def compute(sh):   # the oracle stands in for the engine call on this CPU-only host
    o = orc.run_batch(sh.xyzr, sh.struct_off, seg_be=sh.seg_be, struct_seg_off=sh.struct_seg_off)
    prot = np.stack([orc.protein_totals(o["sasa"][int(sh.struct_off[s]):int(sh.struct_off[s + 1])],
                                        sh.seg_be[int(sh.struct_seg_off[s]):int(sh.struct_seg_off[s + 1])],
                                        sh.seg_polar[int(sh.struct_seg_off[s]):int(sh.struct_seg_off[s + 1])])
                     for s in range(len(sh.struct_off) - 1)]) if len(sh.struct_off) > 1 else np.zeros((0, 3), np.float32)
    return SimpleNamespace(counts=o["counts"], atom_sasa=o["sasa"], seg_sasa=o["seg"], protein=prot)

res = run_sharded(compute, d.xyzr, d.struct_off, d.seg_be, d.struct_seg_off, d.seg_polar)
if rank == 0:
    full = orc.run_batch(d.xyzr, d.struct_off, seg_be=d.seg_be, struct_seg_off=d.struct_seg_off)
    assert np.array_equal(res.counts, full["counts"])
    assert np.array_equal(res.atom_sasa, full["sasa"])
    assert np.array_equal(res.seg_sasa, full["seg"])
    assert res.protein.shape == (d.n_structures, 3)
    assert abs(float(res.protein[:, 0].sum()) - float(full["sasa"].sum())) < 1e-3 * float(full["sasa"].sum())
    print("GATHER_OK", world, res.bounds.tolist())
else:
    assert res.counts is None
dist.barrier()
dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,nstruct", [(2, 9), (2, 1)])
def test_world_size_2_gloo_gather_matches_single_process(tmp_path, world, nstruct):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SASA_ROOT=ROOT, SASA_NSTRUCT=str(nstruct), MASTER_ADDR="127.0.0.1",
               MASTER_PORT=str(_free_port()), WORLD_SIZE=str(world), OMP_NUM_THREADS="2")
    procs = []
    for r in range(world):
        e = dict(env, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "GATHER_OK" in outs[0]
