"""Multi-GPU parity as a collected test: launches tests/multigpu_check.py under torchrun on every visible GPU (2..8), one
process per GPU over NCCL.  Skipped on single-GPU boxes; the CPU (gloo, world_size 2) coverage of the same host logic is
tests/test_shard_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpus() < 2, reason="needs at least two GPUs")
def test_multigpu_check_on_all_gpus():
    n = min(_gpus(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "tests", "multigpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:]
    assert f"sharded batch over {n} GPUs: parity OK" in r.stdout
    assert f"over {n} GPUs + all-reduce: parity OK" in r.stdout
    assert f"over {n} GPUs: equals the oracle fingerprint" in r.stdout
    assert f"peer writes over {n} GPUs: equals the single-GPU result" in r.stdout
