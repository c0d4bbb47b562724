"""The gated single-launch host pipeline as a collected GPU test: tests/gated_check.py in a subprocess with small chunks
(SASA_B200_CHUNK_ATOMS is read once per process), so that batches of a few hundred structures already span many chunks."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("chunk_atoms", ["8000", "30000"])
def test_gated_pipeline_matches_per_chunk_launches_and_oracle(chunk_atoms):
    env = dict(os.environ)
    env["SASA_B200_CHUNK_ATOMS"] = chunk_atoms
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "gated_check.py")], cwd=ROOT, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:]
    assert "ALL OK" in r.stdout and "FAIL " not in r.stdout
