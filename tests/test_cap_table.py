"""CPU check of the cap tables (rustsasa_b200/csrc/sasa_cap.cuh): the host-side builders are compiled into a small checker
(tests/native/cap_table_check.cu, nvcc as host compiler -- no GPU) that samples random directions and levels, looks the bin up
with the kernel's binning arithmetic and verifies by brute force in double precision that an inner bit always means
"occluded" and that every occluded point is in the inner or the ring mask.  That is the property the bit-exactness of the
cap-table occlusion path rests on; the GPU parity tests check its consequence, this checks it directly,
including directions on the folds of the octahedral map and levels exactly on bin boundaries."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "cap_table_check.cu")


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = shutil.which("nvcc")
    if not nvcc:
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("cap") / "cap_table_check")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-o", exe, SRC],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    assert r.returncode == 0, r.stdout
    return exe


@pytest.mark.parametrize("n_points,samples", [(100, 300000), (37, 60000), (128, 60000), (1, 20000), (960, 40000), (200, 40000)])
def test_tables_never_decide_a_point_wrongly(checker, n_points, samples):
    r = subprocess.run([checker, str(n_points), str(samples)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout
    assert "wrong_inner=0 uncovered=0" in r.stdout, r.stdout
