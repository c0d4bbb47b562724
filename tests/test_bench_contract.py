"""bench.py's reference arm on CPU: the JSON line's contract keys, and the regression that made the N > 1 reference arm
single-threaded (torchrun exports OMP_NUM_THREADS=1 to its workers; the arm must still use every host core)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_ref(env_extra, *args):
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--structures", "60", *args], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-800:]
    return r.stdout.strip().splitlines()


def test_reference_arm_line_and_thread_count():
    lines = run_ref({"OMP_NUM_THREADS": "1"})
    d = json.loads(lines[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "atoms/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["e2e"] == {"value": d["value"], "unit": "atoms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and "sample" in cb
    assert cb["cores"] == len(os.sched_getaffinity(0))          # not 1, whatever OMP_NUM_THREADS says
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    assert run_ref({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2") == []
