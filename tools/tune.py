#!/usr/bin/env python
"""Run bench.py under a list of environment settings (kernel tuning knobs) and print value / e2e per setting.

usage: tune.py "SASA_B200_CFGS=023" "SASA_B200_NEAR=3.5 SASA_B200_BCAST_MIN=12" ...
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for spec in sys.argv[1:] or [""]:
    env = dict(os.environ)
    for kv in spec.split():
        k, v = kv.split("=", 1)
        env[k] = v
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu", "--steps", "4", "--warmup", "3"],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"{spec or '(default)':50s} value {d['value'] / 1e6:9.1f} M atoms/s  {d['ms_per_step']:7.3f} ms/step   "
              f"e2e {d['e2e']['value'] / 1e6:9.1f} M atoms/s  same={d['results_identical_device_vs_host_leg']}", flush=True)
    except Exception as e:
        print(spec, "FAILED", e, r.stderr[-500:], flush=True)
