#!/bin/bash
set -u
TAG=${1:-r02e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q --timeout=300 2>&1 | tail -15 ) > $OUT/pytest_gpu.log
( time timeout 600 python bench.py --steps 10 --warmup 3 2> $OUT/bench.err | tail -1 ) > $OUT/bench.json 2> $OUT/bench.time
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2> $OUT/bench_ref.err | tail -1 ) > $OUT/bench_ref.json 2> $OUT/bench_ref.time
tail -4 $OUT/pytest_gpu.log
cat $OUT/bench.time $OUT/bench_ref.time
tail -5 $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("value %.1f e2e %.1f frac %.3f"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["roofline"]["frac"]))
for k,v in d.get("secondary",{}).items():
    print(k, {kk:(round(vv/1e6,1) if kk=="value" else vv) for kk,vv in v.items() if kk in ("value","ms_per_step","parity","error","gather_ms")}, "e2e", v.get("e2e"))
r=json.load(open("$OUT/bench_ref.json")); print("ref", r["value"]/1e6, r["config"]==d["config"])
PY
