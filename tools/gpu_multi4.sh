#!/bin/bash
# N-GPU bench line only (weak + secondary block).  usage (under gpurun --gpus N): bash tools/gpu_multi4.sh N tag
set -u
N=${1:-2}
TAG=${2:-multi4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu 2> $OUT/bench_${N}gpu.err | tail -1 ) > $OUT/bench_${N}gpu.json
python - <<PY
import json
d=json.load(open("$OUT/bench_${N}gpu.json"))
print("N=%d value %.1f e2e %.1f e2e_float4 %.1f"%(d["n_gpus"],d["value"]/1e6,d["e2e"]["value"]/1e6,d.get("e2e_float4",d["e2e"])["value"]/1e6))
print(d.get("per_rank_ms_per_step"))
for k,v in d.get("secondary",{}).items():
    print(k, {kk:(round(vv/1e6,1) if kk=="value" else vv) for kk,vv in v.items() if kk in ("value","ms_per_step","parity","error")}, "e2e", (v.get("e2e") or {}).get("value"))
PY
