#!/bin/bash
# 8-GPU visit: bench line (per-rank e2e, indexed wire format, secondary block) + the CLI over all devices.
set -u
N=${1:-8}
TAG=${2:-multi3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 2> $OUT/bench_${N}gpu.err | tail -1 ) > $OUT/bench_${N}gpu.json 2> $OUT/bench.time
( timeout 600 python tools/cli_dir_bench.py 4400 residue --devices all 2>&1 | tail -8 ) > $OUT/cli_4400_alldev.log
( timeout 600 python -m pytest tests/test_gpu_multi.py -x -q --timeout=600 2>&1 | tail -3 ) > $OUT/pytest_multi.log
cat $OUT/bench.time $OUT/pytest_multi.log $OUT/cli_4400_alldev.log
python - <<PY
import json
d=json.load(open("$OUT/bench_${N}gpu.json"))
print("N=%d value %.1f e2e %.1f e2e_float4 %.1f"%(d["n_gpus"],d["value"]/1e6,d["e2e"]["value"]/1e6,d.get("e2e_float4",d["e2e"])["value"]/1e6))
print(d.get("per_rank_ms_per_step"))
for k,v in d.get("secondary",{}).items():
    print(k, {kk:(round(vv/1e6,1) if kk=="value" else vv) for kk,vv in v.items() if kk in ("value","ms_per_step","parity","error","gather_ms","kernels_only","wall","peer_writes")}, "e2e", v.get("e2e"))
PY
