#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): the collected multi-GPU test, then the N-GPU bench line (with its secondary block).
set -u
N=${1:-2}
TAG=${2:-multi}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
nproc > $OUT/nproc.txt
( timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -s --timeout=900 2>&1 | tail -25 ) > $OUT/pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 2> $OUT/bench_${N}gpu.err | tail -1 ) > $OUT/bench_${N}gpu.json 2> $OUT/bench.time
tail -12 $OUT/pytest_multi.log
cat $OUT/bench.time
grep -v "^\*\*\|OMP_NUM\|^$" $OUT/bench_${N}gpu.err | tail -5
python - <<PY
import json
d=json.load(open("$OUT/bench_${N}gpu.json"))
print("N=%d value %.1f e2e %.1f binding %s"%(d["n_gpus"],d["value"]/1e6,d["e2e"]["value"]/1e6,d.get("host_binding")))
for k,v in d.get("secondary",{}).items():
    print(k, {kk:(round(vv/1e6,1) if kk=="value" else vv) for kk,vv in v.items() if kk in ("value","ms_per_step","parity","error","gather_ms","kernels_only","wall","peer_writes")}, "e2e", v.get("e2e"))
PY
