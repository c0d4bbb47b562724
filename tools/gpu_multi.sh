#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): N-rank parity check, the N-GPU bench line and the cfg5 atom-range split.
# usage: bash tools/gpu_multi.sh N [tag]
set -u
N=${1:-2}
TAG=${2:-multi}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( timeout 600 $TR --master-port 29511 tests/multigpu_check.py 2>&1 | grep -v "^\*\*\|OMP_NUM_THREADS\|^$" | tail -12 ) > $OUT/multigpu_${N}gpu.log
( timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu 2> $OUT/bench_${N}gpu.err | tail -1 ) > $OUT/bench_${N}gpu.json
( timeout 600 $TR --master-port 29513 tools/bench_configs.py cfg5 2> $OUT/cfg5_${N}gpu.err | tail -1 ) > $OUT/cfg5_${N}gpu.json
ls -la $OUT
