#!/bin/bash
# parity tests + the secondary configurations (cfg3/4/5) after a kernel change.  usage (under gpurun): bash tools/gpu_r02b.sh [tag]
set -u
TAG=${1:-r02b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q --timeout=300 2>&1 | tail -25 ) > $OUT/pytest_gpu.log
( timeout 600 python tools/bench_configs.py cfg3 cfg4 cfg5 --cpu5 --frames 4000 2> $OUT/configs.err ) > $OUT/configs.jsonl
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2> $OUT/bench.err | tail -1 ) > $OUT/bench.json
( timeout 120 python tools/latency_single.py 2>&1 | tail -8 ) > $OUT/latency_single.log
tail -5 $OUT/pytest_gpu.log
cat $OUT/configs.jsonl $OUT/latency_single.log
tail -3 $OUT/configs.err
