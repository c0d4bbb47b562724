#!/bin/bash
set -u
TAG=${1:-r02d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q --timeout=300 2>&1 | tail -25 ) > $OUT/pytest_gpu.log
g++ -O2 -std=c++17 -pthread -Iinclude tools/abi_latency.cpp -Lrustsasa_b200 -lsasa_b200 -Wl,-rpath,$PWD/rustsasa_b200 -o /tmp/abi_latency
for n in 1283 2622 6065 32500; do /tmp/abi_latency $n 16 300; done > $OUT/abi_latency.jsonl 2>&1
/tmp/abi_latency 2622 4 300 >> $OUT/abi_latency.jsonl 2>&1
/tmp/abi_latency 2622 32 300 >> $OUT/abi_latency.jsonl 2>&1
( timeout 120 python tools/latency_single.py 2>&1 | tail -8 ) > $OUT/latency_single.log
( timeout 600 python tools/bench_configs.py cfg4 cfg5 2> $OUT/configs.err ) > $OUT/configs.jsonl
tail -8 $OUT/pytest_gpu.log
cat $OUT/abi_latency.jsonl $OUT/latency_single.log $OUT/configs.jsonl
