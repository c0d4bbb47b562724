#!/bin/bash
# One ncu --set full capture of the dominant kernel on the bench workload (+ the plain bench line next to it).
# usage (under gpurun): bash tools/gpu_prof.sh [tag]
set -u
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2> $OUT/bench.err | tail -1 ) > $OUT/bench.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sasa_tight_kernel -s 3 -c 1 \
    -o $OUT/prof python bench.py --steps 1 --warmup 3 --no-cpu > $OUT/prof_bench.log 2>&1
ls -la $OUT
