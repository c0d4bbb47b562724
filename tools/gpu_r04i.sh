#!/bin/bash
set -u
OUT=gpurun_out/r04i
mkdir -p $OUT
bash tools/gpu_check.sh r04i
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sasa_tight_kernel -s 3 -c 1 \
    -o $OUT/prof python bench.py --steps 1 --warmup 3 --no-cpu --no-secondary > $OUT/prof_bench.log 2>&1
ls -la $OUT | tail -4
