#!/bin/bash
# variants + one full ncu capture (with source) of the variant named by $1
set -u
bash tools/gpu_variants.sh r04c
export SASA_B200_LIB=$PWD/rustsasa_b200/variants/libsasa_b200_${1:-best}.so
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sasa_tight_kernel -s 3 -c 1 \
    -o gpurun_out/r04c/prof python bench.py --steps 1 --warmup 3 --no-cpu --no-secondary > gpurun_out/r04c/prof_bench.log 2>&1
ls -la gpurun_out/r04c | tail -5
