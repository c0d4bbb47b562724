#!/bin/bash
# Round-2 first visit: baseline numbers and ncu captures for the NON-headline kernels (cfg3 frames through the fused kernel,
# cfg4 / cfg5 through large_atoms_kernel), plus cfg5 full-size oracle parity.  usage (under gpurun): bash tools/gpu_r02a.sh [tag]
set -u
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nproc > $OUT/nproc.txt
( timeout 600 python tools/bench_configs.py cfg3 cfg4 cfg5 --cpu5 --frames 4000 2> $OUT/configs.err ) > $OUT/configs.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sasa_tight_kernel -s 2 -c 1 \
    -o $OUT/prof_cfg3 python tools/bench_configs.py cfg3 --frames 1500 > $OUT/prof_cfg3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:large_atoms_kernel -s 1 -c 1 \
    -o $OUT/prof_cfg5 python tools/bench_configs.py cfg5 > $OUT/prof_cfg5.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:large_atoms_kernel -s 1 -c 1 \
    -o $OUT/prof_cfg4 python tools/bench_configs.py cfg4 > $OUT/prof_cfg4.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_cfg4.csv \
    python tools/bench_configs.py cfg4 > $OUT/launches_cfg4.log 2>&1
ls -la $OUT
cat $OUT/configs.jsonl
