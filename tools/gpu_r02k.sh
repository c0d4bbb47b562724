#!/bin/bash
# ncu --set full captures of this round's kernels on their own workloads + launch lists.  The reports are condensed to text
# on the box (tools/ncu_summary.py, tools/ncu_funcs.py); only the headline and cfg5 reports travel back (64 MiB limit).
set -u
TAG=${1:-r02k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
cap() {  # name kernel-regex skip command...
    local name=$1 kern=$2 skip=$3; shift 3
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c 1 -o $OUT/prof_$name "$@" > $OUT/prof_$name.log 2>&1
    python tools/ncu_summary.py $OUT/prof_$name.ncu-rep > $OUT/${name}_summary.txt 2>/dev/null
    ncu -i $OUT/prof_$name.ncu-rep --page source --csv --print-source cuda,sass > $OUT/src_$name.csv 2>/dev/null
    python tools/ncu_funcs.py $OUT/src_$name.csv $kern > $OUT/${name}_functions.txt 2>/dev/null
    rm -f $OUT/src_$name.csv
}
cap cfg3 sasa_tight_kernel 2 python tools/bench_configs.py cfg3 --frames 1500
cap cfg5 large_cells_kernel 1 python tools/bench_configs.py cfg5
cap cfg4 large_cells_kernel 1 python tools/bench_configs.py cfg4
cap cfg2 sasa_tight_kernel 3 python bench.py --steps 1 --warmup 3 --no-cpu --no-secondary
rm -f $OUT/prof_cfg3.ncu-rep $OUT/prof_cfg4.ncu-rep
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > $OUT/launches_bench.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches_cfg4.csv \
    python tools/bench_configs.py cfg4 > $OUT/launches_cfg4.log 2>&1
( timeout 120 python tools/latency_single.py 2>&1 | tail -8 ) > $OUT/latency_single.log
du -sh $OUT; ls -la $OUT
