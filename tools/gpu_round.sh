#!/bin/bash
# One GPU-box visit: parity tests, the bench (both arms), the ncu launch list and one full capture of the top kernel.
# usage (under gpurun): bash tools/gpu_round.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt
( timeout 300 python -m pytest tests -m gpu -x -q --timeout=90 2>&1 | tail -15 ) > $OUT/pytest_gpu.log
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > $OUT/smoke.log
( timeout 240 python bench.py --steps 10 --warmup 3 2> $OUT/bench.err | tail -1 ) > $OUT/bench.json
( timeout 240 python bench.py --impl reference --steps 3 --warmup 1 2> $OUT/bench_ref.err | tail -1 ) > $OUT/bench_ref.json
# launch list: every kernel of 1 warm-up + 2 timed steps of the device leg (cold-cache, serialised: compare shares)
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/launches_bench.log 2>&1
# full capture of the dominant kernel on the bench workload itself (2 launches x ~40 replays)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sasa_tight_kernel -s 2 -c 2 \
    -o $OUT/prof python bench.py --steps 1 --warmup 3 --no-cpu > $OUT/prof_bench.log 2>&1
( timeout 120 python tools/latency_single.py 2>&1 | tail -8 ) > $OUT/latency_single.log
ls -la $OUT
