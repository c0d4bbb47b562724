#!/bin/bash
# Parity tests + smoke + one bench line.  usage (under gpurun): bash tools/gpu_check.sh [tag]
set -u
TAG=${1:-check}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 400 python -m pytest tests -m gpu -x -q --timeout=120 2>&1 | tail -15 ) > $OUT/pytest_gpu.log
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > $OUT/smoke.log
( timeout 300 python bench.py --steps 10 --warmup 3 2> $OUT/bench.err | tail -1 ) > $OUT/bench.json
tail -3 $OUT/pytest_gpu.log; cat $OUT/smoke.log; cut -c1-600 $OUT/bench.json
