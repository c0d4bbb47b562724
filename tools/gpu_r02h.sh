#!/bin/bash
set -u
TAG=${1:-r02h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_host_cpp.py -x -q --timeout=300 2>&1 | tail -15 ) > $OUT/pytest_host.log
( timeout 900 python tools/cli_dir_bench.py 4400 residue --cpu 2>&1 | tail -30 ) > $OUT/cli_4400.log
( timeout 600 python tools/cli_dir_bench.py 4400 residue --tile 4400 2>&1 | tail -8 ) > $OUT/cli_4400_onetile.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches_cfg5.csv \
    python tools/bench_configs.py cfg5 > $OUT/launches_cfg5.log 2>&1
tail -5 $OUT/pytest_host.log
cat $OUT/cli_4400.log $OUT/cli_4400_onetile.log
