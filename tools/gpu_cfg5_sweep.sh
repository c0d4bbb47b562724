#!/bin/bash
# cfg5 (1M atoms x 960 points, one GPU): chunked-cap-table grid size (SASA_B200_CAPM_N) x build variants in rustsasa_b200/variants/
set -u
TAG=${1:-cfg5sweep}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for lib in rustsasa_b200/variants/*.so; do
  name=$(basename $lib .so); name=${name#libsasa_b200_}
  for n in ${CAPM_NS:-32 64 96 128}; do
    line=$(SASA_B200_LIB=$PWD/$lib SASA_B200_CAPM_N=$n timeout 200 python tools/bench_configs.py cfg5 2>/dev/null | tail -1)
    echo "$name N=$n $line" | python -c "
import sys,json
l=sys.stdin.read().strip(); h,_,j=l.partition(' {')
try:
    d=json.loads('{'+j); print(h, 'device_ms %.3f  M atoms/s %.0f  sum %d'%(d['device_ms'], d['device_atoms_per_s']/1e6, d['sum_counts']))
except Exception as e: print(h,'FAILED',l[:200])
"
  done
done | tee $OUT/summary.txt
