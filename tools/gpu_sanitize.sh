#!/bin/bash
set -u
TAG=${1:-sanitize}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export SASA_B200_CHUNK_ATOMS=4000   # several chunks even for the small batches of sanitize_paths.py: pipelined and gated host paths
python tools/sanitize_paths.py > $OUT/plain.log 2>&1; echo "plain rc=$?" >> $OUT/plain.log
for tool in memcheck racecheck initcheck; do
  ( timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_paths.py 2>&1 | tail -40 ) > $OUT/$tool.log
done
tail -3 $OUT/plain.log
for tool in memcheck racecheck initcheck; do echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ALL OK|FAILURES|Error|hazard" $OUT/$tool.log | head -12; done
