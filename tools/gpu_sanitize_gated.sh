#!/bin/bash
# compute-sanitizer over the gated single-launch host pipeline (tests/gated_check.py on a small batch cut into many chunks)
set -u
OUT=gpurun_out/${1:-sanitize_gated}
mkdir -p $OUT
export SASA_B200_CHUNK_ATOMS=6000 GATED_CHECK_STRUCTURES=60 GATED_CHECK_FRAMES=16
for tool in memcheck racecheck; do
  ( timeout 280 compute-sanitizer --tool $tool --print-limit 20 python tests/gated_check.py 2>&1 | tail -30 ) > $OUT/gated_$tool.log
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ALL OK|FAILURES|FAIL |Error|hazard|one launch" $OUT/gated_$tool.log | head -12
done
