#!/bin/bash
# Like gpu_variants.sh but with bench.py's secondary block (cfg3 / cfg4 / cfg5) and without the ncu pass.
# usage (under gpurun): bash tools/gpu_variants2.sh [tag]
set -u
TAG=${1:-variants2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for lib in rustsasa_b200/variants/*.so; do
    name=$(basename $lib .so); name=${name#libsasa_b200_}
    export SASA_B200_LIB=$PWD/$lib
    line=$(timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu 2> $OUT/$name.err | tail -1)
    echo "$line" > $OUT/$name.json
    python - "$name" "$OUT/$name.json" <<'PY'
import json, sys
name, j = sys.argv[1:3]
try:
    d = json.load(open(j))
    s = d.get("secondary", {})
    f = lambda k: "%s %.4g ms %s" % (k, s[k]["ms_per_step"], "ok" if s[k].get("parity") else "PARITY?") if k in s else k + " -"
    print("%-14s cfg2 %7.1f M atoms/s e2e %7.1f | %s | %s | %s" % (name, d["value"] / 1e6, d["e2e"]["value"] / 1e6, f("cfg3_md_frames"), f("cfg4_assembly"), f("cfg5_capsid")), flush=True)
except Exception as e:
    print(name, "FAILED", e, flush=True)
PY
done | tee $OUT/summary.txt
