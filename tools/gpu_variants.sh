#!/bin/bash
# Bench every library under rustsasa_b200/variants/ (built by tools/variants.py): device/e2e atoms/s plus the executed
# warp-instruction count and issue utilisation of one launch (ncu, two metrics).  usage (under gpurun): bash tools/gpu_variants.sh [tag]
set -u
TAG=${1:-variants}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for lib in rustsasa_b200/variants/*.so; do
    name=$(basename $lib .so); name=${name#libsasa_b200_}
    export SASA_B200_LIB=$PWD/$lib
    line=$(timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu --no-secondary 2> $OUT/$name.err | tail -1)
    echo "$line" > $OUT/$name.json
    timeout 120 ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum \
        --clock-control none -k regex:sasa_tight_kernel -s 3 -c 1 --csv --log-file $OUT/$name.ncu.csv \
        python bench.py --steps 1 --warmup 3 --no-cpu --no-secondary > /dev/null 2>&1
    python - "$name" "$OUT/$name.json" "$OUT/$name.ncu.csv" <<'PY'
import csv, json, sys
name, j, c = sys.argv[1:4]
try:
    d = json.load(open(j))
    s = "value %7.1f M atoms/s  e2e %7.1f  same=%s" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d.get("results_identical_device_vs_host_leg"))
    atoms = d["config"]["atoms"]
except Exception as e:
    s, atoms = "bench FAILED %s" % e, 0
m = {}
try:
    for r in csv.DictReader(l for l in open(c) if l.startswith('"')):
        m[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
except Exception:
    pass
inst = m.get("smsp__inst_executed.sum", 0.0)
print("%-24s %s  inst/atom %7.1f  issue %5.1f%%" % (name, s, inst / atoms if atoms else 0.0, m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0)), flush=True)
PY
done | tee $OUT/summary.txt
