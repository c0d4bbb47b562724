#!/usr/bin/env python
"""Condense an ncu report into the text summary committed under profiles/.

usage: ncu_summary.py report.ncu-rep [launches.csv] > profiles/<name>.txt
Reads the report with `ncu -i ... --page raw --csv` (works without a GPU)."""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

EXACT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio$")


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def to_bytes(value, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(value.replace(",", "")) * scale.get(unit, 1.0)


def traffic(rep, structures, out, per_step=0):
    """--traffic: profiles/ncu_traffic.json for bench.py (DRAM bytes of ONE step = sum over its launches; per_step > 0:
    the capture holds several steps of per_step launches each, keep the first)."""
    import json
    hdr, units, rows = raw(rep)
    if per_step:
        rows = rows[:per_step]
    col = {h: i for i, h in enumerate(hdr)}
    total, per = 0.0, []
    for r in rows:
        b = sum(to_bytes(r[col[k]], units[col[k]]) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        per.append({"kernel": r[col["Kernel Name"]], "grid": r[col["Grid Size"]], "dram_bytes": b})
        total += b
    json.dump({"structures": structures, "dram_bytes_per_step": total, "launches": per,
               "source": f"ncu --set full --clock-control none, {rep.split('/')[-2]}/{rep.split('/')[-1]}"}, open(out, "w"), indent=1)
    print(out, total)


def main():
    if sys.argv[1] == "--traffic":
        return traffic(sys.argv[2], int(sys.argv[3]), sys.argv[4], int(sys.argv[5]) if len(sys.argv) > 5 else 0)
    rep = sys.argv[1]
    hdr, units, rows = raw(rep)
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full --clock-control none: {rep.split('/')[-1]}  (per-launch values; cold-cache, serialised)")
    for r in rows:
        print(f"\n## {r[col['Kernel Name']]}  grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
        for name in EXACT:
            if name in col and r[col[name]] != "":
                print(f"  {name:72s} {r[col[name]]:>16s} {units[col[name]]}")
        stalls = [(float(r[i] or 0), STALL.search(h).group(1)) for h, i in col.items() if STALL.search(h)]
        tot = sum(v for v, _ in stalls) or 1.0
        print("  warp stall reasons (cycles per issued instruction; share):")
        for v, n in sorted(stalls, reverse=True)[:8]:
            print(f"    {n:28s} {v:7.3f}  {100 * v / tot:5.1f}%")
    if len(sys.argv) > 2:
        print(f"\n# launch list: {sys.argv[2].split('/')[-1]} (ncu --metrics gpu__time_duration.sum --clock-control none)")
        agg, n = defaultdict(float), defaultdict(int)
        lines = [ln for ln in open(sys.argv[2]) if ln.startswith('"')]
        for d in csv.DictReader(lines):
            if d.get("Metric Name") != "gpu__time_duration.sum":
                continue
            k = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "")
            agg[k] += float(d["Metric Value"].replace(",", "")) * (1e-3 if d["Metric Unit"] in ("ns", "nsecond") else 1.0)
            n[k] += 1
        tot = sum(agg.values()) or 1.0
        print(f"  {'kernel':72s} {'launches':>8s} {'total us':>12s} {'share':>7s}")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
            print(f"  {k[:72]:72s} {n[k]:8d} {v:12.1f} {100 * v / tot:6.1f}%")


if __name__ == "__main__":
    main()
