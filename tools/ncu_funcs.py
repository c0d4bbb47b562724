#!/usr/bin/env python
"""Group an `ncu --page source --csv --print-source cuda,sass` dump by device function (line ranges parsed
from the .cuh sources) -- share of executed warp instructions and of stall samples per function."""
import csv
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def func_ranges(path):
    starts = []
    for i, line in enumerate(open(path), 1):
        m = re.match(r"^(?:__device__|__global__|inline|static).*?\b(\w+)\s*\(", line)
        if m and not line.startswith(" "):
            starts.append((i, m.group(1)))
    return starts


def main():
    dump, want = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    ranges = {}
    for f in ("sasa_device.cuh", "sasa_small.cuh", "sasa_tight.cuh", "sasa_large.cuh", "sasa_cap.cuh"):
        ranges[f] = func_ranges(os.path.join(ROOT, "rustsasa_b200", "csrc", f))
    rows = list(csv.reader(open(dump)))
    cur_file = cur_fn = hdr = None
    inst, smp = defaultdict(int), defaultdict(int)
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif len(r) >= 2 and r[0] == "Function Name":
            cur_fn = r[1]
        elif len(r) > 5 and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and cur_fn and want in cur_fn and r[2] in ("-", ""):
            try:
                line = int(r[0])
            except ValueError:
                continue
            name = cur_file
            if cur_file in ranges:
                name = cur_file + ":?"
                for s, fn in ranges[cur_file]:
                    if s <= line:
                        name = fn
            inst[name] += int(r[hdr.index("Instructions Executed")] or 0)
            smp[name] += int(r[hdr.index("# Samples")] or 0)
    ti, ts = sum(inst.values()) or 1, sum(smp.values()) or 1
    for k, v in sorted(inst.items(), key=lambda kv: -kv[1]):
        print(f"{100.0 * v / ti:6.2f}% inst {100.0 * smp[k] / ts:6.2f}% smp  {k}")


if __name__ == "__main__":
    main()
