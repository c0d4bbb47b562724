#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.

usage: ncu_lines.py dump.csv [kernel-substring]   -> per (file, line): executed warp instructions, samples
"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = list(csv.reader(open(path)))
    i = 0
    per_kernel = {}
    cur_file, cur_fn, hdr = None, None, None
    while i < len(rows):
        r = rows[i]
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1]
        elif len(r) >= 2 and r[0] == "Function Name":
            cur_fn = r[1]
        elif len(r) > 5 and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and cur_fn and want in cur_fn:
            d = per_kernel.setdefault(cur_fn, defaultdict(lambda: [0, 0, ""]))
            try:
                line = int(r[0])
            except ValueError:
                i += 1
                continue
            key = (cur_file.split("/")[-1], line)
            # rows with an address are SASS rows belonging to the line; line rows carry the source text
            if r[2] in ("", "-"):
                d[key][2] = r[1]
                ie = hdr.index("Instructions Executed")
                ns = hdr.index("# Samples")
                try:
                    d[key][0] += int(r[ie] or 0)
                    d[key][1] += int(r[ns] or 0)
                except ValueError:
                    pass
        i += 1
    for fn, d in per_kernel.items():
        tot = sum(v[0] for v in d.values()) or 1
        tots = sum(v[1] for v in d.values()) or 1
        print(f"== {fn}: {tot} warp instructions, {tots} samples")
        for (f, line), v in sorted(d.items(), key=lambda kv: -kv[1][0])[:45]:
            print(f"{100.0 * v[0] / tot:6.2f}% inst {100.0 * v[1] / tots:6.2f}% smp  {f}:{line:<4d} {v[2].strip()[:110]}")


if __name__ == "__main__":
    main()
