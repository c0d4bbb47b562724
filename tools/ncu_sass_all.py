#!/usr/bin/env python
"""Whole-kernel SASS listing in address order from an `ncu --page source --csv --print-source cuda,sass` dump:
address, share of executed warp instructions, stall samples, source line, instruction.  usage: ncu_sass_all.py dump.csv"""
import csv, sys
dump = sys.argv[1]
fn = hdr = cur = f = None
out = {}
for r in csv.reader(open(dump)):
    if len(r) >= 2 and r[0] == "File Path":
        f = r[1].split("/")[-1]
    elif len(r) >= 2 and r[0] == "Function Name":
        fn = r[1]
    elif len(r) > 5 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and fn:
        if r[2] in ("-", ""):
            try: cur = int(r[0])
            except ValueError: cur = None
        else:
            try: n = int(r[hdr.index("Instructions Executed")])
            except ValueError: n = 0
            try: s = int(r[hdr.index("# Samples")])
            except ValueError: s = 0
            key = r[2]
            if not key.startswith("0x") and not all(c in "0123456789abcdef" for c in key.lower()): continue
            if key in out:
                out[key][3] += n; out[key][4] += s
                if (f, cur) not in out[key][0]: out[key][0].append((f, cur))
            else:
                out[key] = [[(f, cur)], key, r[3].strip(), n, s]
tot = sum(v[3] for v in out.values())
for k in sorted(out, key=lambda a: int(a, 16)):
    locs, a, s, n, smp = out[k]
    loc = ",".join(f"{x[0].replace('sasa_','').replace('.cuh','').replace('.hpp','')}:{x[1]}" for x in locs[-1:])
    print(f"{a[-5:]} {100.0*n/tot:6.3f}% {smp:7d} {loc:28s} {s}")
