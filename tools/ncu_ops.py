#!/usr/bin/env python
"""Opcode mix (share of executed warp instructions per SASS opcode) from an ncu cuda,sass source dump."""
import csv
import re
import sys
from collections import defaultdict

dump, want = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = list(csv.reader(open(dump)))
cur_fn = hdr = None
ops = defaultdict(int)
tot = 0
for r in rows:
    if len(r) >= 2 and r[0] == "Function Name":
        cur_fn = r[1]
    elif len(r) > 5 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and cur_fn and want in cur_fn and r[2] not in ("-", ""):
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[3].strip())
        if not m:
            continue
        try:
            n = int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        ops[m.group(2).split(".")[0]] += n
        tot += n
print("total warp instructions", tot)
for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:36]:
    print(f"{100 * v / tot:6.2f}%  {k}")
