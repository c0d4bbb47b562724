#!/usr/bin/env python
"""SASS instructions (address order, with executed counts) attributed to given source lines in an ncu source dump.
usage: ncu_sass.py dump.csv kernel-substring file.cuh line [line ...]"""
import csv
import sys

dump, kern, fname = sys.argv[1:4]
lines = {int(x) for x in sys.argv[4:]}
fn = hdr = cur = f = None
out = []
for r in csv.reader(open(dump)):
    if len(r) >= 2 and r[0] == "File Path":
        f = r[1].split("/")[-1]
    elif len(r) >= 2 and r[0] == "Function Name":
        fn = r[1]
    elif len(r) > 5 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and fn and kern in fn:
        if r[2] in ("-", ""):
            try:
                cur = int(r[0])
            except ValueError:
                cur = None
        elif f == fname and cur in lines:
            try:
                n = int(r[hdr.index("Instructions Executed")])
            except ValueError:
                n = 0
            out.append((r[2], cur, r[3].strip(), n))
out.sort()
for a, l, s, n in out:
    print(a[-5:], l, f"{n:>11d}", s)
