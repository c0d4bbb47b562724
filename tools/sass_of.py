#!/usr/bin/env python
"""Print the SASS of one kernel of a built library (instructions only, one per line).
usage: sass_of.py lib.so kernel-substring [> out.sass]"""
import re
import subprocess
import sys

lib, want = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
on = False
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        on = want in m.group(1)
        if on:
            print("//", m.group(1))
        continue
    if on:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*", line)
        if m:
            print(m.group(1), m.group(2).strip())
