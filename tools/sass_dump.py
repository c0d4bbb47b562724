#!/usr/bin/env python
"""Full SASS listings (cuobjdump -sass) of the kernels the design rests on, committed under profiles/ as
<tag>_sass_<kernel>.txt.gz, plus one plain-text summary per kernel: instruction count, mnemonic histogram and the
memory / warp-collective instructions that show how the kernel touches shared memory, L2 and the table
(LDS / STS / LDG.E.*.CONSTANT / REDUX / VOTE / SHFL / ATOMS).   usage: python tools/sass_dump.py [tag]"""
import collections
import gzip
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rustsasa_b200", "libsasa_b200.so")
WANT = {
    "tight_1024": "_ZN4sasa17sasa_tight_kernelILi1024ELi1ELb0ELj16384ELi3EEEvNS_7KParamsE",
    "small_1024": "_ZN4sasa17sasa_small_kernelILi1024ELi1ELb0ELj16384EEEvNS_7KParamsE",
    "large_cells_1": "_ZN4sasa18large_cells_kernelILi1ELb0EEE",
    "large_cells_1_tex": "_ZN4sasa18large_cells_kernelILi1ELb1EEE",
    "large_cells_8": "_ZN4sasa18large_cells_kernelILi8ELb0EEE",
    "large_atoms": "_ZN4sasa18large_atoms_kernelE",
    "large_scan": "_ZN4sasa17large_scan_kernelE",
}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    text = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    parts = re.split(r"(?m)^\s*Function : ", text)
    summary = [f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a); full listings: profiles/{tag}_sass_<kernel>.txt.gz\n"]
    for short, mangled in WANT.items():
        body = next((p for p in parts[1:] if p.startswith(mangled)), None)
        if body is None:
            summary.append(f"## {short}: not found\n")
            continue
        with gzip.open(os.path.join(ROOT, "profiles", f"{tag}_sass_{short}.txt.gz"), "wt") as fh:
            fh.write("Function : " + body)
        ops = collections.Counter()
        n = 0
        for line in body.splitlines():
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                n += 1
                ops[m.group(1)] += 1
        base = collections.Counter()
        for k, v in ops.items():
            base[k.split(".")[0]] += v
        summary.append(f"## {short}  ({mangled[:70]}...)  {n} instructions\n")
        summary.append("  by opcode: " + ", ".join(f"{k} {v}" for k, v in base.most_common(28)) + "\n")
        keys = [k for k in ops if re.match(r"(LDS|STS|LDG|STG|LDC|TLD|TEX|REDUX|CREDUX|VOTE|SHFL|ATOMS|ATOMG|RED|MUFU|BAR|LDGSTS|UTMA|MATCH|POPC|FFMA|FSETP)", k)]
        summary.append("  memory / collective / fp forms: " + ", ".join(f"{k} {ops[k]}" for k in sorted(keys, key=lambda k: -ops[k])[:40]) + "\n\n")
    with open(os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt"), "w") as fh:
        fh.writelines(summary)
    print("".join(summary)[:3000])


if __name__ == "__main__":
    main()
