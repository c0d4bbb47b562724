#!/usr/bin/env python
"""Build tuning variants of libsasa_b200.so (one per set of -D flags, default configurations only) into
rustsasa_b200/variants/, for tools/gpu_variants.sh to bench on the GPU box.

usage: variants.py name1:"-DFOO=1 -DBAR" name2:"" ...      (builds run in parallel)
"""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rustsasa_b200 import build as B  # noqa: E402

VDIR = os.path.join(ROOT, "rustsasa_b200", "variants")


def one(spec):
    name, _, flags = spec.partition(":")
    out = os.path.join(VDIR, f"libsasa_b200_{name}.so")
    B.build(force=True, defines=tuple(flags.split()), out=out)
    return out


def main():
    os.makedirs(VDIR, exist_ok=True)
    for f in os.listdir(VDIR):
        os.remove(os.path.join(VDIR, f))
    with ThreadPoolExecutor(max_workers=4) as ex:
        for o in ex.map(one, sys.argv[1:]):
            print("built", o, flush=True)


if __name__ == "__main__":
    main()
