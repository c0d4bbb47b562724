"""CPU-side checks of the drop-in boundary: libsasa_b200.so loads, exports every symbol include/sasa_b200.h
declares, and refuses to run without a CUDA device (there is no CPU fallback).  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sasa_b200.h")).read()
    return sorted(set(re.findall(r"SASA_B200_API\s+[\w\s\*]+?\b(sasa_b200_\w+)\s*\(", src)))


def test_header_and_binding_agree():
    from rustsasa_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 15
    assert sorted(_lib.EXPORTS) == syms, "rustsasa_b200/_lib.py EXPORTS must list exactly the header's entry points"


def test_library_exports_every_declared_symbol():
    from rustsasa_b200 import _lib
    L = _lib.load()
    for name in header_symbols():
        assert getattr(L, name) is not None, name
    assert L.sasa_b200_abi_version() == 1


def test_struct_layouts_match_header():
    """sizeof of the ctypes mirrors == what the C compiler lays out for the header's structs."""
    import subprocess
    import tempfile
    from rustsasa_b200 import _lib
    prog = ('#include <stdio.h>\n#include "sasa_b200.h"\nint main(void){printf("%zu %zu %zu\\n", '
            'sizeof(sasa_b200_params), sizeof(sasa_b200_outputs), sizeof(sasa_b200_stats));return 0;}\n')
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(src, "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, src], check=True)
        out = subprocess.run([exe], check=True, stdout=subprocess.PIPE, text=True).stdout.split()
    assert [int(x) for x in out] == [C.sizeof(_lib.Params), C.sizeof(_lib.Outputs), C.sizeof(_lib.Stats)]


def test_sphere_points_is_host_side_and_matches_oracle(oracle):
    """sasa_b200_sphere_points needs no device: the golden-spiral table is computed on the host with libm."""
    from rustsasa_b200 import Engine
    for n in (1, 100, 960):
        assert np.array_equal(Engine.sphere_points(n), oracle.sphere_points(n))


def test_no_device_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from rustsasa_b200 import Engine, SasaB200Error, _lib
    with pytest.raises(SasaB200Error) as ei:
        Engine()
    assert ei.value.code == _lib.ERR_CUDA and "no CPU fallback" in str(ei.value)


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "rustsasa_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "sasa_oracle" not in text and "liboracle" not in text, f
