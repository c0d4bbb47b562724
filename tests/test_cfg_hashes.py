"""CPU check of tests/golden/cfg_hashes.json: the stored fingerprint of BASELINE cfg4 is what the oracle computes here (the
1M-atom x 960-point cfg5 entry takes 12+ s of all cores and is re-derived by tools/make_cfg_hashes.py; here its workload hash
and its consistency with the GPU-reported sum are checked)."""
import hashlib
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cfg4_fingerprint_is_the_oracles():
    from oracle import load
    from rustsasa_b200 import workloads as W
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "cfg_hashes.json")))
    a = W.large_assembly(150000)
    g = gold["cfg4"]
    assert a.n_atoms == g["atoms"]
    assert hashlib.sha256(np.ascontiguousarray(a.xyzr).tobytes()).hexdigest() == g["xyzr_sha256"]
    o = load(fast=True).calculate_sasa_internal(a.xyzr, 1.4, g["n_points"], threads=-1)
    assert int(o["counts"].astype(np.int64).sum()) == g["sum_counts"]
    assert hashlib.sha256(np.ascontiguousarray(o["counts"], dtype="<u4").tobytes()).hexdigest() == g["sha256_counts"]
    assert gold["cfg5"]["atoms"] == 999854 and gold["cfg5"]["n_points"] == 960 and gold["cfg5"]["sum_counts"] == 11061027
