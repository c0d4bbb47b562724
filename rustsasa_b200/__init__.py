"""rustsasa_b200 -- B200 (sm_100a) Shrake-Rupley engine behind RustSASA's hot-path interface.

The product is ``libsasa_b200.so`` (hand-written CUDA + a C ABI, include/sasa_b200.h).  This package is
the thin Python host layer used by the tests and the benchmark: a ctypes binding (``_lib``), an
Engine / Batch wrapper (``engine``) and a mirror of the reference's public interface for the path
(``api``: ``calculate_sasa_internal``, ``SASAOptions`` ...).  There is no CPU fallback.
"""
from .api import (Atom, AtomLevel, ChainLevel, ChainResult, ProteinLevel, ProteinResult, ResidueLevel,  # noqa: F401
                  ResidueResult, SASACalcError, SASAOptions, calculate_sasa_internal, default_engine)
from .engine import Batch, BatchResult, Engine, SasaB200Error  # noqa: F401
from .structure import read_structure  # noqa: F401

__all__ = ["Atom", "AtomLevel", "ResidueLevel", "ChainLevel", "ProteinLevel", "ChainResult", "ResidueResult",
           "ProteinResult", "SASACalcError", "SASAOptions", "calculate_sasa_internal", "default_engine", "Engine",
           "Batch", "BatchResult", "SasaB200Error", "read_structure"]
