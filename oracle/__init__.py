"""CPU oracle for the Shrake-Rupley hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package (see sasa_oracle.c).  The product never does.
"""
from .oracle import Oracle, load  # noqa: F401
