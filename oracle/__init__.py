"""CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/fock_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package (fock_rs_b200) never does.
"""
from .ref import *  # noqa: F401,F403
