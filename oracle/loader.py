"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Loads oracle/_build/libesvo_oracle.so (the CPU restatement of the reference; same C
signatures as include/esvo_b200.h under the prefix esvo_oracle_) for the generic ctypes binding in esvo_b200/capi.py.
Only tests/, __graft_entry__.smoke(), scripts/make_golden.py and bench.py's cpu_baseline / parity / --impl reference legs may
import this module; nothing inside the product package esvo_b200/ does."""
import ctypes as C
import os

ORACLE_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libesvo_oracle.so")


def load_oracle():
    from esvo_b200.capi import Library
    if not os.path.exists(ORACLE_LIB):
        raise RuntimeError(f"{ORACLE_LIB} is missing: run `make -C oracle`")
    return Library(C.CDLL(ORACLE_LIB), "esvo_oracle_", ORACLE_LIB)
