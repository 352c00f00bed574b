"""TEST INFRASTRUCTURE: the flat driver API on two back ends.

`Backend("gpu")` is the product (quantum-mg_b200/driver.py over libqmg_host.so).  `Backend("ref")` binds the SAME driver
API exported with the prefix ref_ by oracle/_ref/libqmg_ref.so -- the reference's unmodified headers on the CPU, the
checker.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
import os

import driver
from driver import C, CD, KCycle, Lattice, Multigrid, Stencil, Transfer, carr, np   # noqa: F401  (re-exported for the tests)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libqmg_ref.so")
GPU_LIB = driver.GPU_LIB


def have_ref():
    return os.path.exists(REF_LIB)


class Backend(driver.Backend):
    def __init__(self, kind):
        if kind == "ref":
            driver.Backend.__init__(self, "ref", REF_LIB, "ref_")
        elif kind == "gpu":
            driver.Backend.__init__(self, "gpu")
        else:
            raise ValueError(kind)
