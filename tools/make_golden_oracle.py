#!/usr/bin/env python
"""Write tests/golden/oracle_c3_tiny.npz: theta/tau/tau_b after one stress step of the C3 configuration
shrunk to h = 1 (1,152 cells), computed by the CPU oracle.  A regression pin for the checker."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import Setup, tight  # noqa: E402
from rheotool_b200 import abi, cases  # noqa: E402

spec = cases.by_name("C3", 1 / 19)
s = Setup(spec)
oc = s.oracle(tight(spec.schemes))
oc.store_old_time(); oc.step(s.dt)
np.savez_compressed(ROOT / "tests" / "golden" / "oracle_c3_tiny.npz", n_cells=s.mesh.n_cells, dt=s.dt,
                    theta=oc.get(0, 0, abi.FIELD_THETA), tau=oc.get(0, 0, abi.FIELD_TAU), tau_b=oc.get(0, 0, abi.FIELD_TAU_B))
print("wrote", s.mesh.n_cells, "cells")
