#!/usr/bin/env python
"""tools/grazing_cmp.py OUT.npy -- verdicts of the grazing-trajectory population of tests/test_gpu_parity.py
(test_grazing_and_badly_scaled_trajectories) through the library named by RQCD_LIB, and their differences from the
oracle; used to compare kernel builds on inputs whose verdict hangs on the last bits of the cubic's transcendental core."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import binding as B
from rapidquadcoptercollisiondetection_b200 import abi
from rapidquadcoptercollisiondetection_b200.api import Context
from tests.util import grazing_case

coeffs, t_end, specs = grazing_case()
obs = abi.make_obstacles(specs)
o = B.Oracle().check(coeffs, t_end, obs, len(specs), 0.002, nthreads=os.cpu_count() or 1)
with Context(0) as ctx:
    ctx.set_obstacles(obs, len(specs))
    r = ctx.check(coeffs, t_end, 0.002)
bad = np.argwhere(r["verdict_matrix"] != o["verdict_matrix"])
print(os.environ.get("RQCD_LIB", "default"), "pairs", o["verdict_matrix"].size, "differ from the oracle", len(bad))
np.save(sys.argv[1], bad)
