#!/usr/bin/env python
"""tools/fragile_dump.py -- under gpurun: the GPU's verdict matrices and RQCD_F_FRAGILE marks of the grazing populations,
saved for tools/fragile_cpu.py (which explains the differences from the oracle on the CPU)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rapidquadcoptercollisiondetection_b200 import abi  # noqa: E402
from rapidquadcoptercollisiondetection_b200.api import Context  # noqa: E402
from tests.util import grazing_case  # noqa: E402

out = {}
for log_a in ((1.0, 4.0), (3.0, 6.0), (1.0, 8.0), (6.0, 8.0)):
    coeffs, t_end, specs = grazing_case(log_a=log_a)
    obs = abi.make_obstacles(specs)
    tag = f"{log_a[0]:g}_{log_a[1]:g}"
    with Context(0) as c:
        c.set_obstacles(obs, len(specs))
        r = c.check(coeffs, t_end, 0.002, fragile=True)
    out[f"verdict_{tag}"] = r["verdict_matrix"]
    out[f"marked_{tag}"] = r["fragile_matrix"]
    print(tag, int((r["fragile_matrix"] != 0).sum()), flush=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "fragile_dump.npz"), **out)
