#!/usr/bin/env python
"""tools/fragile_cpu.py -- CPU only: explain the GPU/oracle verdict differences of the grazing populations
(gpurun_out/fragile_dump.npz from tools/fragile_dump.py) with a privately built oracle whose libm outputs can be moved
(tools/oracle_perturbed)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import binding as B  # noqa: E402
from rapidquadcoptercollisiondetection_b200 import abi  # noqa: E402
from tests.util import grazing_case  # noqa: E402


class Pert(B.Oracle):
    def __init__(self):
        B._CpuImpl.__init__(self, os.path.join(ROOT, "tools", "oracle_perturbed", "liboracle_pert.so"))
        self.lib.orc_set_perturb.argtypes = [C.c_double, C.c_double, C.c_int]

    def set(self, a=1.0, d=0.0, mode=0):
        self.lib.orc_set_perturb(a, d, mode)


z = np.load(os.path.join(ROOT, "gpurun_out", "fragile_dump.npz"))
orc = Pert()
for log_a in ((1.0, 4.0), (3.0, 6.0), (1.0, 8.0), (6.0, 8.0)):
    tag = f"{log_a[0]:g}_{log_a[1]:g}"
    coeffs, t_end, specs = grazing_case(log_a=log_a)
    obs = abi.make_obstacles(specs)
    gpu, marked = z[f"verdict_{tag}"], z[f"marked_{tag}"] != 0
    orc.set()
    o = orc.check(coeffs, t_end, obs, len(specs), 0.002, nthreads=8)["verdict_matrix"]
    differs = gpu != o
    un = differs & ~marked
    print(tag, "differs", int(differs.sum()), "marked", int(marked.sum()), "unmarked differs", int(un.sum()))
    orc.set(mode=1)
    oc = orc.check(coeffs, t_end, obs, len(specs), 0.002, nthreads=8)["verdict_matrix"]
    print("   oracle with cbrt: differs from gpu", int((gpu != oc).sum()), "of which unmarked", int(((gpu != oc) & ~marked).sum()),
          "; differs from plain oracle", int((oc != o).sum()))
    # is every difference reproduced by SOME small relative move of pow's result?
    explained = np.zeros_like(differs)
    for k in range(-53, -39):
        for sgn in (+1, -1):
            for mode in (0, 1):
                orc.set(a=1.0 + sgn * 2.0 ** k, mode=mode)
                op = orc.check(coeffs, t_end, obs, len(specs), 0.002, nthreads=8)["verdict_matrix"]
                explained |= differs & (op == gpu)
    print("   explained by a relative move of pow within 2^-40:", int(explained.sum()), "; unexplained:", int((differs & ~explained).sum()),
          "; unexplained and unmarked:", int((un & ~explained).sum()))
    bad = np.argwhere(un)
    print("   first unmarked:", bad[:5].tolist(), "gpu", gpu[un][:5].tolist(), "oracle", o[un][:5].tolist(), "oracle-cbrt", oc[un][:5].tolist())
