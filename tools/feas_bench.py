#!/usr/bin/env python
"""tools/feas_bench.py [log2_n] [steps] -- under gpurun: device-timed rqcd_check_input_feasibility on synthetic candidates
(c3's distribution) resident in HBM; the command the ncu capture of k_feas runs."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rapidquadcoptercollisiondetection_b200 import abi  # noqa: E402
from rapidquadcoptercollisiondetection_b200.api import Context  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 22)
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ctx = Context(0)
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
ctx.set_stream(s.cuda_stream)
d_bc = torch.empty((19, n), dtype=torch.float64, device="cuda")
ctx.synth_bc(1, 0, n, d_bc.data_ptr(), n)
d_res = torch.empty(n, dtype=torch.uint8, device="cuda")
inp = abi.BcSoa(d_bc.data_ptr(), n, abi.GOAL_ALL, 0, 0, 0)
grav = (C.c_double * 3)(0.0, 0.0, -9.81)


def f():
    assert ctx.lib.rqcd_check_input_feasibility(ctx.h, C.byref(inp), n, grav, 5.0, 30.0, 20.0, 0.002, None, abi.F_DEVICE_PTRS,
                                                C.c_void_p(d_res.data_ptr())) == 0


for _ in range(2):
    f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
for _ in range(steps):
    f()
e1.record(s)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"n {n} ms {ms:.3f} {n / ms * 1e3:.4g} candidates/s", np.bincount(d_res.cpu().numpy(), minlength=5).tolist())
