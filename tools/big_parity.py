#!/usr/bin/env python
"""tools/big_parity.py [log2_n=21] [m=256] [static|moving|forest] -- one-off large verdict-parity run: n candidates x m
obstacles through the CUDA path (C ABI) and through the CPU oracle on all host threads; prints the number of pairs and
of differing verdicts. static: the C5 table; moving: C4-style translating obstacles; forest: the five Forest prisms
with the driver's early exit."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import binding as B
from rapidquadcoptercollisiondetection_b200 import abi, workloads as W
from rapidquadcoptercollisiondetection_b200.api import Context

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 21
m = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n = 1 << log2n
kind = sys.argv[3] if len(sys.argv) > 3 else "static"
orc = B.Oracle()
flags = 0
if kind == "moving":
    specs = W.moving_obstacles(4, m, lambda bc: orc.generate(bc)["coeffs"])
elif kind == "forest":
    specs = W.forest_obstacles()
    m = len(specs)
    flags = abi.F_EARLY_EXIT
else:
    specs = W.static_obstacles(6, m // 2, m - m // 2)
obs = abi.make_obstacles(specs)
total = diff = 0
chunk = 1 << 18
with Context(0) as ctx:
    ctx.set_obstacles(obs, m)
    for base in range(0, n, chunk):
        bc = W.synth_bc(5, (1 << 24) + base, chunk)
        t0 = time.time()
        g = ctx.generate_check(bc, 0.002, flags=flags)
        t1 = time.time()
        o = orc.generate_check(bc, obs, m, 0.002, flags=flags, nthreads=os.cpu_count() or 1)
        t2 = time.time()
        d = int((g["verdict"] != o["verdict"]).sum()) if flags else int((g["verdict_matrix"] != o["verdict_matrix"]).sum())
        d += int((g["first_obstacle"] != o["first_obstacle"]).sum())
        total += chunk * m
        diff += d
        print(f"chunk {base >> 18}: {chunk * m} pairs, {d} differences (gpu call {t1 - t0:.2f} s, oracle {t2 - t1:.2f} s)", flush=True)
print(f"TOTAL {total} pairs, {diff} differing verdicts / first-hit indices")
