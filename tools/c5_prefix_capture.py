#!/usr/bin/env python
"""tools/c5_prefix_capture.py [log2_n=21] -- the command of the c5 ncu capture when GPU minutes are short: the first 2^log2_n
candidates of BASELINE config 5 against its 256-obstacle table, host buffers in, three calls. Same kernel variant as the full
configuration (k_pool<0,0,384,1>, one CTA of 384 threads per SM); a full-size launch takes 0.75 s and ncu replays it 40 times."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rapidquadcoptercollisiondetection_b200 import abi, workloads as W  # noqa: E402
from rapidquadcoptercollisiondetection_b200.api import Context  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 21)
specs = W.static_obstacles(6, 128, 128)  # bench.py's c5 table (obs_seed 6)
obs = abi.make_obstacles(specs)
bc = W.synth_bc(5, 0, n)                  # bench.py's c5 candidates (seed 5), global indices 0 .. n-1
with Context(0) as ctx:
    ctx.set_obstacles(obs, len(specs))
    for _ in range(3):
        r = ctx.generate_check(bc, 0.002, matrix=False)
    print(n, "candidates x", len(specs), "obstacles:", r["summary"])
