#!/usr/bin/env python
"""tools/fragile_study.py -- under gpurun: for the grazing populations of tests/util.py, which pairs differ from the oracle and
which pairs RQCD_F_FRAGILE marks at each single probe magnitude 2^k; saved to gpurun_out/fragile_study.npz for offline
choice of the default magnitudes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import binding as B  # noqa: E402
from rapidquadcoptercollisiondetection_b200 import abi  # noqa: E402
from rapidquadcoptercollisiondetection_b200.api import Context  # noqa: E402
from tests.util import grazing_case  # noqa: E402

MAGS = [-50]
MARGINS = [(4, 8)]
out = {}
orc = B.Oracle()
for log_a in ((1.0, 4.0), (3.0, 6.0), (1.0, 8.0), (6.0, 8.0)):
    coeffs, t_end, specs = grazing_case(log_a=log_a)
    obs = abi.make_obstacles(specs)
    o = orc.check(coeffs, t_end, obs, len(specs), 0.002, nthreads=8)
    tag = f"{log_a[0]:g}_{log_a[1]:g}"
    for k in MAGS:
      for mg in MARGINS:
        os.environ["RQCD_FRAGILE_LOG2"] = str(k)
        os.environ["RQCD_FRAGILE_MARGIN"] = str(mg[0])
        os.environ["RQCD_FRAGILE_FLOOR"] = str(mg[1])
        with Context(0) as c:
            c.set_obstacles(obs, len(specs))
            r = c.check(coeffs, t_end, 0.002, fragile=True)
            print("   reasons", {int(k_): int(v_) for k_, v_ in enumerate(c.fragile_reasons()) if v_})
        differs = r["verdict_matrix"] != o["verdict_matrix"]
        marked = r["fragile_matrix"] != 0
        out[f"marked_{tag}_{k}_{mg}"] = np.packbits(marked)
        if k == MAGS[0]:
            out[f"differs_{tag}"] = np.packbits(differs)
        print(tag, k, mg, int(differs.sum()), int(marked.sum()), int((differs & ~marked).sum()), flush=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "fragile_study.npz"), **out)
# how specific is the mark on a Monte Carlo population (c3's distribution, 2^16 x 64)?
from rapidquadcoptercollisiondetection_b200 import workloads as W  # noqa: E402
specs = W.static_obstacles(2, 32, 32)
obs = abi.make_obstacles(specs)
bc = W.synth_bc(1, 0, 1 << 16)
o = orc.generate_check(bc, obs, len(specs), 0.002, nthreads=8)
for k, mg in (("-50", (4, 8)),):
    os.environ["RQCD_FRAGILE_LOG2"] = k
    os.environ["RQCD_FRAGILE_MARGIN"] = str(mg[0])
    os.environ["RQCD_FRAGILE_FLOOR"] = str(mg[1])
    with Context(0) as c:
        c.set_obstacles(obs, len(specs))
        r = c.generate_check(bc, 0.002, fragile=True)
        print("   reasons", {int(k_): int(v_) for k_, v_ in enumerate(c.fragile_reasons()) if v_})
    differs = r["verdict_matrix"] != o["verdict_matrix"]
    marked = r["fragile_matrix"] != 0
    print("c3-sample", k, mg, int(differs.sum()), int(marked.sum()), marked.size, int(r["fragile"].sum()), flush=True)

# the Forest workload (rest-to-rest primitives: an exact double root of every velocity polynomial at t = Tf) with early exit
fspecs = W.forest_obstacles()
fobs = abi.make_obstacles(fspecs)
c2 = W.c2_candidates(1000)
of = orc.generate_check(c2, fobs, 5, 0.002, flags=abi.F_EARLY_EXIT, matrix=False, nthreads=8)
oa = orc.generate_check(c2, fobs, 5, 0.002, nthreads=8)
os.environ["RQCD_FRAGILE_LOG2"], os.environ["RQCD_FRAGILE_MARGIN"], os.environ["RQCD_FRAGILE_FLOOR"] = "-50", "4", "8"
with Context(0) as c:
    c.set_obstacles(fobs, 5)
    r = c.generate_check(c2, 0.002, flags=abi.F_EARLY_EXIT, fragile=True)
    print("forest early exit: verdict diffs", int((r["verdict"] != of["verdict"]).sum()), "trajectories marked", int(r["fragile"].sum()), "of", c2.shape[1],
          "reasons", {int(k_): int(v_) for k_, v_ in enumerate(c.fragile_reasons()) if v_}, flush=True)
with Context(0) as c:
    c.set_obstacles(fobs, 5)
    r = c.generate_check(c2, 0.002, fragile=True)
    print("forest all pairs: pair diffs", int((r["verdict_matrix"] != oa["verdict_matrix"]).sum()), "pairs marked", int((r["fragile_matrix"] != 0).sum()), "of", r["fragile_matrix"].size,
          "reasons", {int(k_): int(v_) for k_, v_ in enumerate(c.fragile_reasons()) if v_}, flush=True)
