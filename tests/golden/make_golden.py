"""Regenerates the fixtures in tests/golden/ (run in the build container, where /root/reference exists).

  python tests/golden/make_golden.py

Sources of truth:
  * the reference's own checked-in outputs: data/ForestPerfEval.csv, data/ForestPerfEvaltreeParamsCSV.csv,
    data/PerformanceEval.txt:23-25, data/ForestPerfEval.txt:28-29  -> forest_golden.npz, golden_counts.json
  * the UNMODIFIED reference classes compiled into oracle/_ref/libref.so (oracle/Makefile, ref_shim.cpp),
    run here on seeded inputs -> ref_vectors.npz

Nothing from the reference is copied: the fixtures hold inputs we drew and the numbers the reference printed
or returned for them.
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import binding as B  # noqa: E402
from rapidquadcoptercollisiondetection_b200 import abi, workloads as W  # noqa: E402

REF = os.environ.get("RQCD_REFERENCE", "/root/reference")


def adversarial_quartics():
    """Monic quartic coefficient sets that exercise every branch of src/quartic.cpp:77-156."""
    rows = [
        (0, 0, 0, 1),          # t^4 + 1: no real roots, NaN path of the D==0 sub-branch
        (0, 0, 0, 0),          # quadruple root at 0
        (0, -2, 0, 1),         # (t^2-1)^2: two double roots
        (-10, 35, -50, 24),    # roots 1,2,3,4
        (0, -5, 0, 4),         # +-1, +-2
        (-4, 6, -4, 1),        # (t-1)^4
        (0, 0, 0, -1),         # t^4 - 1: two real roots
        (1e-8, 1, 1e-8, 1),    # tiny odd coefficients
        (1e6, 1, 1, 1),        # huge a
        (0, 1e-13, 0, 0),      # |D| just under/over eps
        (2, 3, 2, 1),          # (t^2+t+1)^2 no real roots, repeated
        (0, 2, 0, 1),          # (t^2+1)^2
        (-2, -1, 2, 0),        # root at 0
        (3, 3, 1, 0),          # t(t+1)^3
        (np.inf, 1, 1, 1), (np.nan, 0, 0, 0), (0, 0, 0, np.inf),
    ]
    return np.array(rows, dtype=np.float64).T.copy()


def adversarial_cubics():
    rows = [
        (0, 0, 0), (-6, 11, -6), (0, 0, 1), (0, 0, -1), (-3, 3, -1), (0, -3, 2), (0, -3, -2), (1e6, 1, 1),
        (0, 1, 0), (0, -1, 0), (3, 3, 1 + 1e-13), (np.nan, 1, 1), (np.inf, 0, 0), (0, -1e-20, 0),
    ]
    return np.array(rows, dtype=np.float64).T.copy()


def main():
    ref = B.Reference()
    out = {}

    # ---- the reference's checked-in files -------------------------------------------------------------
    csv = np.genfromtxt(os.path.join(REF, "data", "ForestPerfEval.csv"), delimiter=",")[:, :21]
    trees = np.genfromtxt(os.path.join(REF, "data", "ForestPerfEvaltreeParamsCSV.csv"), delimiter=",")[:, :15]
    np.savez_compressed(os.path.join(HERE, "forest_golden.npz"), csv=csv, trees=trees)

    def pct(path, label):
        txt = open(os.path.join(REF, "data", path)).read()
        return float(re.search(re.escape(label) + r"\s*=\s*([0-9.eE+-]+)", txt).group(1))

    counts = {
        "c1_n": 10_000_000,
        "c1_no_collision": round(pct("PerformanceEval.txt", "Percent of trajectories that were collision-free ") * 1e5),
        "c1_collision": round(pct("PerformanceEval.txt", "Percent of trajectories with collision ") * 1e5),
        "c1_indeterminable": round(pct("PerformanceEval.txt", "Percent of trajectories with collision indeterminable ") * 1e5),
        "c2_n": 1_000_000,
        "c2_input_feasible": round(pct("ForestPerfEval.txt", "Percent of trajectories that were input feasible") * 1e4),
        "c2_collision_free": round(pct("ForestPerfEval.txt", "Percent of trajectories that were collision free") * 1e4),
        "source": "data/PerformanceEval.txt:23-25, data/ForestPerfEval.txt:28-29 (percentages x N)",
    }
    json.dump(counts, open(os.path.join(HERE, "golden_counts.json"), "w"), indent=1)

    # ---- libref on the drivers' own input streams -----------------------------------------------------
    bc, sph, verdict, trials = ref.c1_run(4096, True)
    out.update(c1_bc=bc, c1_spheres=sph, c1_verdict=verdict, c1_trials=np.int64(trials))
    bc, feas, verdict, pairs = ref.c2_run(30, 100)
    out.update(c2_bc=bc, c2_input_feas=feas, c2_verdict=verdict, c2_pairs=np.int64(pairs))

    # ---- static mixed set, all pairs --------------------------------------------------------------------
    specs = W.static_obstacles(21, 12, 12)
    specs.append({"type": "prism", "center": [0.5, -0.25, 1.0], "side": [1.0, 2.0, 0.5],
                  "quat": [0.9, 0.3, -0.2, 0.4]})  # deliberately not normalised
    obs = B.make_obstacles(specs)
    bc = W.synth_bc(11, 0, 1024)
    r = ref.generate_check(bc, obs, len(specs), 0.002)
    for k, v in W.specs_to_arrays(specs).items():
        out["st_obs_" + k] = v
    out.update(st_seed=np.int64(11), st_n=np.int64(1024), st_matrix=r["verdict_matrix"], st_verdict=r["verdict"],
               st_first=r["first_obstacle"], st_cost=r["cost"], st_coeffs=r["coeffs"],
               st_best_cost=np.float64(r["summary"]["best_cost"]), st_best_index=np.int64(r["summary"]["best_index"]))

    # ---- every goal-flag combination (SingleAxisTrajectory.cpp:71-116) + a mixed per-axis mask ---------
    masks = [0x1FF]
    for combo in range(8):
        masks.append(sum(((combo >> c) & 1) << (3 * ax + c) for ax in range(3) for c in range(3)))
    masks.append(abi.goal_pos(0) | abi.goal_vel(0) | abi.goal_acc(0) | abi.goal_pos(1) | abi.goal_vel(2))
    bc = W.synth_bc(12, 0, 64)
    out.update(gm_masks=np.array(masks, dtype=np.int64), gm_seed=np.int64(12), gm_n=np.int64(64))
    for j, mk in enumerate(masks):
        g = ref.generate(bc, mk)
        out[f"gm_cost_{j}"] = g["cost"]
        out[f"gm_coeffs_{j}"] = g["coeffs"]
        out[f"gm_abg_{j}"] = g["abg"]

    # ---- moving obstacles (Trajectory::operator-, Trajectory.hpp:121-128) --------------------------------
    mspecs = W.moving_obstacles(31, 8, lambda b: ref.generate(b)["coeffs"])
    mspecs[3]["motion_t_start"] = 0.5 * mspecs[3]["motion_t_end"]  # a window that starts late
    mspecs[5]["motion_t_end"] = 0.05             # one that ends before most trajectories do
    mspecs.append(W.static_obstacles(32, 1, 0)[0])  # one static sphere mixed in
    mobs = B.make_obstacles(mspecs)
    bc = W.synth_bc(13, 0, 512)
    r = ref.generate_check(bc, mobs, len(mspecs), 0.002)
    for k, v in W.specs_to_arrays(mspecs).items():
        out["mv_obs_" + k] = v
    out.update(mv_seed=np.int64(13), mv_n=np.int64(512), mv_matrix=r["verdict_matrix"], mv_verdict=r["verdict"],
               mv_first=r["first_obstacle"])

    # ---- pre-built quintics with their own windows (rqcd_check) ---------------------------------------
    g = ref.generate(W.synth_bc(14, 0, 256))
    rng = np.random.default_rng(14)
    t_end = W.synth_bc(14, 0, 256)[abi.BC_TF]
    t_start = t_end * rng.uniform(0.0, 0.5, 256)
    cspecs = W.static_obstacles(22, 3, 3)
    r = ref.check(g["coeffs"], t_end, B.make_obstacles(cspecs), 6, 0.002, t_start=t_start, cost=g["cost"])
    for k, v in W.specs_to_arrays(cspecs).items():
        out["ck_obs_" + k] = v
    out.update(ck_coeffs=g["coeffs"], ck_t_start=t_start, ck_t_end=t_end, ck_cost=g["cost"],
               ck_matrix=r["verdict_matrix"], ck_verdict=r["verdict"],
               ck_best_index=np.int64(r["summary"]["best_index"]))

    # ---- solver vectors -------------------------------------------------------------------------------
    rng = np.random.default_rng(41)
    q = np.concatenate([adversarial_quartics(), rng.uniform(-10, 10, (4, 2048)), rng.normal(0, 1e3, (4, 512))], axis=1)
    roots, cnt = ref.solve_quartic(q)
    out.update(sq_coef=q, sq_roots=roots, sq_count=cnt)
    c = np.concatenate([adversarial_cubics(), rng.uniform(-10, 10, (3, 2048)), rng.normal(0, 1e3, (3, 512))], axis=1)
    roots, cnt = ref.solve_cubic(c)
    out.update(sc_coef=c, sc_roots=roots, sc_count=cnt)

    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)
    for f in ("forest_golden.npz", "golden_counts.json", "ref_vectors.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
    print(counts)


if __name__ == "__main__":
    main()
