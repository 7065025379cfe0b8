"""CPU: pins oracle/oracle.c (the restatement) to the reference's checked-in known answers and to vectors
returned by the unmodified reference classes (tests/golden/make_golden.py). No GPU involved."""
import numpy as np
import pytest

from rapidquadcoptercollisiondetection_b200 import abi, workloads as W
from tests.util import degenerate_case, obstacles_from

MIN_T = 0.002


def test_c2_forest_counts_match_reference_data_file(oracle, golden_counts):
    """data/ForestPerfEval.txt:28-29: 63.9533 % input feasible, 60.381 % collision free of 10^6."""
    bc = oracle.c2_inputs(10_000, 100)
    r = oracle.generate_check(bc, oracle.forest_obstacles(), 5, MIN_T, flags=abi.F_EARLY_EXIT, matrix=False, nthreads=8)
    assert r["summary"]["traj_counts"][0] == golden_counts["c2_collision_free"]
    group = np.repeat(np.arange(10_000, dtype=np.int64), 100)  # one generator object per batch
    feas = oracle.input_feasibility(bc, group=group)
    assert int((feas == 0).sum()) == golden_counts["c2_input_feasible"]


def test_c1_performance_eval_counts_match_reference_data_file(oracle, golden_counts):
    """data/PerformanceEval.txt:23-25: 9 600 380 / 399 612 / 8 of 10^7 accepted trials."""
    n = golden_counts["c1_n"]
    bc, sph, trials = oracle.c1_inputs(n, filter=True)
    r = oracle.generate_check_pairwise(bc, sph, MIN_T, nthreads=8)
    assert r["summary"]["traj_counts"][:3] == [golden_counts["c1_no_collision"], golden_counts["c1_collision"],
                                              golden_counts["c1_indeterminable"]]
    assert trials > n


def test_forest_csv_batch0(oracle, forest_golden):
    """data/ForestPerfEval.csv: coefficients to 6 significant digits, Tf, inputFeas, stateFeas of batch 0;
    data/ForestPerfEvaltreeParamsCSV.csv: the five prisms."""
    csv = forest_golden["csv"]
    bc = oracle.c2_inputs(1, 100)
    obs = oracle.forest_obstacles()
    r = oracle.generate_check(bc, obs, 5, MIN_T, flags=abi.F_EARLY_EXIT, matrix=False)
    np.testing.assert_allclose(r["coeffs"].T, csv[:, :18], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(bc[abi.BC_TF], csv[:, 18], rtol=1e-5)
    np.testing.assert_array_equal(r["verdict"], csv[:, 20].astype(np.uint8))
    feas = oracle.input_feasibility(bc, group=np.zeros(100, dtype=np.int64))
    np.testing.assert_array_equal(feas, csv[:, 19].astype(np.uint8))
    trees = forest_golden["trees"]
    for k in range(5):
        np.testing.assert_allclose(list(obs[k].center), trees[k, 0:3])
        np.testing.assert_allclose(list(obs[k].side), trees[k, 3:6])
        np.testing.assert_allclose(oracle.rotation_matrix(list(obs[k].quat)), trees[k, 6:15], atol=1e-6)
    # the host-side workload builder must hand the GPU exactly the same five prisms
    mine = abi.make_obstacles(W.forest_obstacles())
    for k in range(5):
        assert list(mine[k].quat) == list(obs[k].quat) and list(mine[k].center) == list(obs[k].center)


def test_driver_streams_match_reference_run(oracle, vectors):
    """mt19937(0) + libstdc++ uniform_real_distribution draw order, acceptance filter and verdicts of the
    unmodified drivers' loops (ref_shim.cpp ref_c1_run / ref_c2_run)."""
    n = vectors["c1_bc"].shape[1]
    bc, sph, trials = oracle.c1_inputs(n, filter=True)
    np.testing.assert_array_equal(bc, vectors["c1_bc"])
    np.testing.assert_array_equal(sph, vectors["c1_spheres"])
    assert trials == int(vectors["c1_trials"])
    r = oracle.generate_check_pairwise(bc, sph, MIN_T)
    np.testing.assert_array_equal(r["verdict"], vectors["c1_verdict"])

    nb = vectors["c2_bc"].shape[1] // 100
    bc = oracle.c2_inputs(nb, 100)
    np.testing.assert_array_equal(bc, vectors["c2_bc"])
    r = oracle.generate_check(bc, oracle.forest_obstacles(), 5, MIN_T, flags=abi.F_EARLY_EXIT, matrix=False)
    np.testing.assert_array_equal(r["verdict"], vectors["c2_verdict"])
    assert r["summary"]["pairs_checked"] == int(vectors["c2_pairs"])
    feas = oracle.input_feasibility(bc, group=np.repeat(np.arange(nb, dtype=np.int64), 100))
    np.testing.assert_array_equal(feas, vectors["c2_input_feas"])


def test_static_set_all_pairs(oracle, vectors):
    obs, m = obstacles_from(vectors, "st_obs_")
    bc = W.synth_bc(int(vectors["st_seed"]), 0, int(vectors["st_n"]))
    np.testing.assert_array_equal(bc, oracle.synth_bc(int(vectors["st_seed"]), 0, int(vectors["st_n"])))
    r = oracle.generate_check(bc, obs, m, MIN_T)
    np.testing.assert_array_equal(r["verdict_matrix"], vectors["st_matrix"])
    np.testing.assert_array_equal(r["verdict"], vectors["st_verdict"])
    np.testing.assert_array_equal(r["first_obstacle"], vectors["st_first"])
    np.testing.assert_array_equal(r["cost"], vectors["st_cost"])      # bit for bit: same operation order
    np.testing.assert_array_equal(r["coeffs"], vectors["st_coeffs"])
    assert r["summary"]["best_index"] == int(vectors["st_best_index"])
    assert r["summary"]["best_cost"] == float(vectors["st_best_cost"])
    assert (vectors["st_matrix"] == 1).any() and (vectors["st_matrix"] == 0).any()


def test_goal_flag_combinations(oracle, vectors):
    """All eight branches of SingleAxisTrajectory.cpp:71-116 plus a mixed per-axis mask."""
    bc = W.synth_bc(int(vectors["gm_seed"]), 0, int(vectors["gm_n"]))
    for j, mask in enumerate(vectors["gm_masks"]):
        g = oracle.generate(bc, int(mask))
        np.testing.assert_array_equal(g["cost"], vectors[f"gm_cost_{j}"])
        np.testing.assert_array_equal(g["coeffs"], vectors[f"gm_coeffs_{j}"])
        np.testing.assert_array_equal(g["abg"], vectors[f"gm_abg_{j}"])


def test_moving_obstacles(oracle, vectors):
    obs, m = obstacles_from(vectors, "mv_obs_")
    bc = W.synth_bc(int(vectors["mv_seed"]), 0, int(vectors["mv_n"]))
    r = oracle.generate_check(bc, obs, m, MIN_T)
    np.testing.assert_array_equal(r["verdict_matrix"], vectors["mv_matrix"])
    np.testing.assert_array_equal(r["verdict"], vectors["mv_verdict"])
    np.testing.assert_array_equal(r["first_obstacle"], vectors["mv_first"])
    assert (vectors["mv_matrix"] == 1).any()


def test_prebuilt_quintics_with_windows(oracle, vectors):
    obs, m = obstacles_from(vectors, "ck_obs_")
    r = oracle.check(vectors["ck_coeffs"], vectors["ck_t_end"], obs, m, MIN_T, t_start=vectors["ck_t_start"],
                     cost=vectors["ck_cost"])
    np.testing.assert_array_equal(r["verdict_matrix"], vectors["ck_matrix"])
    np.testing.assert_array_equal(r["verdict"], vectors["ck_verdict"])
    assert r["summary"]["best_index"] == int(vectors["ck_best_index"])


def test_solver_vectors(oracle, vectors):
    """Quartic::solve_quartic / solveP3 incl. t^4+1 (NaN roots), repeated roots, non-finite input."""
    roots, cnt = oracle.solve_quartic(vectors["sq_coef"])
    np.testing.assert_array_equal(cnt, vectors["sq_count"])
    np.testing.assert_array_equal(roots, vectors["sq_roots"])  # same libm, same order: bit for bit (NaN == NaN)
    roots, cnt = oracle.solve_cubic(vectors["sc_coef"])
    np.testing.assert_array_equal(cnt, vectors["sc_count"])
    np.testing.assert_array_equal(roots, vectors["sc_roots"])


def test_oracle_equals_live_reference_on_fresh_inputs(oracle, reference):
    """Where oracle/_ref/libref.so is present: fresh seeds, restatement vs unmodified reference, bit for bit."""
    specs = W.static_obstacles(77, 6, 6)
    obs = abi.make_obstacles(specs)
    bc = W.synth_bc(78, 1000, 3000)
    a = oracle.generate_check(bc, obs, 12, MIN_T)
    b = reference.generate_check(bc, obs, 12, MIN_T)
    for k in ("verdict_matrix", "verdict", "first_obstacle", "cost", "coeffs"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    assert a["summary"] == b["summary"]
    rng = np.random.default_rng(5)
    q = rng.uniform(-50, 50, (4, 20000))
    ra, ca = oracle.solve_quartic(q)
    rb, cb = reference.solve_quartic(q)
    np.testing.assert_array_equal(ca, cb)
    np.testing.assert_array_equal(ra, rb)
    feas_a = oracle.input_feasibility(bc)
    feas_b = reference.input_feasibility(bc)
    np.testing.assert_array_equal(feas_a, feas_b)


def test_degenerate_inputs_oracle_equals_live_reference(oracle, reference):
    """Zero / tiny / huge durations, overflow, NaN states, points exactly on a surface, a zero quaternion: the
    restatement follows the unmodified reference through every NaN and inf."""
    bc, specs = degenerate_case()
    obs = abi.make_obstacles(specs)
    for min_t in (0.002, 1e-5):
        a = oracle.generate_check(bc, obs, len(specs), min_t)
        b = reference.generate_check(bc, obs, len(specs), min_t)
        np.testing.assert_array_equal(a["verdict_matrix"], b["verdict_matrix"])
        assert a["summary"] == b["summary"]
