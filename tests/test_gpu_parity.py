"""GPU: the CUDA path, called through the C ABI (librqcd.so), against the committed reference vectors
(tests/golden/) and against the CPU oracle on the same seeded inputs. Verdicts, first-hit indices, counts and
best-candidate indices are compared bit-exact; coefficients, costs and roots within 1e-9 relative
(BASELINE.json), written in tests/util.py."""
import numpy as np
import pytest

from rapidquadcoptercollisiondetection_b200 import abi, workloads as W
from tests.util import assert_close, assert_roots_match, degenerate_case, grazing_case, obstacles_from

pytestmark = pytest.mark.gpu
MIN_T = 0.002


def test_device_is_blackwell(ctx):
    info = ctx.device_info()
    assert info["cc"][0] >= 10 and info["sm_count"] > 0


def test_static_set_matches_reference_vectors(ctx, vectors):
    obs, m = obstacles_from(vectors, "st_obs_")
    ctx.set_obstacles(obs, m)
    bc = W.synth_bc(int(vectors["st_seed"]), 0, int(vectors["st_n"]))
    before = ctx.launch_count()
    r = ctx.generate_check(bc, MIN_T)
    assert ctx.launch_count() == before + 2
    assert r["status"] == abi.OK
    np.testing.assert_array_equal(r["verdict_matrix"], vectors["st_matrix"])
    np.testing.assert_array_equal(r["verdict"], vectors["st_verdict"])
    np.testing.assert_array_equal(r["first_obstacle"], vectors["st_first"])
    assert_close(r["cost"], vectors["st_cost"], what="cost")
    assert_close(r["coeffs"], vectors["st_coeffs"], what="coeffs")
    s = r["summary"]
    assert s["best_index"] == int(vectors["st_best_index"])
    assert s["pairs_checked"] == bc.shape[1] * m
    assert s["pair_counts"][:3] == [int((vectors["st_matrix"] == v).sum()) for v in range(3)]
    assert s["traj_counts"][:3] == [int((vectors["st_verdict"] == v).sum()) for v in range(3)]


def test_early_exit_is_the_forest_loop(ctx, vectors, oracle):
    """ForestPerformanceEval.cpp:216-225: stop at the first obstacle whose check is not NoCollision."""
    obs, m = obstacles_from(vectors, "st_obs_")
    ctx.set_obstacles(obs, m)
    bc = W.synth_bc(int(vectors["st_seed"]), 0, int(vectors["st_n"]))
    r = ctx.generate_check(bc, MIN_T, flags=abi.F_EARLY_EXIT)
    o = oracle.generate_check(bc, obs, m, MIN_T, flags=abi.F_EARLY_EXIT, matrix=False)
    np.testing.assert_array_equal(r["verdict"], vectors["st_verdict"])
    np.testing.assert_array_equal(r["first_obstacle"], vectors["st_first"])
    assert r["summary"]["pairs_checked"] == o["summary"]["pairs_checked"]
    assert r["summary"]["pair_counts"] == o["summary"]["pair_counts"]


def test_forest_driver_vectors(ctx, vectors, forest_golden):
    """Batch stream of ForestPerformanceEval (ref_c2_run) incl. batch 0 = data/ForestPerfEval.csv."""
    ctx.set_obstacles(abi.make_obstacles(W.forest_obstacles()), 5)
    bc = vectors["c2_bc"]
    r = ctx.generate_check(bc, MIN_T, flags=abi.F_EARLY_EXIT)
    np.testing.assert_array_equal(r["verdict"], vectors["c2_verdict"])
    assert r["summary"]["pairs_checked"] == int(vectors["c2_pairs"])
    csv = forest_golden["csv"]
    np.testing.assert_array_equal(r["verdict"][:100], csv[:, 20].astype(np.uint8))
    np.testing.assert_allclose(r["coeffs"][:, :100].T, csv[:, :18], rtol=2e-5, atol=1e-6)


def test_performance_eval_driver_vectors(ctx, vectors):
    """Trial stream of PerformanceEval (ref_c1_run): trajectory i against its own sphere i."""
    r = ctx.generate_check_pairwise(vectors["c1_bc"], vectors["c1_spheres"], MIN_T)
    np.testing.assert_array_equal(r["verdict"], vectors["c1_verdict"])
    n = vectors["c1_bc"].shape[1]
    assert r["summary"]["pairs_checked"] == n and sum(r["summary"]["traj_counts"]) == n


def test_goal_flag_combinations(ctx, vectors):
    bc = W.synth_bc(int(vectors["gm_seed"]), 0, int(vectors["gm_n"]))
    for j, mask in enumerate(vectors["gm_masks"]):
        g = ctx.generate(bc, int(mask))
        assert_close(g["cost"], vectors[f"gm_cost_{j}"], what=f"cost mask {int(mask):#x}")
        assert_close(g["coeffs"], vectors[f"gm_coeffs_{j}"], what=f"coeffs mask {int(mask):#x}")
        # GetAxisParamAlpha/Beta/Gamma of the unmodified reference (rqcd_out::params)
        assert_close(g["abg"], vectors[f"gm_abg_{j}"], what=f"alpha/beta/gamma mask {int(mask):#x}")


def test_moving_obstacles_match_reference_vectors(ctx, vectors):
    obs, m = obstacles_from(vectors, "mv_obs_")
    ctx.set_obstacles(obs, m)
    bc = W.synth_bc(int(vectors["mv_seed"]), 0, int(vectors["mv_n"]))
    r = ctx.generate_check(bc, MIN_T)
    np.testing.assert_array_equal(r["verdict_matrix"], vectors["mv_matrix"])
    np.testing.assert_array_equal(r["verdict"], vectors["mv_verdict"])
    np.testing.assert_array_equal(r["first_obstacle"], vectors["mv_first"])


def test_prebuilt_quintics_with_windows(ctx, vectors):
    obs, m = obstacles_from(vectors, "ck_obs_")
    ctx.set_obstacles(obs, m)
    r = ctx.check(vectors["ck_coeffs"], vectors["ck_t_end"], MIN_T, t_start=vectors["ck_t_start"], cost=vectors["ck_cost"])
    np.testing.assert_array_equal(r["verdict_matrix"], vectors["ck_matrix"])
    np.testing.assert_array_equal(r["verdict"], vectors["ck_verdict"])
    assert r["summary"]["best_index"] == int(vectors["ck_best_index"])


def test_solver_vectors(ctx, vectors):
    """The reference's own solver outputs (tests/golden). Cubic entry 7 is x^3 + 1e6 x^2 + x + 1: r^2 - q^3 cancels
    catastrophically, the imaginary part the reference reports (4e-10, true value 1e-3) is rounding noise of ITS pow()
    and decides `fabs(x[2]) < eps` -- rqcd_solve_conditioning must say so (branch-fragile), not a skip list."""
    roots, cnt = ctx.solve_quartic(vectors["sq_coef"])
    cond, fr = ctx.solve_conditioning(vectors["sq_coef"])
    print("quartic vectors:", assert_roots_match(vectors["sq_coef"], roots, cnt, vectors["sq_roots"], vectors["sq_count"], cond, fr, "quartic"))
    roots, cnt = ctx.solve_cubic(vectors["sc_coef"])
    cond, fr = ctx.solve_conditioning(vectors["sc_coef"])
    print("cubic vectors:", assert_roots_match(vectors["sc_coef"], roots, cnt, vectors["sc_roots"], vectors["sc_count"], cond, fr, "cubic"))
    assert fr[7] == 1


def test_solver_random_against_oracle(ctx, oracle):
    rng = np.random.default_rng(9)
    q = np.concatenate([rng.uniform(-10, 10, (4, 200_000)), rng.normal(0, 100, (4, 50_000))], axis=1)
    ra, ca = ctx.solve_quartic(q)
    rb, cb = oracle.solve_quartic(q)
    cond, fr = ctx.solve_conditioning(q)
    st = assert_roots_match(q, ra, ca, rb, cb, cond, fr, "quartic")
    print("quartic random:", st)
    assert st["branch_fragile"] <= 1e-3 * st["problems"]  # the flag is specific
    c = rng.uniform(-10, 10, (3, 200_000))
    ra, ca = ctx.solve_cubic(c)
    rb, cb = oracle.solve_cubic(c)
    cond, fr = ctx.solve_conditioning(c)
    st = assert_roots_match(c, ra, ca, rb, cb, cond, fr, "cubic")
    print("cubic random:", st)
    assert st["branch_fragile"] <= 1e-3 * st["problems"]


@pytest.mark.parametrize("n,ns,npz,seed", [(20_000, 32, 32, 101), (50_000, 5, 0, 102), (3_000, 0, 40, 103), (777, 1, 0, 104)])
def test_random_sets_against_oracle(ctx, oracle, n, ns, npz, seed):
    """Same seeded inputs through the CUDA path and the oracle: every pair's verdict, first hits, counts, best."""
    specs = W.static_obstacles(seed, ns, npz)
    obs = abi.make_obstacles(specs)
    m = len(specs)
    ctx.set_obstacles(obs, m)
    bc = W.synth_bc(seed, 12345, n)
    r = ctx.generate_check(bc, MIN_T, index_base=12345)
    o = oracle.generate_check(bc, obs, m, MIN_T, index_base=12345, nthreads=8)
    np.testing.assert_array_equal(r["verdict_matrix"], o["verdict_matrix"])
    np.testing.assert_array_equal(r["verdict"], o["verdict"])
    np.testing.assert_array_equal(r["first_obstacle"], o["first_obstacle"])
    assert_close(r["cost"], o["cost"], what="cost")
    assert_close(r["coeffs"], o["coeffs"], what="coeffs")
    assert r["summary"]["traj_counts"] == o["summary"]["traj_counts"]
    assert r["summary"]["pair_counts"] == o["summary"]["pair_counts"]
    assert r["summary"]["best_index"] == o["summary"]["best_index"]
    assert_close(r["summary"]["best_cost"], o["summary"]["best_cost"], what="best cost")


def test_both_kernels_of_the_static_all_pairs_shape(oracle, monkeypatch):
    """Obstacle tables are checked by k_pool (staged sections: plane stage with the certified solver skip, solve stage,
    pair tasks in a warp pool); RQCD_POOL=0 selects k_check (lane-owns-trajectory, the kernel of the pairwise shape)
    for them too. Both must give the oracle's verdicts, first hits, counts and best candidate -- with prebuilt
    coefficients and ragged sizes as well (a last batch of 9 candidates; a skipped prefix of the index space)."""
    from rapidquadcoptercollisiondetection_b200.api import Context
    specs = W.static_obstacles(201, 20, 17)
    obs = abi.make_obstacles(specs)
    m = len(specs)
    n = 32 * 311 + 9
    bc = W.synth_bc(202, 777, n)
    o = oracle.generate_check(bc, obs, m, MIN_T, index_base=777, nthreads=8)
    for pool in ("1", "0"):
        monkeypatch.setenv("RQCD_POOL", pool)
        with Context(0) as c2:
            c2.set_obstacles(obs, m)
            r = c2.generate_check(bc, MIN_T, index_base=777)
            assert c2.last_kernel_name().startswith("k_pool" if pool == "1" else "k_check")
            np.testing.assert_array_equal(r["verdict_matrix"], o["verdict_matrix"])
            np.testing.assert_array_equal(r["verdict"], o["verdict"])
            np.testing.assert_array_equal(r["first_obstacle"], o["first_obstacle"])
            assert r["summary"]["traj_counts"] == o["summary"]["traj_counts"]
            assert r["summary"]["pair_counts"] == o["summary"]["pair_counts"]
            assert r["summary"]["best_index"] == o["summary"]["best_index"]
            rc = c2.check(o["coeffs"], bc[abi.BC_TF], MIN_T, cost=o["cost"], index_base=777)
            np.testing.assert_array_equal(rc["verdict_matrix"], o["verdict_matrix"])
            assert rc["summary"]["best_index"] == o["summary"]["best_index"]


def test_edge_cases(ctx, oracle):
    """Empty batch, empty obstacle set, a one-trajectory batch, ragged strides, invalid arguments."""
    specs = W.static_obstacles(5, 2, 2)
    ctx.set_obstacles(abi.make_obstacles(specs), 4)
    r = ctx.generate_check(np.zeros((abi.BC_ROWS, 0)), MIN_T)
    assert r["summary"]["pairs_checked"] == 0 and r["summary"]["best_index"] == -1
    bc = W.synth_bc(6, 0, 1)
    r = ctx.generate_check(bc, MIN_T)
    o = oracle.generate_check(bc, abi.make_obstacles(specs), 4, MIN_T)
    np.testing.assert_array_equal(r["verdict_matrix"], o["verdict_matrix"])
    ctx.set_obstacles(abi.make_obstacles([]), 0)
    bc = W.synth_bc(6, 0, 100)
    r = ctx.generate_check(bc, MIN_T, matrix=False)
    assert r["summary"]["traj_counts"][0] == 100 and (r["verdict"] == 0).all() and (r["first_obstacle"] == -1).all()
    # minT <= 0 would make the reference recurse without bound: rejected
    from rapidquadcoptercollisiondetection_b200.api import RqcdError
    with pytest.raises(RqcdError):
        ctx.generate_check(bc, 0.0)
    with pytest.raises(RqcdError):
        ctx.set_obstacles(abi.make_obstacles([{"type": "sphere", "center": [0, 0, 0], "radius": -1.0}]), 1)


def test_indeterminable_and_deep_recursion(ctx, oracle):
    """Trajectories grazing a sphere: long recursions, Indeterminable verdicts, small minT (deep stacks)."""
    n = 4096
    rng = np.random.default_rng(3)
    bc = np.zeros((abi.BC_ROWS, n))
    # straight-ish lines passing at distance ~r from the origin
    bc[abi.BC_P0 + 0] = -3.0
    bc[abi.BC_P0 + 1] = 1.0 + rng.normal(0, 1e-7, n)
    bc[abi.BC_V0 + 0] = 2.0
    bc[abi.BC_PF + 0] = 3.0
    bc[abi.BC_PF + 1] = 1.0 + rng.normal(0, 1e-7, n)
    bc[abi.BC_VF + 0] = 2.0
    bc[abi.BC_TF] = 3.0
    specs = [{"type": "sphere", "center": [0.0, 0.0, 0.0], "radius": 1.0},
             {"type": "prism", "center": [0.0, 2.0, 0.0], "side": [1.0, 2.0, 1.0], "quat": [1.0, 0, 0, 0]}]
    obs = abi.make_obstacles(specs)
    ctx.set_obstacles(obs, 2)
    for min_t in (0.002, 1e-6, 1e-9):
        r = ctx.generate_check(bc, min_t)
        o = oracle.generate_check(bc, obs, 2, min_t, nthreads=8)
        np.testing.assert_array_equal(r["verdict_matrix"], o["verdict_matrix"])
        assert r["summary"]["pair_counts"] == o["summary"]["pair_counts"]
        assert r["status"] == abi.OK
    assert (o["verdict_matrix"] == 2).any() or (o["verdict_matrix"] == 1).any()


def test_plane_boundaries_against_oracle_intent(ctx, oracle):
    """CollisionCheck(Boundary) with the documented semantics (the reference's own code is UB there; the oracle
    restates the intent -- parity for this entry is unpinned against the reference binary)."""
    bc = W.synth_bc(55, 0, 20_000)
    g = oracle.generate(bc)
    rng = np.random.default_rng(56)
    planes = np.concatenate([rng.uniform(-3, 3, (12, 3)), rng.normal(0, 1, (12, 3))], axis=1)
    planes[0] = [0, 0, -1.0, 0, 0, 5.0]   # a floor at z = -1 with a non-unit normal
    planes[1] = [0, 0, 0, 1e-9, 0, 0]     # a tiny normal: normalisation goes through the float norm
    v, mtx = ctx.check_planes(g["coeffs"], bc[abi.BC_TF], planes)
    vo, mo = oracle.check_planes(g["coeffs"], bc[abi.BC_TF], planes)
    np.testing.assert_array_equal(mtx, mo)
    np.testing.assert_array_equal(v, vo)
    assert 0 < int(mo.sum()) < mo.size
    v2, m2 = ctx.generate_check_planes(bc, planes)   # same thing from boundary conditions
    np.testing.assert_array_equal(m2, mo)
    # windows that start late
    t_start = bc[abi.BC_TF] * 0.5
    v3, m3 = ctx.check_planes(g["coeffs"], bc[abi.BC_TF], planes, t_start=t_start)
    vo3, mo3 = oracle.check_planes(g["coeffs"], bc[abi.BC_TF], planes, t_start=t_start)
    np.testing.assert_array_equal(m3, mo3)


def test_input_feasibility_fresh_objects(ctx, oracle):
    """CheckInputFeasibility(5, 30, 20, 0.002) as PerformanceEval calls it: a new generator per trial."""
    bc = W.synth_bc(61, 0, 200_000)
    got = ctx.input_feasibility(bc)
    want = oracle.input_feasibility(bc)
    np.testing.assert_array_equal(got, want)
    assert len(np.unique(want)) >= 4   # feasible, indeterminable, thrust high, thrust low all occur
    for mask in (abi.goal_pos(0) | abi.goal_pos(1) | abi.goal_pos(2), 0x1FF & ~abi.goal_acc(2)):
        np.testing.assert_array_equal(ctx.input_feasibility(bc[:, :20_000], goal_mask=mask),
                                      oracle.input_feasibility(bc[:, :20_000], goal_mask=mask))


def test_input_feasibility_limits_at_the_edge_of_a_comparison(ctx, oracle):
    """k_feas decides a section's comparisons on squares (feas_frame_sq) and takes the roots themselves only when a
    square is within the roundings of a limit's. Limits placed exactly ON a candidate's thrust at t = 0 (and one ulp to
    either side), limits whose squares leave the certified range, zero and negative limits: the results stay the
    reference's for every candidate of the batch."""
    bc = W.synth_bc(62, 0, 8192)
    g = np.array([0.0, 0.0, -9.81])
    a0 = bc[abi.BC_A0:abi.BC_A0 + 3]
    e = a0 - g[:, None]
    thrust0 = np.sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2])  # RapidTrajectoryGenerator.hpp:192-194 at t = 0
    cases = []
    order = np.argsort(thrust0)
    for q, which in ((0.03, "fmin"), (0.10, "fmin"), (0.90, "fmax"), (0.97, "fmax")):
        f = float(thrust0[order[int(q * bc.shape[1])]])
        for lim in (f, np.nextafter(f, 0.0), np.nextafter(f, np.inf)):
            cases.append(dict(fmin=float(lim), fmax=30.0, wmax=20.0) if which == "fmin" else dict(fmin=5.0, fmax=float(lim), wmax=20.0))
    cases += [dict(fmin=0.0, fmax=30.0, wmax=20.0), dict(fmin=-1.0, fmax=30.0, wmax=20.0), dict(fmin=5.0, fmax=30.0, wmax=1e200),
              dict(fmin=5.0, fmax=30.0, wmax=1e-150), dict(fmin=1e-160, fmax=1e160, wmax=20.0),
              dict(fmin=5.0, fmax=float("inf"), wmax=float("inf")), dict(fmin=5.0, fmax=30.0, wmax=0.0)]
    seen = set()
    for kw in cases:
        got = ctx.input_feasibility(bc, **kw)
        want = oracle.input_feasibility(bc, **kw)
        np.testing.assert_array_equal(got, want, err_msg=str(kw))
        seen |= set(np.unique(want).tolist())
    assert seen >= {0, 1, 2, 3}


def test_input_feasibility_forest_batches_match_reference_counts(ctx, oracle, vectors, forest_golden, golden_counts):
    """The Forest driver reuses one generator per batch (stale acceleration-peak cache): its own stream gives
    63.9533 % input feasible of 10^6 (data/ForestPerfEval.txt:28) and the inputFeas column of the CSV."""
    bc = vectors["c2_bc"]
    nb = bc.shape[1] // 100
    group = np.repeat(np.arange(nb, dtype=np.int64), 100)
    got = ctx.input_feasibility(bc, group=group)
    np.testing.assert_array_equal(got, vectors["c2_input_feas"])             # the unmodified reference's values
    np.testing.assert_array_equal(got[:100], forest_golden["csv"][:, 19].astype(np.uint8))
    bc = oracle.c2_inputs(10_000, 100)
    group = np.repeat(np.arange(10_000, dtype=np.int64), 100)
    got = ctx.input_feasibility(bc, group=group)
    assert int((got == 0).sum()) == golden_counts["c2_input_feasible"]
    np.testing.assert_array_equal(got, oracle.input_feasibility(bc, group=group))


def test_reference_drivers_known_answers_at_full_size(ctx, oracle, golden_counts):
    """The reference's checked-in outputs reproduced by the CUDA path on the drivers' own input streams:
    data/PerformanceEval.txt:23-25 -> 9 600 380 / 399 612 / 8 of 10^7 accepted trials (pairwise spheres), and
    data/ForestPerfEval.txt:28-29 -> 639 533 input feasible / 603 810 collision free of 10^6 (five prisms, early exit).
    (The oracle only draws the inputs here -- mt19937 stream and acceptance filter; the counts come from the GPU.)"""
    bc = oracle.c2_inputs(10_000, 100)
    ctx.set_obstacles(abi.make_obstacles(W.forest_obstacles()), 5)
    r = ctx.generate_check(bc, MIN_T, flags=abi.F_EARLY_EXIT, matrix=False)
    assert r["summary"]["traj_counts"][0] == golden_counts["c2_collision_free"]
    feas = ctx.input_feasibility(bc, group=np.repeat(np.arange(10_000, dtype=np.int64), 100))
    assert int((feas == 0).sum()) == golden_counts["c2_input_feasible"]

    n = golden_counts["c1_n"]
    bc, sph, _ = oracle.c1_inputs(n, filter=True)
    # the acceptance filter itself, on the device: every accepted trial is input feasible for a fresh generator
    assert int((ctx.input_feasibility(bc[:, :500_000].copy()) != 0).sum()) == 0
    r = ctx.generate_check_pairwise(bc, sph, MIN_T)
    assert r["summary"]["traj_counts"][:3] == [golden_counts["c1_no_collision"], golden_counts["c1_collision"],
                                              golden_counts["c1_indeterminable"]]


def test_degenerate_inputs_follow_the_reference_arithmetic(ctx, oracle):
    """Inputs the reference neither rejects nor handles specially -- zero / tiny / huge durations, overflowing states,
    NaN boundary conditions, points exactly on an obstacle's surface -- must take the same IEEE path on the GPU
    (NaN compares false everywhere, `<=` counts touching as inside, tiny windows are Indeterminable)."""
    bc, specs = degenerate_case()
    obs = abi.make_obstacles(specs)
    ctx.set_obstacles(obs, len(specs))
    for min_t in (0.002, 1e-5):
        r = ctx.generate_check(bc, min_t)
        o = oracle.generate_check(bc, obs, len(specs), min_t, nthreads=8)
        np.testing.assert_array_equal(r["verdict_matrix"], o["verdict_matrix"])
        np.testing.assert_array_equal(r["verdict"], o["verdict"])
        assert r["summary"]["traj_counts"] == o["summary"]["traj_counts"]
        assert r["summary"]["best_index"] == o["summary"]["best_index"]
    assert set(np.unique(o["verdict_matrix"])) == {0, 1, 2}


def test_fragile_mark_is_rare_on_monte_carlo(ctx, oracle):
    """RQCD_F_FRAGILE on the Monte Carlo shapes (static table, moving table, early exit, pairwise): verdicts equal the
    oracle's on every pair, the marks change nothing, and fewer than 1e-4 of the pairs are marked."""
    specs = W.static_obstacles(2, 32, 32)
    obs = abi.make_obstacles(specs)
    bc = W.synth_bc(1, 0, 1 << 15)
    ctx.set_obstacles(obs, len(specs))
    r = ctx.generate_check(bc, MIN_T, fragile=True)
    o = oracle.generate_check(bc, obs, len(specs), MIN_T, nthreads=8)
    np.testing.assert_array_equal(r["verdict_matrix"], o["verdict_matrix"])
    np.testing.assert_array_equal(r["verdict"], o["verdict"])
    assert (r["fragile_matrix"] != 0).mean() < 1e-4
    assert r["fragile"].max() <= 1
    e = ctx.generate_check(bc, MIN_T, flags=abi.F_EARLY_EXIT, fragile=True)
    oe = oracle.generate_check(bc, obs, len(specs), MIN_T, flags=abi.F_EARLY_EXIT, matrix=False, nthreads=8)
    np.testing.assert_array_equal(e["verdict"], oe["verdict"])
    np.testing.assert_array_equal(e["first_obstacle"], oe["first_obstacle"])
    assert e["fragile"].mean() < 1e-3
    # a trajectory marked in early-exit mode is marked in all-pairs mode too (same pairs up to the first hit)
    assert not ((e["fragile"] != 0) & (r["fragile"] == 0)).any()
    mspecs = W.moving_obstacles(4, 16, lambda b: ctx.generate(b)["coeffs"])
    mobs = abi.make_obstacles(mspecs)
    ctx.set_obstacles(mobs, len(mspecs))
    rm = ctx.generate_check(bc[:, :8192], MIN_T, fragile=True)
    om = oracle.generate_check(bc[:, :8192], mobs, len(mspecs), MIN_T, nthreads=8)
    np.testing.assert_array_equal(rm["verdict_matrix"], om["verdict_matrix"])
    assert (rm["fragile_matrix"] != 0).mean() < 1e-3
    u = W.synth_bc(8, 0, 1 << 15)
    sph = np.empty((4, 1 << 15))
    sph[0:3] = u[3:6]
    sph[3] = 0.1 + (u[6] + 4.0) / 8.0 * 1.4
    rp = ctx.generate_check_pairwise(bc, sph, MIN_T, fragile=True)
    op = oracle.generate_check_pairwise(bc, sph, MIN_T, nthreads=8)
    np.testing.assert_array_equal(rp["verdict"], op["verdict"])
    assert rp["fragile"].mean() < 1e-3
    # rest-to-rest primitives (the Forest driver's: goal velocity = acceleration = 0) have an exact double root of every
    # velocity polynomial at t = Tf; test points that come and go there decide nothing and must not flood the mark
    ctx.set_obstacles(abi.make_obstacles(W.forest_obstacles()), 5)
    c2 = W.c2_candidates(200)
    rf = ctx.generate_check(c2, MIN_T, flags=abi.F_EARLY_EXIT, fragile=True)
    of = oracle.generate_check(c2, abi.make_obstacles(W.forest_obstacles()), 5, MIN_T, flags=abi.F_EARLY_EXIT, matrix=False, nthreads=8)
    np.testing.assert_array_equal(rf["verdict"], of["verdict"])
    assert rf["fragile"].mean() < 0.1
    print("fragile marks: static", int((r["fragile_matrix"] != 0).sum()), "early-exit trajectories", int(e["fragile"].sum()),
          "moving", int((rm["fragile_matrix"] != 0).sum()), "pairwise", int(rp["fragile"].sum()), "reasons", ctx.fragile_reasons().tolist())


def test_grazing_and_badly_scaled_trajectories(ctx, oracle):
    """The two places where the kernel leaves its fast forms and must land on the reference's literal arithmetic:
    (1) plane-side tests whose value is within the certified margin of zero -- trajectories that graze an obstacle's
    tangent plane at a critical time, so that the SIGN depends on GetValue's operation order -- are re-evaluated
    literally; (2) operands outside the range of the in-line division/sqrt sequences (scenes scaled by 1e-150, 1e-80,
    1e+60, 1e+100) redo the frame section with the compiler's IEEE operators. Verdicts must equal the oracle's.
    With coefficient ratios of 1e3..1e8 the reference's own closed-form roots are off by more than the grazing
    distance, and a last-bit difference in acos / cos / pow (glibc there, polynomials here) moves the computed
    critical point across the plane. Those pairs are not tolerated by a quota: RQCD_F_FRAGILE marks the pairs in whose
    recursion some sign decision is closer to zero than the cubic's transcendental outputs (moved by 2^-50) and the
    decision's own rounding noise can move it, and EVERY pair that differs from the oracle must carry the mark, at every
    ratio -- outside the marked set verdicts are bit-exact. (On Monte Carlo batches the mark is rare:
    test_fragile_mark_is_rare_on_monte_carlo.)"""
    coeffs, t_end, specs = grazing_case(log_a=(1.0, 2.0))
    obs = abi.make_obstacles(specs)
    ctx.set_obstacles(obs, len(specs))  # 20 obstacles: end-point pre-pass kernel
    r = ctx.check(coeffs, t_end, MIN_T)
    o = oracle.check(coeffs, t_end, obs, len(specs), MIN_T, nthreads=8)
    np.testing.assert_array_equal(r["verdict_matrix"], o["verdict_matrix"])
    assert all(len(np.unique(o["verdict_matrix"][:, k])) == 3 for k in range(3))  # all three verdicts occur
    ctx.set_obstacles(obs, 3)  # 3 obstacles: end-point tests inside the frames
    r = ctx.check(coeffs, t_end, MIN_T)
    np.testing.assert_array_equal(r["verdict_matrix"], o["verdict_matrix"][:, :3])

    for log_a in ((1.0, 2.0), (1.0, 4.0), (3.0, 6.0), (1.0, 8.0), (6.0, 8.0)):
        coeffs, t_end, specs = grazing_case(log_a=log_a)
        ctx.set_obstacles(obs, len(specs))
        r = ctx.check(coeffs, t_end, MIN_T, fragile=True)
        plain = ctx.check(coeffs, t_end, MIN_T)
        np.testing.assert_array_equal(r["verdict_matrix"], plain["verdict_matrix"])  # the marks change no verdict
        o = oracle.check(coeffs, t_end, obs, len(specs), MIN_T, nthreads=8)
        differs = r["verdict_matrix"] != o["verdict_matrix"]
        marked = r["fragile_matrix"] != 0
        print(f"grazing log_a={log_a}: {int(differs.sum())} pairs differ, {int(marked.sum())} of {marked.size} marked fragile")
        assert not (differs & ~marked).any(), f"{int((differs & ~marked).sum())} unmarked mismatches at log_a={log_a}"
        # the three near obstacles of this population graze by construction (within 1e-13 .. 1e-6 of the surface with
        # coefficients up to 1e8: below the rounding noise of the reference's own plane-side evaluation for most of them);
        # the 17 far spheres are crossed at speeds up to 1e8 and are rarely marked
        assert marked[:, :3].mean() < 0.75 and marked[:, 3:].mean() < 0.01
        # per trajectory: a differing verdict / first hit implies the trajectory's mark
        t_differs = (r["verdict"] != o["verdict"]) | (r["first_obstacle"] != o["first_obstacle"])
        assert not (t_differs & (r["fragile"] == 0)).any()

    # badly scaled scenes: the same random scene multiplied by a power of ten (positions, sizes and coefficients)
    base = W.synth_bc(88, 0, 4096)
    for scale in (1e-150, 1e-80, 1e60, 1e100):
        bc = base.copy()
        bc[:abi.BC_TF] *= scale
        sp = [{"type": "sphere", "center": [1.5 * scale, 0.5 * scale, 0.0], "radius": 1.0 * scale},
              {"type": "prism", "center": [-1.0 * scale, 1.0 * scale, 0.5 * scale], "side": [1.0 * scale, 2.0 * scale, 0.7 * scale],
               "quat": [0.9238795325112867, 0, 0.3826834323650898, 0]}] + \
             [{"type": "sphere", "center": [(k - 8.0) * scale, -2.0 * scale, 1.0 * scale], "radius": 0.8 * scale} for k in range(16)]
        ob = abi.make_obstacles(sp)
        ctx.set_obstacles(ob, len(sp))
        r = ctx.generate_check(bc, MIN_T)
        o = oracle.generate_check(bc, ob, len(sp), MIN_T, nthreads=8)
        np.testing.assert_array_equal(r["verdict_matrix"], o["verdict_matrix"])
        np.testing.assert_array_equal(r["verdict"], o["verdict"])


def test_device_pointer_mode_equals_host_mode(ctx, vectors):
    """RQCD_F_DEVICE_PTRS: every entry point reads and writes device memory on the context's stream; results equal
    the host-pointer calls. (torch is used only to own the device buffers.)"""
    import ctypes as C

    import torch
    dev = torch.device("cuda", 0)
    lib, h = ctx.lib, ctx.h
    obs, m = obstacles_from(vectors, "ck_obs_")
    ctx.set_obstacles(obs, m)
    coeffs, t_end, t_start, cost = (vectors[k] for k in ("ck_coeffs", "ck_t_end", "ck_t_start", "ck_cost"))
    n = coeffs.shape[1]
    host = ctx.check(coeffs, t_end, MIN_T, t_start=t_start, cost=cost)
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in
         dict(coeffs=coeffs, t_end=t_end, t_start=t_start, cost=cost).items()}
    ver = torch.zeros(n, dtype=torch.uint8, device=dev)
    first = torch.zeros(n, dtype=torch.int32, device=dev)
    mtx = torch.zeros((n, m), dtype=torch.uint8, device=dev)
    summ = abi.Summary()
    inp = abi.CoeffSoa(d["coeffs"].data_ptr(), n, d["t_start"].data_ptr(), d["t_end"].data_ptr())
    out = abi.Out(verdict=ver.data_ptr(), first_obstacle=first.data_ptr(), verdict_matrix=mtx.data_ptr(), summary=C.pointer(summ))
    torch.cuda.synchronize()
    assert lib.rqcd_check(h, C.byref(inp), C.c_void_p(d["cost"].data_ptr()), n, MIN_T, abi.F_DEVICE_PTRS, 0, C.byref(out)) == abi.OK
    np.testing.assert_array_equal(mtx.cpu().numpy(), host["verdict_matrix"])
    np.testing.assert_array_equal(first.cpu().numpy(), host["first_obstacle"])
    assert summ.as_dict() == host["summary"]

    # planes + input feasibility + generate, device mode
    bc = W.synth_bc(91, 0, 10_000)
    planes = np.array([[0, 0, -1.0, 0, 0, 1.0], [2.0, 0, 0, -1.0, 0.2, 0]])
    pv_host, pm_host = ctx.generate_check_planes(bc, planes)
    feas_host = ctx.input_feasibility(bc)
    gen_host = ctx.generate(bc)
    d_bc = torch.from_numpy(bc).to(dev)
    nb = bc.shape[1]
    pv = torch.zeros(nb, dtype=torch.uint8, device=dev)
    pm = torch.zeros((nb, 2), dtype=torch.uint8, device=dev)
    fe = torch.zeros(nb, dtype=torch.uint8, device=dev)
    co = torch.zeros(nb, dtype=torch.float64, device=dev)
    cf = torch.zeros((18, nb), dtype=torch.float64, device=dev)
    binp = abi.BcSoa(d_bc.data_ptr(), nb, abi.GOAL_ALL, 0)
    torch.cuda.synchronize()
    assert lib.rqcd_generate_check_planes(h, C.byref(binp), nb, planes.ctypes.data_as(C.c_void_p), 2, abi.F_DEVICE_PTRS,
                                          C.c_void_p(pv.data_ptr()), C.c_void_p(pm.data_ptr())) == abi.OK
    g = np.array([0.0, 0.0, -9.81])
    assert lib.rqcd_check_input_feasibility(h, C.byref(binp), nb, g.ctypes.data_as(C.c_void_p), 5.0, 30.0, 20.0, 0.002, None,
                                            abi.F_DEVICE_PTRS, C.c_void_p(fe.data_ptr())) == abi.OK
    gout = abi.Out(cost=co.data_ptr(), coeffs=cf.data_ptr(), coeff_stride=nb)
    assert lib.rqcd_generate(h, C.byref(binp), nb, abi.F_DEVICE_PTRS, C.byref(gout)) == abi.OK
    ctx.synchronize()
    np.testing.assert_array_equal(pm.cpu().numpy(), pm_host)
    np.testing.assert_array_equal(pv.cpu().numpy(), pv_host)
    np.testing.assert_array_equal(fe.cpu().numpy(), feas_host)
    np.testing.assert_array_equal(co.cpu().numpy(), gen_host["cost"])
    np.testing.assert_array_equal(cf.cpu().numpy(), gen_host["coeffs"])


def test_planner_chain_equals_the_stages_composed(ctx, oracle):
    """rqcd_plan_best = cost prune -> CheckInputFeasibility -> CollisionCheck(Boundary) -> CollisionCheck(obstacles)
    -> argmin GetCost, against the same chain composed from the oracle's stage functions."""
    n = 60_000
    bc = W.synth_bc(81, 500, n)
    bc[abi.BC_PF:abi.BC_PF + 3] *= 0.5     # shorter hops: more input-feasible candidates
    specs = W.static_obstacles(82, 3, 3, extent=3.0)
    obs = abi.make_obstacles(specs)
    ctx.set_obstacles(obs, len(specs))
    planes = np.array([[0, 0, -1.5, 0, 0, 1.0], [0, 3.0, 0, 0, -2.0, 0]])
    g = oracle.generate(bc)
    max_cost = float(np.percentile(g["cost"], 80))
    for use_input, use_planes in ((True, True), (False, True), (True, False), (False, False)):
        r = ctx.plan_best(bc, MIN_T, max_cost=max_cost, input_limits=(5.0, 30.0, 20.0) if use_input else None,
                          planes=planes if use_planes else None, index_base=500)
        want = np.zeros(n, dtype=np.uint8)
        want[~(g["cost"] <= max_cost)] = 1
        if use_input:
            feas = oracle.input_feasibility(bc)
            want[(want == 0) & (feas != 0)] = 2
        if use_planes:
            pv, _ = oracle.check_planes(g["coeffs"], bc[abi.BC_TF], planes)
            want[(want == 0) & (pv != 0)] = 3
        o = oracle.generate_check(bc, obs, len(specs), MIN_T, flags=abi.F_EARLY_EXIT, matrix=False, nthreads=8)
        hit = (want == 0) & (o["verdict"] != 0)
        want[hit] = 3 + o["verdict"][hit]
        np.testing.assert_array_equal(r["stage"], want)
        alive = np.flatnonzero(want == 0)
        assert alive.size > 0 and len(np.unique(want)) >= (3 + use_input + use_planes)
        best = alive[np.argmin(g["cost"][alive])]
        assert r["summary"]["best_index"] == 500 + best
        np.testing.assert_allclose(r["summary"]["best_cost"], g["cost"][best], rtol=1e-12)
        reached = want == 0
        reached |= want >= 4
        assert sum(r["summary"]["traj_counts"]) == int(reached.sum())
    # every candidate rejected before the obstacle stage: the kernel must come back with nothing, not hang
    r = ctx.plan_best(bc, MIN_T, max_cost=-1.0)
    assert (r["stage"] == 1).all() and r["summary"]["best_index"] == -1 and r["summary"]["pairs_checked"] == 0


def test_obstacle_table_limits(ctx, oracle):
    """The prepared table must fit one CTA's shared memory next to the per-thread rows: ~700 static obstacles still
    run (one CTA per SM), a set that cannot fit is refused with RQCD_ERR_UNSUPPORTED -- never truncated."""
    from rapidquadcoptercollisiondetection_b200.api import RqcdError
    specs = W.static_obstacles(15, 350, 350)
    obs = abi.make_obstacles(specs)
    ctx.set_obstacles(obs, 700)
    bc = W.synth_bc(16, 0, 2000)
    r = ctx.generate_check(bc, MIN_T)
    o = oracle.generate_check(bc, obs, 700, MIN_T, nthreads=16)
    np.testing.assert_array_equal(r["verdict_matrix"], o["verdict_matrix"])
    assert r["summary"]["pair_counts"] == o["summary"]["pair_counts"]
    big = W.static_obstacles(17, 1500, 1500)
    with pytest.raises(RqcdError) as e:
        ctx.set_obstacles(abi.make_obstacles(big), 3000)
    assert e.value.status == abi.ERR_UNSUPPORTED
    assert ctx.num_obstacles() == 700  # the previous set stays in place


def test_input_filter_reproduces_performance_eval(ctx, oracle):
    """rqcd_set_input_filter + rqcd_generate_check_pairwise on the driver's own trial stream (accepted or not) = the whole
    trial loop of test/PerformanceEval.cpp:117-182 on the device: CheckInputFeasibility keeps 64 % of the trials, only those
    are checked and counted. Against the oracle's filtered stream (pinned to data/PerformanceEval.txt by
    tests/test_oracle_golden.py): same accepted set, same verdicts."""
    n_acc = 200_000
    bc_a, sph_a, trials = oracle.c1_inputs(n_acc, filter=True)   # the accepted trials, in order
    bc, sph = W.c1_trials(trials)                                # every trial the driver drew to get them
    ctx.set_input_filter()
    try:
        r = ctx.generate_check_pairwise(bc, sph, MIN_T)
    finally:
        ctx.set_input_filter(on=False)
    feas = oracle.input_feasibility(bc)
    np.testing.assert_array_equal(r["input_feasibility"], feas)
    keep = feas == 0
    assert int(keep.sum()) == n_acc and keep[-1]
    np.testing.assert_array_equal(bc[:, keep], bc_a)
    o = oracle.generate_check_pairwise(bc_a, sph_a, MIN_T, nthreads=8)
    np.testing.assert_array_equal(r["verdict"][keep], o["verdict"])
    assert (r["verdict"][~keep] == abi.NOT_CHECKED).all()
    assert r["summary"]["traj_counts"] == o["summary"]["traj_counts"]
    assert r["summary"]["pairs_checked"] == n_acc
    # the table shape with the filter, early exit (Forest loop with the driver's inputFeas test in front of the check)
    specs = W.forest_obstacles()
    obs = abi.make_obstacles(specs)
    ctx.set_obstacles(obs, len(specs))
    c2 = W.c2_candidates(50)
    ctx.set_input_filter()
    try:
        f = ctx.generate_check(c2, MIN_T, flags=abi.F_EARLY_EXIT)
    finally:
        ctx.set_input_filter(on=False)
    fe = oracle.input_feasibility(c2)   # a fresh object per candidate (the filter's semantics)
    np.testing.assert_array_equal(f["input_feasibility"], fe)
    k2 = fe == 0
    o2 = oracle.generate_check(np.ascontiguousarray(c2[:, k2]), obs, len(specs), MIN_T, flags=abi.F_EARLY_EXIT, matrix=False, nthreads=4)
    np.testing.assert_array_equal(f["verdict"][k2], o2["verdict"])
    np.testing.assert_array_equal(f["first_obstacle"][k2], o2["first_obstacle"])
    assert (f["verdict"][~k2] == abi.NOT_CHECKED).all() and (f["first_obstacle"][~k2] == -2).all()
    assert f["summary"]["traj_counts"] == o2["summary"]["traj_counts"]
    # a table long enough for the staged kernel (k_pool takes its work from the filter's accept list too), all pairs
    specs = W.static_obstacles(77, 12, 12)
    obs = abi.make_obstacles(specs)
    ctx.set_obstacles(obs, len(specs))
    b3 = W.synth_bc(78, 0, 20_011)
    ctx.set_input_filter()
    try:
        g = ctx.generate_check(b3, MIN_T)
        assert ctx.last_kernel_name().startswith("k_pool")
    finally:
        ctx.set_input_filter(on=False)
    f3 = oracle.input_feasibility(b3)
    np.testing.assert_array_equal(g["input_feasibility"], f3)
    k3 = f3 == 0
    o3 = oracle.generate_check(np.ascontiguousarray(b3[:, k3]), obs, len(specs), MIN_T, nthreads=8)
    np.testing.assert_array_equal(g["verdict_matrix"][k3], o3["verdict_matrix"])
    np.testing.assert_array_equal(g["verdict"][k3], o3["verdict"])
    assert (g["verdict"][~k3] == abi.NOT_CHECKED).all()
    assert g["summary"]["pair_counts"] == o3["summary"]["pair_counts"]
    # the best candidate is named by its index in the caller's array, whatever order the accept list had
    want_best = np.flatnonzero(k3)[o3["summary"]["best_index"]] if o3["summary"]["best_index"] >= 0 else -1
    assert g["summary"]["best_index"] == want_best


def test_fast_fma_flag_is_opt_in_and_close(ctx, oracle):
    """RQCD_F_FAST_FMA runs the kernels' FMA-contracted build: costs and coefficients within 1e-12 of the exact build, verdicts
    equal on (nearly) every Monte Carlo pair -- the count is printed, a handful in 1e6 would be legitimate -- while the
    default call stays bit-exact against the oracle."""
    specs = W.static_obstacles(2, 32, 32)
    obs = abi.make_obstacles(specs)
    bc = W.synth_bc(1, 0, 1 << 15)
    ctx.set_obstacles(obs, len(specs))
    exact = ctx.generate_check(bc, MIN_T)
    fma = ctx.generate_check(bc, MIN_T, flags=abi.F_FAST_FMA)
    o = oracle.generate_check(bc, obs, len(specs), MIN_T, nthreads=8)
    np.testing.assert_array_equal(exact["verdict_matrix"], o["verdict_matrix"])
    changed = int((fma["verdict_matrix"] != exact["verdict_matrix"]).sum())
    print(f"fast fma: {changed} of {exact['verdict_matrix'].size} pair verdicts changed")
    assert changed <= 1e-5 * exact["verdict_matrix"].size
    assert_close(fma["cost"], exact["cost"], tol=1e-12, what="cost, contracted build")
    # the other shapes run too: pairwise (k_check), early exit over a short table (k_check), moving table (k_pool)
    u = W.synth_bc(8, 0, 1 << 14)
    sph = np.empty((4, 1 << 14))
    sph[0:3] = u[3:6]
    sph[3] = 0.1 + (u[6] + 4.0) / 8.0 * 1.4
    b14 = np.ascontiguousarray(bc[:, :1 << 14])
    pe = ctx.generate_check_pairwise(b14, sph, MIN_T)
    ctx.set_obstacles(abi.make_obstacles(W.forest_obstacles()), 5)
    c2 = W.c2_candidates(100)
    fe = ctx.generate_check(c2, MIN_T, flags=abi.F_EARLY_EXIT)
    ff = ctx.generate_check(c2, MIN_T, flags=abi.F_EARLY_EXIT | abi.F_FAST_FMA)
    assert int((fe["verdict"] != ff["verdict"]).sum()) <= 2
    assert pe["summary"]["pairs_checked"] == 1 << 14


def test_performance_eval_end_to_end_on_the_device_at_full_size(ctx, golden_counts):
    """test/PerformanceEval.cpp:117-182 as ONE call: the product package draws the driver's 15 571 728 trials (no oracle), the
    input filter keeps exactly 10^7 of them -- the last one drawn among them, as in the driver's loop -- and their verdict
    counts are the reference's checked-in 9 600 380 / 399 612 / 8 (data/PerformanceEval.txt:23-25)."""
    bc, sph = W.c1_trials(golden_counts["c1_trials"])
    ctx.set_input_filter()
    try:
        r = ctx.generate_check_pairwise(bc, sph, MIN_T)
    finally:
        ctx.set_input_filter(on=False)
    accepted = r["input_feasibility"] == 0
    assert int(accepted.sum()) == golden_counts["c1_n"] and accepted[-1]
    assert r["summary"]["pairs_checked"] == golden_counts["c1_n"]
    assert r["summary"]["traj_counts"][:3] == [golden_counts["c1_no_collision"], golden_counts["c1_collision"],
                                              golden_counts["c1_indeterminable"]]
    assert (r["verdict"][~accepted] == abi.NOT_CHECKED).all() and (r["verdict"][accepted] <= 2).all()
