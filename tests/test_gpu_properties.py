"""GPU: parity at scale and size-independent properties at BASELINE.json's full single-GPU size (C3: 2^20 primitives
x 64 obstacles), where the CPU oracle covers a sample and the rest is pinned by invariants of the domain:
pair checks are independent, so obstacle order, early exit, sharding and the fused/unfused split must not change
any verdict."""
import numpy as np
import pytest

from rapidquadcoptercollisiondetection_b200 import abi, workloads as W
from rapidquadcoptercollisiondetection_b200.api import merge_summaries

pytestmark = pytest.mark.gpu
MIN_T = 0.002
N_FULL = 1 << 20


@pytest.fixture(scope="module")
def c3(ctx):
    specs = W.static_obstacles(2, 32, 32)
    bc = W.synth_bc(1, 0, N_FULL)
    ctx.set_obstacles(abi.make_obstacles(specs), 64)
    full = ctx.generate_check(bc, MIN_T)  # all pairs: verdict matrix 2^20 x 64
    return specs, bc, full


def test_sample_of_c3_against_oracle(ctx, oracle, c3):
    """1.7e7 pairs (every 4th candidate of the first 2^20) through the oracle on all host threads."""
    specs, bc, full = c3
    sel = np.arange(0, N_FULL, 4)
    o = oracle.generate_check(np.ascontiguousarray(bc[:, sel]), abi.make_obstacles(specs), 64, MIN_T, nthreads=16)
    np.testing.assert_array_equal(full["verdict_matrix"][sel], o["verdict_matrix"])
    np.testing.assert_array_equal(full["verdict"][sel], o["verdict"])
    np.testing.assert_array_equal(full["first_obstacle"][sel], o["first_obstacle"])
    np.testing.assert_allclose(full["cost"][sel], o["cost"], rtol=1e-9)


def test_matrix_summary_and_first_hit_are_consistent(c3):
    _, _, full = c3
    mtx, s = full["verdict_matrix"], full["summary"]
    assert s["pairs_checked"] == N_FULL * 64
    assert s["pair_counts"][:3] == [int((mtx == v).sum()) for v in range(3)] and s["pair_counts"][3] == 0
    hit = mtx != 0
    first = np.where(hit.any(axis=1), hit.argmax(axis=1), -1)
    np.testing.assert_array_equal(full["first_obstacle"], first)
    want = np.where(first >= 0, mtx[np.arange(N_FULL), np.maximum(first, 0)], 0)
    np.testing.assert_array_equal(full["verdict"], want)
    assert s["traj_counts"][:3] == [int((want == v).sum()) for v in range(3)]
    free = np.flatnonzero(want == 0)
    if free.size:
        best = free[np.argmin(full["cost"][free])]
        assert s["best_index"] == best and s["best_cost"] == full["cost"][best]
    else:
        assert s["best_index"] == -1


def test_early_exit_does_not_change_verdicts(ctx, c3):
    specs, bc, full = c3
    r = ctx.generate_check(bc, MIN_T, flags=abi.F_EARLY_EXIT)
    np.testing.assert_array_equal(r["verdict"], full["verdict"])
    np.testing.assert_array_equal(r["first_obstacle"], full["first_obstacle"])
    executed = np.where(full["first_obstacle"] >= 0, full["first_obstacle"] + 1, 64).sum()
    assert r["summary"]["pairs_checked"] == int(executed)


def test_obstacle_order_only_permutes_the_matrix(ctx, c3):
    specs, bc, full = c3
    perm = np.random.default_rng(7).permutation(64)
    ctx.set_obstacles(abi.make_obstacles([specs[j] for j in perm]), 64)
    n = 1 << 18
    r = ctx.generate_check(bc[:, :n].copy(), MIN_T)
    np.testing.assert_array_equal(r["verdict_matrix"], full["verdict_matrix"][:n][:, perm])
    ctx.set_obstacles(abi.make_obstacles(specs), 64)


def test_ragged_shards_merge_to_the_single_shot_summary(ctx, c3):
    """The multi-GPU exchange on one GPU: three uneven candidate ranges, summaries merged by rqcd_merge_summaries."""
    specs, bc, full = c3
    cuts = [0, 300_001, 777_777, N_FULL]
    parts = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        r = ctx.generate_check(np.ascontiguousarray(bc[:, lo:hi]), MIN_T, index_base=lo, matrix=False)
        np.testing.assert_array_equal(r["verdict"], full["verdict"][lo:hi])
        parts.append(r["summary"])
    assert merge_summaries(parts) == full["summary"]


def test_unfused_generate_then_check_equals_fused(ctx, c3):
    specs, bc, full = c3
    n = 1 << 18
    g = ctx.generate(bc[:, :n].copy())
    np.testing.assert_array_equal(g["coeffs"], full["coeffs"][:, :n])   # same kernel arithmetic: bit for bit
    np.testing.assert_array_equal(g["cost"], full["cost"][:n])
    r = ctx.check(g["coeffs"], bc[abi.BC_TF, :n], MIN_T, cost=g["cost"])
    np.testing.assert_array_equal(r["verdict_matrix"], full["verdict_matrix"][:n])


def test_pairwise_equals_one_obstacle_sets(ctx, oracle):
    """PerformanceEval shape against the table path: sphere i for candidate i == a one-sphere set per candidate."""
    n = 4096
    bc = W.synth_bc(9, 0, n)
    rng = np.random.default_rng(9)
    sph = np.vstack([rng.uniform(-4, 4, (3, n)), rng.uniform(0.1, 1.5, (1, n))])
    r = ctx.generate_check_pairwise(bc, sph, MIN_T)
    o = oracle.generate_check_pairwise(bc, sph, MIN_T)
    np.testing.assert_array_equal(r["verdict"], o["verdict"])
    for i in range(0, n, 512):
        ctx.set_obstacles(abi.make_obstacles([{"type": "sphere", "center": sph[:3, i].tolist(), "radius": float(sph[3, i])}]), 1)
        one = ctx.generate_check(bc[:, i:i + 1].copy(), MIN_T)
        assert one["verdict"][0] == r["verdict"][i]


def test_moving_set_sample_of_c4_against_oracle(ctx, oracle):
    """BASELINE config 4 (moving, non-rotating obstacles): 2^16 primitives x 32 moving obstacles = 2.1e6 pairs."""
    n = 1 << 16
    specs = W.moving_obstacles(4, 32, lambda b: ctx.generate(b)["coeffs"])
    obs = abi.make_obstacles(specs)
    ctx.set_obstacles(obs, 32)
    bc = W.synth_bc(3, 0, n)
    r = ctx.generate_check(bc, MIN_T)
    o = oracle.generate_check(bc, obs, 32, MIN_T, nthreads=16)
    np.testing.assert_array_equal(r["verdict_matrix"], o["verdict_matrix"])
    np.testing.assert_array_equal(r["first_obstacle"], o["first_obstacle"])
    assert r["summary"] == {**o["summary"], "best_cost": r["summary"]["best_cost"]}
    assert (o["verdict_matrix"] == 1).any() and (o["verdict_matrix"] == 2).any()


def test_large_table_sample_of_c5_against_oracle(ctx, oracle):
    """BASELINE config 5's obstacle set (128 spheres + 128 prisms: a 51 KB table per CTA): 2^14 primitives taken from
    the middle of the 2^26 index space x 256 obstacles = 4.2e6 pairs."""
    n = 1 << 14
    specs = W.static_obstacles(6, 128, 128)
    obs = abi.make_obstacles(specs)
    ctx.set_obstacles(obs, 256)
    base = (1 << 25) + 12345
    bc = W.synth_bc(5, base, n)
    r = ctx.generate_check(bc, MIN_T, index_base=base)
    o = oracle.generate_check(bc, obs, 256, MIN_T, index_base=base, nthreads=16)
    np.testing.assert_array_equal(r["verdict_matrix"], o["verdict_matrix"])
    np.testing.assert_array_equal(r["verdict"], o["verdict"])
    assert r["summary"]["pair_counts"] == o["summary"]["pair_counts"]
    assert r["summary"]["best_index"] == o["summary"]["best_index"]


def test_device_synth_equals_host_synth(ctx):
    """rqcd_synth_bc (device) and workloads.synth_bc (numpy) are the same counter-based generator, so any candidate of
    the 2^26-candidate configuration can be regenerated on the host for the oracle."""
    import torch
    n, base = 4096, (1 << 26) - 4096
    d = torch.empty((abi.BC_ROWS, n), dtype=torch.float64, device="cuda")
    ctx.synth_bc(5, base, n, d.data_ptr(), n)
    ctx.synchronize()
    np.testing.assert_array_equal(d.cpu().numpy(), W.synth_bc(5, base, n))
