"""CPU checks (numpy, no GPU) of the two certified filters the kernels use in place of literal reference arithmetic
(rqcd_device.cuh: frame() plane-side tests, ends_inside() sphere test). They restate the device formulas in numpy
and confirm, on random and on deliberately near-degenerate inputs, that whenever the filter claims certainty its
answer equals the literal evaluation in the reference's operation order -- i.e. that the margins really cover the
rounding errors. (The device code itself is exercised by tests/test_gpu_parity.py.)"""
import numpy as np


def _literal_side(c, t, pp, pn):
    """(GetValue(t) - point) . normal with Trajectory.hpp:79-81's operation order (c[k][dim], k = 0 is t^5)."""
    X = np.empty((3, t.size))
    for d in range(3):
        X[d] = (c[0, d] * t * t * t * t * t + c[1, d] * t * t * t * t + c[2, d] * t * t * t + c[3, d] * t * t
                + c[4, d] * t + c[5, d])
    w = X - pp
    return w[0] * pn[0] + w[1] * pn[1] + w[2] * pn[2]


def test_plane_side_filter_margin_covers_both_evaluations():
    rng = np.random.default_rng(7)
    n = 2_000_000
    Tf = rng.uniform(0.2, 4.0, n)
    c = np.empty((6, 3, n))
    for k in range(6):
        c[k] = rng.uniform(-4, 4, (3, n)) / Tf ** (5 - k) * rng.choice([0.1, 1.0, 1.0, 10.0, 1e3], size=n)
    t = rng.uniform(0, 1, n) * Tf
    pn = rng.normal(size=(3, n))
    pn /= np.sqrt((pn ** 2).sum(0)).astype(np.float32).astype(np.float64)  # Vec3::GetUnitVector's float norm
    # half of the planes pass (almost) through X(t): the value to be signed is ~0, +-1e-14, +-1e-11 relative
    X = sum(c[k] * t ** (5 - k) for k in range(6))
    shift = rng.choice([0.0, 1e-14, 1e-11], size=n) * rng.normal(size=n)
    pp = np.where(rng.random(n) < 0.5, X - pn * shift, rng.uniform(-5, 5, (3, n)))

    d_ref = _literal_side(c, t, pp, pn)
    # device form: pc[k] as critical_times builds it, g_k = pc[k] / (5 - k), g_5 = (c_5 - point) . n, Horner
    g = []
    for k in range(5):
        acc = np.zeros(n)
        for dim in range(3):
            acc = acc + pn[dim] * ((5.0 - k) * c[k, dim])
        g.append(acc * [0.2, 0.25, 1.0 / 3.0, 0.5, 1.0][k])
    e = c[5] - pp
    g5 = e[0] * pn[0] + e[1] * pn[1] + e[2] * pn[2]
    d = g[0]
    for k in range(1, 5):
        d = d * t + g[k]  # two roundings per step here, one (FMA) on the device: the device is at least as accurate
    d = d * t + g5
    amax = np.zeros(n)
    for k in range(6):
        amax = amax * Tf + np.abs(c[k]).sum(0)  # side_filter_bound
    tau = 1e-12 * np.abs(pn).sum(0) * (amax + np.abs(pp).sum(0))
    certain = np.abs(d) > tau
    assert 0.4 < certain.mean() < 0.9  # both regimes are populated
    assert not np.any(certain & ((d <= 0) != (d_ref <= 0)))
    # and the margin is not marginal: the two evaluations differ by a small multiple of u (A + P), tau is ~9000 u (A + P)
    scale = amax + np.abs(pp).sum(0)
    assert np.max(np.abs(d - d_ref) / scale) < 64 * 2.0 ** -53


def test_sphere_inside_filter_equals_the_rounded_root():
    rng = np.random.default_rng(8)
    n = 2_000_000
    r = 10.0 ** rng.uniform(-3, 3, n)
    # points at distance r (1 + k ulp-ish) from the centre along random directions, plus generic ones
    u = rng.normal(size=(3, n))
    u /= np.sqrt((u ** 2).sum(0))
    rel = rng.choice([0.0, 1e-16, 2e-16, 5e-16, 1e-15, 3e-15, 1e-12, 0.3], size=n) * rng.choice([-1.0, 1.0], size=n)
    w = u * (r * (1.0 + rel))
    d = w[0] * w[0] + w[1] * w[1] + w[2] * w[2]
    literal = np.sqrt(d) <= r  # Sphere.hpp:70-72
    r2 = r * r
    lo, hi = r2 * (1.0 - 1e-15), r2 * (1.0 + 1e-15)
    inside, outside = d < lo, d > hi
    assert inside.sum() > 0.1 * n and outside.sum() > 0.1 * n and (~inside & ~outside).sum() > 0.1 * n
    assert np.all(literal[inside]) and not np.any(literal[outside])


def test_feasibility_section_comparisons_on_squares():
    """feas_frame_sq (rqcd_device.cuh): every root of a CheckInputFeasibilitySection is only compared with a limit
    (RapidTrajectoryGenerator.cpp:85-147); the kernel decides on the squares when they are apart by more than the
    roundings. Here: whenever the squares claim certainty, the literal comparison of the rounded root agrees."""
    rng = np.random.default_rng(9)
    n = 2_000_000
    A = np.concatenate([np.full(n // 2, 5.0), 10.0 ** rng.uniform(-3, 3, n - n // 2)])  # fminA = 5 and generic limits
    rel = rng.choice([0.0, 1e-16, 2e-16, 4e-16, 8e-16, 1.2e-15, 2e-15, 1e-12, 0.3], size=n) * rng.choice([-1.0, 1.0], size=n)
    root = A * (1.0 + rel)
    d = root * root * (1.0 + rng.choice([0.0, 1.1e-16, -1.1e-16, 2.2e-16], size=n))  # the squared quantity, around A^2
    a2 = A * A
    lo, hi = a2 * (1.0 - 1e-15), a2 * (1.0 + 1e-15)
    f = np.sqrt(d)
    below, above = d < lo, d > hi
    assert below.sum() > 0.1 * n and above.sum() > 0.1 * n and (~below & ~above).sum() > 0.1 * n
    assert np.all(f[below] < A[below]) and np.all(f[above] > A[above])
    # the rate bound: sqrt(j / fminSqr) > wmaxA decided by j against wmaxA^2 fminSqr (1 -+ 4e-15)
    W = np.concatenate([np.full(n // 2, 20.0), 10.0 ** rng.uniform(-2, 3, n - n // 2)])
    fm = 10.0 ** rng.uniform(-6, 4, n)
    rel = rng.choice([0.0, 1e-16, 4e-16, 1e-15, 2e-15, 3e-15, 5e-15, 1e-12, 0.3], size=n) * rng.choice([-1.0, 1.0], size=n)
    j = (W * W) * fm * (1.0 + rel)
    P = (W * W) * fm
    plo, phi = P * (1.0 - 4e-15), P * (1.0 + 4e-15)
    wq = np.sqrt(j / fm)
    below, above = j < plo, j > phi
    assert below.sum() > 0.1 * n and above.sum() > 0.1 * n and (~below & ~above).sum() > 0.1 * n
    assert np.all(wq[below] < W[below]) and np.all(wq[above] > W[above])
