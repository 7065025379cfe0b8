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


def _feas_section_both_ways(rng, n):
    """One CheckInputFeasibilitySection (RapidTrajectoryGenerator.cpp:74-147) on random axes and intervals, restated twice
    in numpy: literally (roots taken, feas_frame_sel's expressions) and on squares (feas_frame_sq's). Returns
    (literal result, squares result, open)."""
    fminA, fmaxA, wmaxA, minT = 5.0, 30.0, 20.0, 0.002
    grav = np.array([0.0, 0.0, -9.81])
    Tf = rng.uniform(0.2, 4.0, n)
    a0 = rng.uniform(-4, 4, (3, n))
    al = rng.uniform(-4, 4, (3, n)) * 60 / Tf ** 3 * rng.choice([0.0, 0.3, 1.0], size=(3, n))
    be = rng.uniform(-4, 4, (3, n)) * 24 / Tf ** 2 * rng.choice([0.0, 0.3, 1.0], size=(3, n))
    ga = rng.uniform(-4, 4, (3, n)) * 6 / Tf * rng.choice([0.0, 0.3, 1.0], size=(3, n))
    lo = rng.uniform(0, 1, n) * rng.choice([0.0, 1.0], size=n)
    t1 = lo * Tf
    t2 = t1 + (Tf - t1) * rng.choice([1.0, 0.5, 0.25, 1e-3, 1e-4], size=n)

    def acc(k, t):  # SingleAxisTrajectory.hpp:56-58
        return a0[k] + ga[k] * t + (1 / 2.0) * be[k] * t * t + (1 / 6.0) * al[k] * t * t * t

    def jerk(k, t):  # :54
        return ga[k] + be[k] * t + (1 / 2.0) * al[k] * t * t

    with np.errstate(all="ignore"):
        def thrust_sq(t):
            e = [acc(k, t) - grav[k] for k in range(3)]
            return e[0] * e[0] + e[1] * e[1] + e[2] * e[2]

        d1, d2 = thrust_sq(t1), thrust_sq(t2)
        fmin_sq = np.zeros(n)
        fmax_sq = np.zeros(n)
        jmax_sq = np.zeros(n)
        axis_high = np.zeros(n, bool)
        for k in range(3):
            has = al[k] != 0.0
            det = be[k] * be[k] - 2 * ga[k] * al[k]
            sa = np.where(has, al[k], 1.0)
            root = np.sqrt(np.where(det < 0, 0.0, det))
            pk0 = np.where(has, np.where(det < 0, 0.0, (-be[k] + root) / sa), np.where(be[k] != 0, -ga[k] / np.where(be[k] != 0, be[k], 1.0), 0.0))
            pk1 = np.where(has, np.where(det < 0, 0.0, (-be[k] - root) / sa), 0.0)
            x1, x2 = acc(k, t1), acc(k, t2)
            amin, amax = np.minimum(x1, x2), np.maximum(x1, x2)
            for pk in (pk0, pk1):  # SingleAxisTrajectory.cpp:160-171
                inside = ~(pk <= t1) & ~(pk >= t2)
                ap = acc(k, pk)
                amin = np.where(inside, np.minimum(amin, ap), amin)
                amax = np.where(inside, np.maximum(amax, ap), amax)
            v1, v2 = amin - grav[k], amax - grav[k]
            axis_high |= np.maximum(v1 * v1, v2 * v2) > fmaxA * fmaxA
            m = np.minimum(np.abs(v1), np.abs(v2))
            fmin_sq += np.where(v1 * v2 < 0, 0.0, m * m)
            M = np.maximum(np.abs(v1), np.abs(v2))
            fmax_sq += M * M
            j = np.maximum(jerk(k, t1) ** 2, jerk(k, t2) ** 2)  # :182-197
            tmax = -be[k] / sa
            jm = np.maximum(jerk(k, tmax) ** 2, j)
            jmax_sq += np.where(has & (tmax > t1) & (tmax < t2), jm, j)
        big = fmin_sq > 1e-6
        SPLIT, OK, IND, HIGH, LOW = 100, 0, 1, 2, 3

        def pick(split, f_gt_max, F_lt_min, e_lt_min, e_gt_max):
            r = np.where(split, SPLIT, OK)
            r = np.where(f_gt_max, HIGH, r)
            r = np.where(F_lt_min, LOW, r)
            r = np.where(axis_high, HIGH, r)
            r = np.where(e_lt_min, LOW, r)
            r = np.where(e_gt_max, HIGH, r)
            return np.where(t2 - t1 < minT, IND, r)

        # literal: the roots themselves
        f1, f2, fmin, fmax = np.sqrt(d1), np.sqrt(d2), np.sqrt(fmin_sq), np.sqrt(fmax_sq)
        wb = np.where(big, np.sqrt(jmax_sq / np.where(big, fmin_sq, 1.0)), 1.7976931348623157e308)
        lit = pick((fmin < fminA) | (fmax > fmaxA) | (wb > wmaxA), fmin > fmaxA, fmax < fminA,
                   np.minimum(f1, f2) < fminA, np.maximum(f1, f2) > fmaxA)
        # on squares
        a2, b2, w2 = fminA * fminA, fmaxA * fmaxA, wmaxA * wmaxA
        lo_min, hi_min, lo_max, hi_max = a2 * (1 - 1e-15), a2 * (1 + 1e-15), b2 * (1 - 1e-15), b2 * (1 + 1e-15)
        P = w2 * fmin_sq
        plo, phi = P * (1 - 4e-15), P * (1 + 4e-15)

        def und(x):
            return (~(x < lo_min) & ~(x > hi_min)) | (~(x < lo_max) & ~(x > hi_max))

        opened = und(d1) | und(d2) | und(fmin_sq) | und(fmax_sq) | (big & ~(jmax_sq < plo) & ~(jmax_sq > phi))
        w_gt = np.where(big, jmax_sq > phi, 1.7976931348623157e308 > wmaxA)
        sq = pick((fmin_sq < lo_min) | (fmax_sq > hi_max) | w_gt, fmin_sq > hi_max, fmax_sq < lo_min,
                  (d1 < lo_min) | (d2 < lo_min), (d1 > hi_max) | (d2 > hi_max))
    return lit, sq, opened


def test_feasibility_section_on_squares_equals_the_literal_section():
    """feas_frame_sq against feas_frame_sel, both restated in numpy, on 3e6 random sections (whole windows, halves, slivers
    below minT, axes with vanishing alpha / beta / gamma): wherever the squares claim certainty the results agree, and
    they claim it almost always."""
    lit, sq, opened = _feas_section_both_ways(np.random.default_rng(10), 3_000_000)
    assert opened.mean() < 1e-6
    assert np.array_equal(lit[~opened], sq[~opened])
    assert len(np.unique(lit)) == 5  # feasible, indeterminable, thrust high, thrust low and "split" all occur
