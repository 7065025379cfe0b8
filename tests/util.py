"""Shared helpers of the parity tests."""
import numpy as np

from rapidquadcoptercollisiondetection_b200 import abi, workloads as W

REL_TOL = 1e-9  # BASELINE.json: coefficients, costs and crossing times within 1e-9 relative (FP64 mode)


def obstacles_from(vectors, prefix):
    specs = W.arrays_to_specs(vectors, prefix)
    return abi.make_obstacles(specs), len(specs)


def assert_close(a, b, tol=REL_TOL, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, what
    both_nan = np.isnan(a) & np.isnan(b)
    same_inf = np.isinf(a) & (a == b)
    ok = both_nan | same_inf
    with np.errstate(invalid="ignore"):
        err = np.abs(a - b)
        scale = np.maximum(np.abs(a), np.abs(b))
        ok |= err <= tol * scale
    bad = np.argwhere(~ok)
    assert bad.size == 0, f"{what}: {len(bad)} entries differ beyond {tol} relative, first {bad[:3].tolist()}: " \
                          f"{a[tuple(bad[0])]} vs {b[tuple(bad[0])]}"


def _polyval_monic(coef, x):
    """x^k + coef[0] x^(k-1) + ... + coef[k-1] (the monic form Quartic::solveP3 / solve_quartic take)."""
    y = np.ones_like(x)
    for row in coef:
        y = y * x + row
    return y


def assert_roots_match(coef, roots_a, cnt_a, roots_b, cnt_b, cond, fragile, what=""):
    """(a = CUDA path, b = reference/oracle.)  No quota: every problem is held to a bound of its own.

    Everything in src/quartic.cpp is IEEE arithmetic the CUDA path repeats operation by operation, except acos / cos /
    pow(x, 1/3), where glibc and the device polynomials differ by an ulp or two. rqcd_solve_conditioning measures what
    that ulp is worth per problem: `cond` = root movement per unit of libm error, `fragile` = a branch of the solver
    (root count!) sits within that movement of its threshold. So
      * root counts must be equal wherever the problem is not branch-fragile;
      * every root must agree within 1e-9 relative (the bar of BASELINE.json) OR within cond * 2^-49 absolute -- eight
        ulps of libm difference amplified by the problem's own conditioning;
      * on branch-fragile problems with equal counts the CUDA root must still be as good a root of the polynomial as the
        reference's (residual test)."""
    n = cnt_a.shape[0]
    fragile = np.asarray(fragile) != 0
    diff = np.flatnonzero((cnt_a != cnt_b) & ~fragile)
    assert diff.size == 0, (f"{what}: root counts differ on problems that are not branch-fragile, at {diff[:8]}: "
                            f"{cnt_a[diff[:8]]} vs {cnt_b[diff[:8]]}; coef {coef[:, diff[:4]].T.tolist()}")
    keep = cnt_a == cnt_b
    rows = roots_a.shape[0]
    with np.errstate(invalid="ignore", over="ignore"):
        scale = np.maximum(1.0, np.nanmax(np.abs(np.where(np.isfinite(coef), coef, 0.0)), axis=0))
        allow = np.where(np.isfinite(cond), cond, np.inf) * 2.0 ** -49
        for k in range(rows):
            a, b = roots_a[k], roots_b[k]
            live = (cnt_a > k) if rows == 4 else np.ones(n, dtype=bool)  # cubic: x[1], x[2] hold Re/Im when 1 root
            live &= keep
            both_nan = np.isnan(a) & np.isnan(b)
            same_inf = np.isinf(a) & (a == b)
            err = np.abs(a - b)
            rel = err / np.maximum(np.maximum(np.abs(a), np.abs(b)), scale)
            ok = both_nan | same_inf | (rel <= REL_TOL) | (err <= allow) | ~live
            bad = np.flatnonzero(~ok & ~fragile)
            assert bad.size == 0, (f"{what}: root {k} differs beyond 1e-9 relative and beyond its conditioning at {bad[:5]}: "
                                   f"{a[bad[:5]]} vs {b[bad[:5]]}, cond {cond[bad[:5]]}")
            fr = np.flatnonzero(~ok & fragile)
            if fr.size and (rows == 4 or k == 0):  # real roots: ours must be as good a root as the reference's
                res_a = np.abs(_polyval_monic(coef[:, fr], a[fr]))
                res_b = np.abs(_polyval_monic(coef[:, fr], b[fr]))
                floor = 1e-12 * scale[fr] ** rows
                assert np.all(res_a <= 16 * res_b + floor), \
                    f"{what}: root {k} at {fr[:5]} is a worse root than the reference's: {res_a[:5]} vs {res_b[:5]}"
    return {"problems": int(n), "branch_fragile": int(fragile.sum()), "count_differs": int((cnt_a != cnt_b).sum()),
            "beyond_1e-9": int(sum(((np.abs(roots_a[k] - roots_b[k]) / np.maximum(np.maximum(np.abs(roots_a[k]), np.abs(roots_b[k])), scale)
                                     > REL_TOL) & keep & ((cnt_a > k) if rows == 4 else True)).sum() for k in range(rows)))}


def degenerate_case():
    """Inputs the reference neither rejects nor handles specially (it returns whatever IEEE arithmetic gives)."""
    n = 4096
    bc = W.synth_bc(71, 0, n)
    rng = np.random.default_rng(71)
    bc[abi.BC_TF, 0:64] = 0.0                       # division by zero in the closed forms -> NaN coefficients
    bc[abi.BC_TF, 64:128] = 1e-4                    # whole window below minTimeSection
    bc[abi.BC_TF, 128:192] = 1e3                    # long flights: deep recursions
    bc[abi.BC_PF:abi.BC_PF + 3, 192:256] = 1e300    # overflow to inf in the coefficients
    bc[abi.BC_V0, 256:320] = np.nan                 # NaN initial state
    bc[abi.BC_A0:abi.BC_A0 + 3, 320:384] = 0.0      # start exactly on a sphere's surface, at rest:
    bc[abi.BC_V0:abi.BC_V0 + 3, 320:384] = 0.0
    bc[abi.BC_P0, 320:384] = 1.0                    #   |p0 - c| == r for the unit sphere at the origin
    bc[abi.BC_P0 + 1:abi.BC_P0 + 3, 320:384] = 0.0
    bc[:, 384:448] = 0.0                            # everything zero incl. Tf
    bc[abi.BC_TF, 448:512] = rng.uniform(1e-3, 3e-3, 64)  # windows around minTimeSection itself
    specs = [
        {"type": "sphere", "center": [0.0, 0.0, 0.0], "radius": 1.0},
        {"type": "sphere", "center": [0.0, 0.0, 0.0], "radius": 1e6},   # everything starts inside
        {"type": "sphere", "center": [1e9, 0.0, 0.0], "radius": 1e-3},  # nothing comes near
        {"type": "prism", "center": [0.5, 0.0, 0.0], "side": [1.0, 1.0, 1.0], "quat": [1.0, 0, 0, 0]},  # p0 = (1,0,0) on a face
        {"type": "prism", "center": [0.0, 0.0, 2.0], "side": [1e-6, 8.0, 8.0], "quat": [0.7071067811865476, 0, 0.7071067811865476, 0]},
        {"type": "prism", "center": [0.0, 0.0, 0.0], "side": [0.3, 0.2, 0.1], "quat": [0.0, 0.0, 0.0, 0.0]},  # zero quaternion
    ]
    return bc, specs


def grazing_case(n=20_000, seed=2024, log_a=(1.0, 8.0)):
    """Trajectories that come within rounding noise of an obstacle's surface at a critical time:
    x(t) = A (t - t*)^2 + E (t - t*)^5 + 1 + delta, y(t) = b (t - t*), z = 0 against the unit sphere at the origin and
    the face x = 1 of two prisms (+ 17 far spheres). A quintic term is needed: the reference's solver divides by zero
    on velocity polynomials of degree < 3. A = 10^U(log_a): from ~1e3 on the reference's closed-form roots are off by
    more than delta / A (cancellation in q cos(.) - a/3), and WHICH side of the plane the computed critical point
    lands on is decided by the last bits of libm's acos / cos / pow."""
    rng = np.random.default_rng(seed)
    t_end = rng.uniform(0.5, 1.5, n)
    ts = rng.uniform(0.1, 0.9, n) * t_end
    A = 10.0 ** rng.uniform(log_a[0], log_a[1], n)
    E = rng.choice([-1.0, 1.0], n) * rng.uniform(0.5, 2.0, n)
    delta = rng.choice([-1.0, 1.0], n) * 10.0 ** rng.uniform(-13, -6, n)
    bq = rng.uniform(-1e-2, 1e-2, n)
    coeffs = np.zeros((18, n))
    coeffs[0] = E
    coeffs[3] = -5 * E * ts
    coeffs[6] = 10 * E * ts**2
    coeffs[9] = -10 * E * ts**3 + A
    coeffs[12] = 5 * E * ts**4 - 2 * A * ts
    coeffs[15] = -E * ts**5 + A * ts * ts + 1.0 + delta
    coeffs[13] = bq
    coeffs[16] = -bq * ts
    specs = [
        {"type": "sphere", "center": [0.0, 0.0, 0.0], "radius": 1.0},
        {"type": "prism", "center": [0.0, 0.0, 0.0], "side": [2.0, 2.0, 2.0], "quat": [1.0, 0, 0, 0]},
        {"type": "prism", "center": [0.0, 0.0, 0.0], "side": [2.0, 1.0, 3.0], "quat": [0.9238795325112867, 0.3826834323650898, 0, 0]},
    ] + [{"type": "sphere", "center": [3.0 + k, 1.0, 0.0], "radius": 0.5} for k in range(17)]
    return coeffs, t_end, specs
