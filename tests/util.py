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


def assert_roots_match(coef, roots_a, cnt_a, roots_b, cnt_b, what=""):
    """Root counts exact; roots equal within 1e-9 relative to the magnitude of the polynomial's own scale
    (closed-form roots of nearly repeated roots lose digits to cancellation in BOTH implementations, so
    the comparison is relative to max(|root|, coefficient scale))."""
    np.testing.assert_array_equal(cnt_a, cnt_b, err_msg=what + " root counts")
    rows = roots_a.shape[0]
    with np.errstate(invalid="ignore", over="ignore"):
        scale = np.maximum(1.0, np.nanmax(np.abs(np.where(np.isfinite(coef), coef, 0.0)), axis=0))
        for k in range(rows):
            a, b = roots_a[k], roots_b[k]
            live = cnt_a > k if rows == 4 else np.ones_like(cnt_a, dtype=bool)
            both_nan = np.isnan(a) & np.isnan(b)
            same_inf = np.isinf(a) & (a == b)
            err = np.abs(a - b)
            ok = both_nan | same_inf | (err <= REL_TOL * np.maximum(np.maximum(np.abs(a), np.abs(b)), scale)) | ~live
            bad = np.flatnonzero(~ok)
            assert bad.size == 0, f"{what}: root {k} differs at {bad[:5]}: {a[bad[:5]]} vs {b[bad[:5]]}, coef {coef[:, bad[0]]}"
