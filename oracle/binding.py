"""TEST INFRASTRUCTURE ONLY: numpy/ctypes access to the CPU checkers.

 * ``Oracle()``    -> oracle/liboracle.so, the plain-C restatement (oracle.c)
 * ``Reference()`` -> oracle/_ref/libref.so, the unmodified reference sources + ref_shim.cpp

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from rapidquadcoptercollisiondetection_b200 import abi  # noqa: E402  (struct mirrors of include/rqcd.h)

ORACLE_PATH = os.path.join(HERE, "liboracle.so")
REF_PATH = os.path.join(HERE, "_ref", "libref.so")


class Stats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in (
        "checks", "sections", "tangent_sections", "test_points", "trig", "cardano", "root_pairs",
        "cubic_fallback", "max_depth", "sphere_checks", "prism_checks", "skip_first", "skip_child",
        "child_tangent_sections")]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_obstacles(specs):
    """See abi.make_obstacles (the struct builder lives with the struct mirrors)."""
    specs = [dict(s, motion=np.asarray(s["motion"]).reshape(18).tolist()) if "motion" in s else s for s in specs]
    return abi.make_obstacles(specs)


class _CpuImpl:
    prefix = ""
    has_stats = False

    def __init__(self, path):
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        self.lib = C.CDLL(path)
        self.path = path

    def _f(self, name, restype=None):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype = restype
        return fn

    # ---- solver ----
    def solve_cubic(self, coef):
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        n = coef.shape[1]
        roots = np.full((3, n), np.nan)
        cnt = np.zeros(n, dtype=np.int32)
        self._f("solve_cubic_batch")(_ptr(coef), C.c_int64(n), _ptr(roots), _ptr(cnt))
        return roots, cnt

    def solve_quartic(self, coef):
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        n = coef.shape[1]
        roots = np.full((4, n), np.nan)
        cnt = np.zeros(n, dtype=np.int32)
        self._f("solve_quartic_batch")(_ptr(coef), C.c_int64(n), _ptr(roots), _ptr(cnt))
        return roots, cnt

    # ---- generation ----
    def generate(self, bc, goal_mask=abi.GOAL_ALL):
        bc = np.ascontiguousarray(bc, dtype=np.float64)
        n = bc.shape[1]
        cost = np.empty(n)
        coeffs = np.empty((18, n))
        abg = np.empty((9, n))
        out = abi.Out(cost=_ptr(cost), coeffs=_ptr(coeffs), coeff_stride=n)
        inp = abi.BcSoa(_ptr(bc), n, goal_mask, 0)
        self._f("generate")(C.byref(inp), C.c_int64(n), C.byref(out), _ptr(abg))
        return {"cost": cost, "coeffs": coeffs, "abg": abg}

    def _outputs(self, n, m, matrix, coeffs=True):
        res = {
            "verdict": np.zeros(n, dtype=np.uint8),
            "first_obstacle": np.zeros(n, dtype=np.int32),
            "cost": np.empty(n),
        }
        if coeffs:
            res["coeffs"] = np.empty((18, n))
        if matrix:
            res["verdict_matrix"] = np.zeros((n, max(m, 1)), dtype=np.uint8)
        summ = abi.Summary()
        out = abi.Out(
            verdict=_ptr(res["verdict"]), first_obstacle=_ptr(res["first_obstacle"]),
            verdict_matrix=_ptr(res.get("verdict_matrix")), cost=_ptr(res["cost"]),
            coeffs=_ptr(res.get("coeffs")), coeff_stride=n, summary=C.pointer(summ))
        return res, out, summ

    def generate_check(self, bc, obstacles, m, min_t, goal_mask=abi.GOAL_ALL, flags=0, index_base=0,
                       matrix=True, nthreads=1, stats=False):
        bc = np.ascontiguousarray(bc, dtype=np.float64)
        n = bc.shape[1]
        res, out, summ = self._outputs(n, m, matrix and not (flags & abi.F_EARLY_EXIT))
        inp = abi.BcSoa(_ptr(bc), n, goal_mask, 0)
        args = [C.byref(inp), C.c_int64(n), obstacles, C.c_int32(m), C.c_double(min_t), C.c_uint32(flags),
                C.c_int64(index_base), C.byref(out), C.c_int(nthreads)]
        st = Stats()
        if self.has_stats:
            args.append(C.byref(st) if stats else None)
        self._f("generate_check")(*args)
        res["summary"] = summ.as_dict()
        if stats and self.has_stats:
            res["stats"] = st.as_dict()
        return res

    def check(self, coeffs, t_end, obstacles, m, min_t, t_start=None, cost=None, flags=0, index_base=0,
              matrix=True, nthreads=1, stats=False):
        coeffs = np.ascontiguousarray(coeffs, dtype=np.float64)
        n = coeffs.shape[1]
        t_end = np.ascontiguousarray(t_end, dtype=np.float64)
        if t_start is not None:
            t_start = np.ascontiguousarray(t_start, dtype=np.float64)
        if cost is not None:
            cost = np.ascontiguousarray(cost, dtype=np.float64)
        res, out, summ = self._outputs(n, m, matrix and not (flags & abi.F_EARLY_EXIT), coeffs=False)
        inp = abi.CoeffSoa(_ptr(coeffs), n, _ptr(t_start), _ptr(t_end))
        args = [C.byref(inp), _ptr(cost), C.c_int64(n), obstacles, C.c_int32(m), C.c_double(min_t),
                C.c_uint32(flags), C.c_int64(index_base), C.byref(out), C.c_int(nthreads)]
        st = Stats()
        if self.has_stats:
            args.append(C.byref(st) if stats else None)
        self._f("check")(*args)
        res.pop("cost")
        res["summary"] = summ.as_dict()
        if stats and self.has_stats:
            res["stats"] = st.as_dict()
        return res

    def generate_check_pairwise(self, bc, spheres, min_t, goal_mask=abi.GOAL_ALL, index_base=0, nthreads=1,
                                stats=False):
        bc = np.ascontiguousarray(bc, dtype=np.float64)
        spheres = np.ascontiguousarray(spheres, dtype=np.float64)
        n = bc.shape[1]
        res, out, summ = self._outputs(n, 1, False)
        inp = abi.BcSoa(_ptr(bc), n, goal_mask, 0)
        args = [C.byref(inp), _ptr(spheres), C.c_int64(n), C.c_int64(n), C.c_double(min_t), C.c_int64(index_base),
                C.byref(out), C.c_int(nthreads)]
        st = Stats()
        if self.has_stats:
            args.append(C.byref(st) if stats else None)
        self._f("generate_check_pairwise")(*args)
        res["summary"] = summ.as_dict()
        if stats and self.has_stats:
            res["stats"] = st.as_dict()
        return res

    def check_one(self, coef18, t0, t1, obstacle, min_t):
        coef18 = np.ascontiguousarray(coef18, dtype=np.float64).reshape(18)
        args = [_ptr(coef18), C.c_double(t0), C.c_double(t1), C.byref(obstacle), C.c_double(min_t)]
        if self.has_stats:
            args.append(None)
        return int(self._f("check_one", C.c_int)(*args))

    def input_feasibility(self, bc, goal_mask=abi.GOAL_ALL, gravity=(0.0, 0.0, -9.81), fmin=5.0, fmax=30.0,
                          wmax=20.0, min_t=0.002, group=None):
        bc = np.ascontiguousarray(bc, dtype=np.float64)
        n = bc.shape[1]
        g = np.asarray(gravity, dtype=np.float64)
        res = np.zeros(n, dtype=np.uint8)
        if group is not None:
            group = np.ascontiguousarray(group, dtype=np.int64)
        inp = abi.BcSoa(_ptr(bc), n, goal_mask, 0)
        self._f("input_feasibility")(C.byref(inp), C.c_int64(n), _ptr(g), C.c_double(fmin), C.c_double(fmax),
                                     C.c_double(wmax), C.c_double(min_t), _ptr(group), _ptr(res))
        return res

    def feas_counters(self, reset=True):
        """(sections, axis evaluations, full sections) of input_feasibility calls since the last reset (oracle only)."""
        out = (C.c_uint64 * 3)()
        self._f("feas_counters")(out, C.c_int(1 if reset else 0))
        return [int(x) for x in out]

    def rotation_matrix(self, quat):
        q = np.asarray(quat, dtype=np.float64)
        R = np.empty(9)
        self._f("rotation_matrix")(_ptr(q), _ptr(R))
        return R

    def traj_value(self, coef18, t):
        c = np.ascontiguousarray(coef18, dtype=np.float64).reshape(18)
        out = np.empty(3)
        self._f("traj_value")(_ptr(c), C.c_double(t), _ptr(out))
        return out

    def forest_obstacles(self):
        arr = (abi.Obstacle * 5)()
        self._f("forest_obstacles")(arr)
        return arr


class Oracle(_CpuImpl):
    """oracle/oracle.c -- the C restatement ("port")."""
    prefix = "orc_"
    has_stats = True
    kind = "port"

    def __init__(self):
        super().__init__(ORACLE_PATH)

    def synth_bc(self, seed, index_base, n):
        bc = np.empty((abi.BC_ROWS, n))
        self._f("synth_bc")(C.c_uint64(seed), C.c_int64(index_base), C.c_int64(n), _ptr(bc), C.c_int64(n))
        return bc

    def c1_inputs(self, n_accept, filter=True):
        bc = np.empty((abi.BC_ROWS, n_accept))
        sph = np.empty((4, n_accept))
        trials = self._f("c1_inputs", C.c_int64)(C.c_int64(n_accept), C.c_int(1 if filter else 0), _ptr(bc),
                                                 C.c_int64(n_accept), _ptr(sph), C.c_int64(n_accept))
        return bc, sph, int(trials)

    def c2_inputs(self, n_batches, per_batch=100):
        n = n_batches * per_batch
        bc = np.empty((abi.BC_ROWS, n))
        self._f("c2_inputs")(C.c_int64(n_batches), C.c_int64(per_batch), _ptr(bc), C.c_int64(n))
        return bc


    def check_planes(self, coeffs, t_end, planes, t_start=None):
        coeffs = np.ascontiguousarray(coeffs, dtype=np.float64)
        planes = np.ascontiguousarray(planes, dtype=np.float64).reshape(-1, 6)
        n, p = coeffs.shape[1], planes.shape[0]
        t_end = np.ascontiguousarray(t_end, dtype=np.float64)
        if t_start is not None:
            t_start = np.ascontiguousarray(t_start, dtype=np.float64)
        verdict = np.zeros(n, dtype=np.uint8)
        matrix = np.zeros((n, max(p, 1)), dtype=np.uint8)
        inp = abi.CoeffSoa(_ptr(coeffs), n, _ptr(t_start), _ptr(t_end))
        self._f("check_planes")(C.byref(inp), C.c_int64(n), _ptr(planes), C.c_int32(p), _ptr(verdict), _ptr(matrix))
        return verdict, matrix


class Reference(_CpuImpl):
    """oracle/_ref/libref.so -- the unmodified reference classes behind ref_shim.cpp."""
    prefix = "ref_"
    has_stats = False
    kind = "reference"

    def __init__(self):
        super().__init__(REF_PATH)

    def c1_run(self, n_accept, filter=True):
        bc = np.empty((abi.BC_ROWS, n_accept))
        sph = np.empty((4, n_accept))
        verdict = np.zeros(n_accept, dtype=np.uint8)
        trials = self._f("c1_run", C.c_int64)(C.c_int64(n_accept), C.c_int(1 if filter else 0), _ptr(bc),
                                              C.c_int64(n_accept), _ptr(sph), C.c_int64(n_accept), _ptr(verdict))
        return bc, sph, verdict, int(trials)

    def c2_run(self, n_batches, per_batch=100):
        n = n_batches * per_batch
        bc = np.empty((abi.BC_ROWS, n))
        feas = np.zeros(n, dtype=np.uint8)
        verdict = np.zeros(n, dtype=np.uint8)
        pairs = C.c_int64(0)
        self._f("c2_run")(C.c_int64(n_batches), C.c_int64(per_batch), _ptr(bc), C.c_int64(n), _ptr(feas),
                          _ptr(verdict), C.byref(pairs))
        return bc, feas, verdict, int(pairs.value)

    def hardware_threads(self):
        return int(self._f("hardware_threads", C.c_int)())


def have_reference() -> bool:
    return os.path.exists(REF_PATH)
