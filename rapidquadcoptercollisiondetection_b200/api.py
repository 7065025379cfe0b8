"""Thin Python handle over the C ABI (include/rqcd.h) for tests/ and bench.py.

Every method is one C-ABI call; no arithmetic of the path happens in Python and there is no fallback:
constructing a Context without librqcd.so or without a CUDA device raises.

Host mode takes/returns numpy arrays (the library stages them); device mode takes raw device pointers
(ints, e.g. torch.Tensor.data_ptr()) and only enqueues work on the context's stream.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi


class RqcdError(RuntimeError):
    def __init__(self, status: int, text: str, detail: str = ""):
        super().__init__(f"rqcd status {status}: {text}" + (f" [{detail}]" if detail else ""))
        self.status = status


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def bc_soa(bc, n=None, goal_mask=abi.GOAL_ALL, shared_rows=0, group_size=0, first=0, stride=None):
    """rqcd_bc_soa over a (19, stride) numpy array or a raw pointer. shared_rows: iterable of row numbers or a bit mask."""
    if not isinstance(shared_rows, int):
        mask = 0
        for r in shared_rows:
            mask |= 1 << int(r)
        shared_rows = mask
    if stride is None:
        stride = bc.shape[1]
    return abi.BcSoa(_ptr(bc), stride, goal_mask, shared_rows, group_size, first)


def shared_row_mask(rows) -> int:
    m = 0
    for r in rows:
        m |= 1 << int(r)
    return m


class Context:
    """rqcd_ctx: one GPU, one stream."""

    def __init__(self, device: int = 0, _handle=None):
        self.lib = abi.load()
        self._owned = _handle is None
        if _handle is None:
            h = C.c_void_p()
            st = self.lib.rqcd_create(C.byref(h), device)
            if st != abi.OK:
                raise RqcdError(st, self.lib.rqcd_strerror(st).decode())
        else:
            h = C.c_void_p(_handle)  # a context owned by an rqcd_group
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            if self._owned:
                self.lib.rqcd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _chk(self, st, allow=()):
        if st != abi.OK and st not in allow:
            raise RqcdError(st, self.lib.rqcd_strerror(st).decode(), self.lib.rqcd_last_error(self.h).decode())
        return st

    # ---- context ----
    def device_info(self):
        sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
        name = C.create_string_buffer(256)
        self._chk(self.lib.rqcd_device_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi), name, 256))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "name": name.value.decode()}

    def set_stream(self, cuda_stream: int | None):
        self._chk(self.lib.rqcd_set_stream(self.h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def get_stream(self) -> int:
        return int(self.lib.rqcd_get_stream(self.h) or 0)

    def synchronize(self):
        self._chk(self.lib.rqcd_synchronize(self.h))

    def launch_count(self) -> int:
        return int(self.lib.rqcd_launch_count(self.h))

    def last_kernel_name(self) -> str:
        return (self.lib.rqcd_last_kernel_name(self.h) or b"").decode()

    def fragile_reasons(self):
        """rqcd_fragile_reasons: marked pairs of this context by the first decision that was not pinned (32 counters)."""
        out = np.zeros(32, dtype=np.uint64)
        self._chk(self.lib.rqcd_fragile_reasons(self.h, _ptr(out)))
        return out

    def set_input_filter(self, gravity=(0.0, 0.0, -9.81), fmin=5.0, fmax=30.0, wmax=20.0, min_t=0.002, on=True):
        """rqcd_set_input_filter: generate_check / generate_check_pairwise then skip (verdict 255) the candidates that are
        not InputFeasible, like PerformanceEval.cpp:163-173; on=False switches it off."""
        if not on:
            self._chk(self.lib.rqcd_set_input_filter(self.h, None))
            self._filter = False
            return
        lim = abi.InputLimits()
        lim.gravity[:] = [float(x) for x in gravity]
        lim.fmin, lim.fmax, lim.wmax, lim.min_time_section = float(fmin), float(fmax), float(wmax), float(min_t)
        self._chk(self.lib.rqcd_set_input_filter(self.h, C.byref(lim)))
        self._filter = True

    def read_summary(self) -> dict:
        s = abi.Summary()
        self._chk(self.lib.rqcd_read_summary(self.h, C.byref(s)), allow=(abi.ERR_DEPTH_CAP,))
        return s.as_dict()

    def set_obstacles(self, obstacles, m: int | None = None):
        """obstacles: ctypes array of abi.Obstacle (see oracle.binding.make_obstacles / abi.make_obstacles)."""
        if m is None:
            m = len(obstacles)
        self._chk(self.lib.rqcd_set_obstacles(self.h, obstacles, m))
        self.m = m

    def num_obstacles(self) -> int:
        return int(self.lib.rqcd_num_obstacles(self.h))

    # ---- host-mode calls (numpy in / numpy out) ----
    def generate(self, bc, goal_mask=abi.GOAL_ALL, n=None, shared_rows=0, group_size=0, first=0):
        bc = np.ascontiguousarray(bc, dtype=np.float64)
        if n is None:
            n = bc.shape[1] - first
        cost = np.empty(n)
        coeffs = np.empty((abi.COEFF_ROWS, n))
        params = np.empty((abi.PARAM_ROWS, n))
        out = abi.Out(cost=_ptr(cost), coeffs=_ptr(coeffs), coeff_stride=n, params=_ptr(params), param_stride=n)
        inp = bc_soa(bc, n, goal_mask, shared_rows, group_size, first)
        self._chk(self.lib.rqcd_generate(self.h, C.byref(inp), n, 0, C.byref(out)))
        return {"cost": cost, "coeffs": coeffs, "abg": params}

    @staticmethod
    def _split_fragile(res):
        """RQCD_F_FRAGILE: bit 7 of a matrix entry marks a fragile pair; hand it out as its own array."""
        if "verdict_matrix" in res:
            res["fragile_matrix"] = (res["verdict_matrix"] >> 7).astype(np.uint8)
            res["verdict_matrix"] &= 0x7F

    def _host_out(self, n, m, matrix, coeffs, first=True, fragile=False):
        res = {"verdict": np.full(n, 255, dtype=np.uint8)}
        if getattr(self, "_filter", False):
            res["input_feasibility"] = np.full(n, 255, dtype=np.uint8)
        if fragile:
            res["fragile"] = np.full(n, 255, dtype=np.uint8)
        if first:
            res["first_obstacle"] = np.full(n, -2, dtype=np.int32)
        if coeffs:
            res["cost"] = np.empty(n)
            res["coeffs"] = np.empty((abi.COEFF_ROWS, n))
        if matrix:
            res["verdict_matrix"] = np.full((n, max(m, 1)), 255, dtype=np.uint8)
        summ = abi.Summary()
        out = abi.Out(verdict=_ptr(res["verdict"]), first_obstacle=_ptr(res.get("first_obstacle")),
                      verdict_matrix=_ptr(res.get("verdict_matrix")), cost=_ptr(res.get("cost")),
                      coeffs=_ptr(res.get("coeffs")), coeff_stride=n, summary=C.pointer(summ),
                      fragile=_ptr(res.get("fragile")), input_feasibility=_ptr(res.get("input_feasibility")))
        return res, out, summ

    def generate_check(self, bc, min_t, goal_mask=abi.GOAL_ALL, flags=0, index_base=0, matrix=True, n=None, shared_rows=0,
                       group_size=0, first=0, fragile=False):
        """bc: (19, stride) array. With shared_rows (bit mask or row numbers) pass n explicitly: the shared rows hold one
        value per group of group_size candidates, the per-candidate rows n values from element `first`."""
        bc = np.ascontiguousarray(bc, dtype=np.float64)
        if n is None:
            n = bc.shape[1] - first
        m = self.num_obstacles()
        if fragile:
            flags |= abi.F_FRAGILE
        res, out, summ = self._host_out(n, m, matrix and not (flags & abi.F_EARLY_EXIT), True, fragile=fragile)
        inp = bc_soa(bc, n, goal_mask, shared_rows, group_size, first)
        st = self._chk(self.lib.rqcd_generate_check(self.h, C.byref(inp), n, min_t, flags, index_base, C.byref(out)),
                       allow=(abi.ERR_DEPTH_CAP,))
        if fragile:
            self._split_fragile(res)
        res["summary"] = summ.as_dict()
        res["status"] = st
        return res

    def check(self, coeffs, t_end, min_t, t_start=None, cost=None, flags=0, index_base=0, matrix=True, fragile=False):
        coeffs = np.ascontiguousarray(coeffs, dtype=np.float64)
        n = coeffs.shape[1]
        m = self.num_obstacles()
        t_end = np.ascontiguousarray(t_end, dtype=np.float64)
        if t_start is not None:
            t_start = np.ascontiguousarray(t_start, dtype=np.float64)
        if cost is not None:
            cost = np.ascontiguousarray(cost, dtype=np.float64)
        if fragile:
            flags |= abi.F_FRAGILE
        res, out, summ = self._host_out(n, m, matrix and not (flags & abi.F_EARLY_EXIT), False, fragile=fragile)
        inp = abi.CoeffSoa(_ptr(coeffs), n, _ptr(t_start), _ptr(t_end))
        st = self._chk(self.lib.rqcd_check(self.h, C.byref(inp), _ptr(cost), n, min_t, flags, index_base, C.byref(out)),
                       allow=(abi.ERR_DEPTH_CAP,))
        if fragile:
            self._split_fragile(res)
        res["summary"] = summ.as_dict()
        res["status"] = st
        return res

    def generate_check_pairwise(self, bc, spheres, min_t, goal_mask=abi.GOAL_ALL, index_base=0, fragile=False):
        bc = np.ascontiguousarray(bc, dtype=np.float64)
        spheres = np.ascontiguousarray(spheres, dtype=np.float64)
        n = bc.shape[1]
        res, out, summ = self._host_out(n, 1, False, True, first=False, fragile=fragile)
        inp = abi.BcSoa(_ptr(bc), n, goal_mask, 0)
        st = self._chk(self.lib.rqcd_generate_check_pairwise(self.h, C.byref(inp), _ptr(spheres), n, n, min_t,
                                                             abi.F_FRAGILE if fragile else 0,
                                                             index_base, C.byref(out)), allow=(abi.ERR_DEPTH_CAP,))
        res["summary"] = summ.as_dict()
        res["status"] = st
        return res

    def check_planes(self, coeffs, t_end, planes, t_start=None):
        """planes: (p, 6) array of {point, normal}. Returns verdict (n,), verdict_matrix (n, p)."""
        coeffs = np.ascontiguousarray(coeffs, dtype=np.float64)
        planes = np.ascontiguousarray(planes, dtype=np.float64).reshape(-1, 6)
        n, p = coeffs.shape[1], planes.shape[0]
        t_end = np.ascontiguousarray(t_end, dtype=np.float64)
        if t_start is not None:
            t_start = np.ascontiguousarray(t_start, dtype=np.float64)
        verdict = np.full(n, 255, dtype=np.uint8)
        matrix = np.full((n, max(p, 1)), 255, dtype=np.uint8)
        inp = abi.CoeffSoa(_ptr(coeffs), n, _ptr(t_start), _ptr(t_end))
        self._chk(self.lib.rqcd_check_planes(self.h, C.byref(inp), n, _ptr(planes), p, 0, _ptr(verdict), _ptr(matrix)))
        return verdict, matrix

    def generate_check_planes(self, bc, planes, goal_mask=abi.GOAL_ALL):
        bc = np.ascontiguousarray(bc, dtype=np.float64)
        planes = np.ascontiguousarray(planes, dtype=np.float64).reshape(-1, 6)
        n, p = bc.shape[1], planes.shape[0]
        verdict = np.full(n, 255, dtype=np.uint8)
        matrix = np.full((n, max(p, 1)), 255, dtype=np.uint8)
        inp = abi.BcSoa(_ptr(bc), n, goal_mask, 0)
        self._chk(self.lib.rqcd_generate_check_planes(self.h, C.byref(inp), n, _ptr(planes), p, 0, _ptr(verdict), _ptr(matrix)))
        return verdict, matrix

    def input_feasibility(self, bc, goal_mask=abi.GOAL_ALL, gravity=(0.0, 0.0, -9.81), fmin=5.0, fmax=30.0, wmax=20.0,
                          min_t=0.002, group=None, n=None, shared_rows=0, group_size=0, first=0):
        bc = np.ascontiguousarray(bc, dtype=np.float64)
        if n is None:
            n = bc.shape[1] - first
        g = np.asarray(gravity, dtype=np.float64)
        if group is not None:
            group = np.ascontiguousarray(group, dtype=np.int64)
        res = np.full(n, 255, dtype=np.uint8)
        inp = bc_soa(bc, n, goal_mask, shared_rows, group_size, first)
        self._chk(self.lib.rqcd_check_input_feasibility(self.h, C.byref(inp), n, _ptr(g), fmin, fmax, wmax, min_t, _ptr(group), 0,
                                                  _ptr(res)))
        return res

    def plan_best(self, bc, min_t, max_cost=float("inf"), input_limits=(5.0, 30.0, 20.0), gravity=(0.0, 0.0, -9.81),
                  planes=None, goal_mask=abi.GOAL_ALL, index_base=0):
        """rqcd_plan_best: cost prune -> input feasibility -> plane boundaries -> obstacles -> cheapest survivor."""
        bc = np.ascontiguousarray(bc, dtype=np.float64)
        n = bc.shape[1]
        prm = abi.PlanParams()
        prm.max_cost = max_cost
        prm.check_input = 1 if input_limits is not None else 0
        prm.gravity[:] = [float(x) for x in gravity]
        if input_limits is not None:
            prm.fmin, prm.fmax, prm.wmax = (float(x) for x in input_limits)
        if planes is not None:
            planes = np.ascontiguousarray(planes, dtype=np.float64).reshape(-1, 6)
            prm.planes = planes.ctypes.data
            prm.n_planes = planes.shape[0]
        prm.min_time_section = min_t
        stage = np.full(n, 255, dtype=np.uint8)
        summ = abi.Summary()
        inp = abi.BcSoa(_ptr(bc), n, goal_mask, 0)
        st = self._chk(self.lib.rqcd_plan_best(self.h, C.byref(inp), n, C.byref(prm), 0, index_base, _ptr(stage), C.byref(summ)),
                       allow=(abi.ERR_DEPTH_CAP,))
        return {"stage": stage, "summary": summ.as_dict(), "status": st}

    def solve_cubic(self, coef):
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        n = coef.shape[1]
        roots = np.full((3, n), np.nan)
        cnt = np.zeros(n, dtype=np.int32)
        self._chk(self.lib.rqcd_solve_cubic(self.h, _ptr(coef), n, 0, _ptr(roots), _ptr(cnt)))
        return roots, cnt

    def solve_quartic(self, coef):
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        n = coef.shape[1]
        roots = np.full((4, n), np.nan)
        cnt = np.zeros(n, dtype=np.int32)
        self._chk(self.lib.rqcd_solve_quartic(self.h, _ptr(coef), n, 0, _ptr(roots), _ptr(cnt)))
        return roots, cnt

    def solve_conditioning(self, coef):
        """rqcd_solve_conditioning: (cond, branch_fragile) per problem; coef has 3 (cubic) or 4 (quartic) rows."""
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        degree, n = coef.shape
        cond = np.zeros(n)
        flag = np.zeros(n, dtype=np.uint8)
        self._chk(self.lib.rqcd_solve_conditioning(self.h, degree, _ptr(coef), n, 0, _ptr(cond), _ptr(flag)))
        return cond, flag

    # ---- device-mode calls (raw device pointers; asynchronous on the context's stream) ----
    def synth_bc(self, seed, index_base, n, d_bc: int, stride: int):
        self._chk(self.lib.rqcd_synth_bc(self.h, seed, index_base, n, C.c_void_p(d_bc), stride))

    def generate_check_dev(self, d_bc: int, stride: int, n: int, min_t: float, goal_mask=abi.GOAL_ALL, flags=0,
                           index_base=0, d_verdict=0, d_first=0, d_matrix=0, d_cost=0, d_coeffs=0, coeff_stride=0,
                           want_summary=False, shared_rows=0, group_size=0, first=0):
        summ = abi.Summary()
        out = abi.Out(verdict=d_verdict or None, first_obstacle=d_first or None, verdict_matrix=d_matrix or None,
                      cost=d_cost or None, coeffs=d_coeffs or None, coeff_stride=coeff_stride,
                      summary=C.pointer(summ) if want_summary else None)
        inp = abi.BcSoa(C.c_void_p(d_bc), stride, goal_mask, shared_rows, group_size, first)
        st = self._chk(self.lib.rqcd_generate_check(self.h, C.byref(inp), n, min_t, flags | abi.F_DEVICE_PTRS,
                                                    index_base, C.byref(out)), allow=(abi.ERR_DEPTH_CAP,))
        return summ.as_dict() if want_summary else st

    def generate_check_pairwise_dev(self, d_bc: int, stride: int, d_spheres: int, sphere_stride: int, n: int,
                                    min_t: float, goal_mask=abi.GOAL_ALL, index_base=0, d_verdict=0, d_cost=0,
                                    want_summary=False, d_input_feas=0):
        summ = abi.Summary()
        out = abi.Out(verdict=d_verdict or None, cost=d_cost or None,
                      summary=C.pointer(summ) if want_summary else None, input_feasibility=d_input_feas or None)
        inp = abi.BcSoa(C.c_void_p(d_bc), stride, goal_mask, 0)
        st = self._chk(self.lib.rqcd_generate_check_pairwise(self.h, C.byref(inp), C.c_void_p(d_spheres), sphere_stride,
                                                             n, min_t, abi.F_DEVICE_PTRS, index_base, C.byref(out)),
                       allow=(abi.ERR_DEPTH_CAP,))
        return summ.as_dict() if want_summary else st

    # ---- one rank per process: the in-library NCCL exchange ----
    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = C.create_string_buffer(bytes(unique_id), abi.COMM_ID_BYTES)
        self._chk(self.lib.rqcd_comm_init(self.h, buf, rank, world))

    def comm_finalize(self):
        self._chk(self.lib.rqcd_comm_finalize(self.h))

    def comm_world(self) -> int:
        return int(self.lib.rqcd_comm_world(self.h))

    def exchange_summaries(self, d_record: int = 0):
        """Enqueue all-gather + merge of the last launch's summary on the context's stream (no host sync)."""
        self._chk(self.lib.rqcd_exchange_summaries(self.h, C.c_void_p(d_record) if d_record else None))

    def read_merged_summary(self) -> dict:
        s = abi.Summary()
        self._chk(self.lib.rqcd_read_merged_summary(self.h, C.byref(s)), allow=(abi.ERR_DEPTH_CAP,))
        return s.as_dict()

    def measure_fp64_peak(self, millis=200):
        a, b = C.c_double(), C.c_double()
        self._chk(self.lib.rqcd_measure_fp64_peak(self.h, millis, C.byref(a), C.byref(b)))
        return {"dfma_flops": a.value, "dmul_dadd_flops": b.value}


def comm_unique_id() -> bytes:
    """rqcd_comm_unique_id: the 128-byte NCCL id one rank creates and the host's transport hands to the others."""
    lib = abi.load()
    buf = C.create_string_buffer(abi.COMM_ID_BYTES)
    st = lib.rqcd_comm_unique_id(buf)
    if st != abi.OK:
        raise RqcdError(st, lib.rqcd_strerror(st).decode())
    return buf.raw


class Group:
    """rqcd_group: several GPUs in one process; candidates sharded by index, summaries merged on the first device."""

    def __init__(self, devices):
        self.lib = abi.load()
        devs = (C.c_int32 * len(devices))(*devices)
        h = C.c_void_p()
        st = self.lib.rqcd_group_create(C.byref(h), devs, len(devices))
        if st != abi.OK:
            raise RqcdError(st, self.lib.rqcd_strerror(st).decode())
        self.h = h
        self.devices = list(devices)

    def close(self):
        if getattr(self, "h", None):
            self.lib.rqcd_group_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, st, allow=()):
        if st != abi.OK and st not in allow:
            raise RqcdError(st, self.lib.rqcd_strerror(st).decode(), self.lib.rqcd_group_last_error(self.h).decode())
        return st

    def size(self) -> int:
        return int(self.lib.rqcd_group_size(self.h))

    def peer_slots(self) -> bool:
        return bool(self.lib.rqcd_group_peer_slots(self.h))

    def ctx(self, k: int) -> Context:
        return Context(self.devices[k], _handle=self.lib.rqcd_group_ctx(self.h, k))

    def launch_count(self) -> int:
        return sum(int(self.lib.rqcd_launch_count(self.lib.rqcd_group_ctx(self.h, k))) for k in range(self.size()))

    def set_obstacles(self, obstacles, m=None):
        if m is None:
            m = len(obstacles)
        self._chk(self.lib.rqcd_group_set_obstacles(self.h, obstacles, m))
        self.m = m

    def generate_check(self, bc, min_t, goal_mask=abi.GOAL_ALL, flags=0, matrix=True, n=None, shared_rows=0, group_size=0,
                       first=0):
        """Host buffers in, host buffers out, the whole group (rqcd_group_generate_check)."""
        bc = np.ascontiguousarray(bc, dtype=np.float64)
        if n is None:
            n = bc.shape[1] - first
        m = self.m
        res = {"verdict": np.full(n, 255, dtype=np.uint8), "first_obstacle": np.full(n, -2, dtype=np.int32),
               "cost": np.empty(n), "coeffs": np.empty((abi.COEFF_ROWS, n))}
        if matrix and not (flags & abi.F_EARLY_EXIT):
            res["verdict_matrix"] = np.full((n, max(m, 1)), 255, dtype=np.uint8)
        summ = abi.Summary()
        out = abi.Out(verdict=_ptr(res["verdict"]), first_obstacle=_ptr(res["first_obstacle"]),
                      verdict_matrix=_ptr(res.get("verdict_matrix")), cost=_ptr(res["cost"]), coeffs=_ptr(res["coeffs"]),
                      coeff_stride=n, summary=C.pointer(summ))
        inp = bc_soa(bc, n, goal_mask, shared_rows, group_size, first)
        st = self._chk(self.lib.rqcd_group_generate_check(self.h, C.byref(inp), n, min_t, flags, C.byref(out)),
                       allow=(abi.ERR_DEPTH_CAP,))
        res["summary"] = summ.as_dict()
        res["status"] = st
        return res

    def generate_check_resident(self, d_bc, strides, ns, min_t, goal_mask=abi.GOAL_ALL, flags=0, index_base=0,
                                d_verdict=None, d_first=None, d_cost=None, want_summary=True):
        """Shards resident on their devices: per-shard lists of device pointers / strides / counts."""
        G = self.size()
        ins = (abi.BcSoa * G)()
        outs = (abi.Out * G)()
        cnt = (C.c_int64 * G)(*ns)
        for k in range(G):
            ins[k] = abi.BcSoa(C.c_void_p(d_bc[k]), strides[k], goal_mask, 0, 0, 0)
            outs[k] = abi.Out(verdict=d_verdict[k] if d_verdict else None, first_obstacle=d_first[k] if d_first else None,
                              cost=d_cost[k] if d_cost else None)
        summ = abi.Summary()
        st = self._chk(self.lib.rqcd_group_generate_check_resident(self.h, ins, cnt, min_t, flags, index_base, outs,
                                                                    C.byref(summ) if want_summary else None),
                       allow=(abi.ERR_DEPTH_CAP,))
        return summ.as_dict() if want_summary else st

    def read_summary(self) -> dict:
        s = abi.Summary()
        self._chk(self.lib.rqcd_group_read_summary(self.h, C.byref(s)), allow=(abi.ERR_DEPTH_CAP,))
        return s.as_dict()


def merge_summaries(summaries) -> dict:
    """rqcd_merge_summaries over a list of summary dicts (the multi-GPU exchange step)."""
    lib = abi.load()
    arr = (abi.Summary * max(len(summaries), 1))()
    for dst, s in zip(arr, summaries):
        dst.traj_counts[:] = s["traj_counts"]
        dst.pair_counts[:] = s["pair_counts"]
        dst.pairs_checked = s["pairs_checked"]
        dst.best_cost = s["best_cost"]
        dst.best_index = s["best_index"]
    out = abi.Summary()
    st = lib.rqcd_merge_summaries(arr, len(summaries), C.byref(out))
    if st != abi.OK:
        raise RqcdError(st, lib.rqcd_strerror(st).decode())
    return out.as_dict()
