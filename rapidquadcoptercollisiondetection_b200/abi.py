"""ctypes mirror of include/rqcd.h and the loader of the CUDA library (librqcd.so).

This package is harness plumbing for tests/ and bench.py: the product is the C-ABI shared library
built from csrc/ (CUDA, sm_100a) and the C++ facade in include/rqcd/.  There is no CPU fallback: if
librqcd.so is missing, load() raises.
"""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.environ.get("RQCD_LIB") or os.path.join(PKG_DIR, "librqcd.so")  # RQCD_LIB: kernel-variant experiments

# rqcd_status
OK, ERR_INVALID_ARG, ERR_CUDA, ERR_NO_DEVICE, ERR_DEPTH_CAP, ERR_NOMEM, ERR_UNSUPPORTED = range(7)
# rqcd_verdict (CollisionChecker.hpp:33-37)
NO_COLLISION, COLLISION, INDETERMINABLE, DEPTH_CAP = range(4)
SPHERE, RECT_PRISM = 0, 1
BC_P0, BC_V0, BC_A0, BC_PF, BC_VF, BC_AF, BC_TF, BC_ROWS = 0, 3, 6, 9, 12, 15, 18, 19
COEFF_ROWS = 18
PARAM_ROWS = 9
GOAL_ALL = 0x1FF
F_DEVICE_PTRS = 0x1
F_EARLY_EXIT = 0x2
F_FRAGILE = 0x4
F_FAST_FMA = 0x8
MAX_DEPTH = 40
COMM_ID_BYTES = 128


def goal_pos(axis: int) -> int:
    return 1 << (3 * axis)


def goal_vel(axis: int) -> int:
    return 1 << (3 * axis + 1)


def goal_acc(axis: int) -> int:
    return 1 << (3 * axis + 2)


class Obstacle(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("moving", C.c_int32),
        ("center", C.c_double * 3),
        ("radius", C.c_double),
        ("side", C.c_double * 3),
        ("quat", C.c_double * 4),
        ("motion", C.c_double * 18),
        ("motion_t_start", C.c_double),
        ("motion_t_end", C.c_double),
    ]


class BcSoa(C.Structure):
    # shared_rows / group_size / first: rows shared by groups of consecutive candidates, see rqcd.h (0, 0, 0 = none)
    _fields_ = [("data", C.c_void_p), ("stride", C.c_int64), ("goal_mask", C.c_uint32), ("shared_rows", C.c_uint32),
                ("group_size", C.c_int64), ("first", C.c_int64)]


class CoeffSoa(C.Structure):
    _fields_ = [("coeffs", C.c_void_p), ("stride", C.c_int64), ("t_start", C.c_void_p), ("t_end", C.c_void_p)]


class Summary(C.Structure):
    _fields_ = [
        ("traj_counts", C.c_uint64 * 4),
        ("pair_counts", C.c_uint64 * 4),
        ("pairs_checked", C.c_uint64),
        ("best_cost", C.c_double),
        ("best_index", C.c_int64),
    ]

    def as_dict(self) -> dict:
        return {
            "traj_counts": list(self.traj_counts),
            "pair_counts": list(self.pair_counts),
            "pairs_checked": int(self.pairs_checked),
            "best_cost": float(self.best_cost),
            "best_index": int(self.best_index),
        }


class PlanParams(C.Structure):
    _fields_ = [
        ("max_cost", C.c_double),
        ("check_input", C.c_int32),
        ("gravity", C.c_double * 3),
        ("fmin", C.c_double),
        ("fmax", C.c_double),
        ("wmax", C.c_double),
        ("planes", C.c_void_p),
        ("n_planes", C.c_int32),
        ("min_time_section", C.c_double),
    ]


class InputLimits(C.Structure):
    _fields_ = [("gravity", C.c_double * 3), ("fmin", C.c_double), ("fmax", C.c_double), ("wmax", C.c_double),
                ("min_time_section", C.c_double)]


NOT_CHECKED = 255


class Out(C.Structure):
    _fields_ = [
        ("verdict", C.c_void_p),
        ("first_obstacle", C.c_void_p),
        ("verdict_matrix", C.c_void_p),
        ("cost", C.c_void_p),
        ("coeffs", C.c_void_p),
        ("coeff_stride", C.c_int64),
        ("summary", C.POINTER(Summary)),
        ("fragile", C.c_void_p),
        ("input_feasibility", C.c_void_p),
        ("params", C.c_void_p),
        ("param_stride", C.c_int64),
    ]


# every symbol include/rqcd.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
SYMBOLS = {
    "rqcd_create": (C.c_int, [C.POINTER(_VP), C.c_int]),
    "rqcd_destroy": (None, [_VP]),
    "rqcd_strerror": (C.c_char_p, [C.c_int]),
    "rqcd_last_error": (C.c_char_p, [_VP]),
    "rqcd_device_info": (C.c_int, [_VP, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.c_size_t]),
    "rqcd_get_stream": (_VP, [_VP]),
    "rqcd_set_stream": (C.c_int, [_VP, _VP]),
    "rqcd_synchronize": (C.c_int, [_VP]),
    "rqcd_read_summary": (C.c_int, [_VP, C.POINTER(Summary)]),
    "rqcd_launch_count": (C.c_uint64, [_VP]),
    "rqcd_last_kernel_name": (C.c_char_p, [_VP]),
    "rqcd_fragile_reasons": (C.c_int, [_VP, _VP]),
    "rqcd_set_input_filter": (C.c_int, [_VP, C.POINTER(InputLimits)]),
    "rqcd_set_obstacles": (C.c_int, [_VP, C.POINTER(Obstacle), C.c_int32]),
    "rqcd_num_obstacles": (C.c_int32, [_VP]),
    "rqcd_generate": (C.c_int, [_VP, C.POINTER(BcSoa), C.c_int64, C.c_uint32, C.POINTER(Out)]),
    "rqcd_generate_check": (C.c_int, [_VP, C.POINTER(BcSoa), C.c_int64, C.c_double, C.c_uint32, C.c_int64, C.POINTER(Out)]),
    "rqcd_check": (C.c_int, [_VP, C.POINTER(CoeffSoa), _VP, C.c_int64, C.c_double, C.c_uint32, C.c_int64, C.POINTER(Out)]),
    "rqcd_generate_check_pairwise": (
        C.c_int,
        [_VP, C.POINTER(BcSoa), _VP, C.c_int64, C.c_int64, C.c_double, C.c_uint32, C.c_int64, C.POINTER(Out)],
    ),
    "rqcd_check_planes": (C.c_int, [_VP, C.POINTER(CoeffSoa), C.c_int64, _VP, C.c_int32, C.c_uint32, _VP, _VP]),
    "rqcd_generate_check_planes": (C.c_int, [_VP, C.POINTER(BcSoa), C.c_int64, _VP, C.c_int32, C.c_uint32, _VP, _VP]),
    "rqcd_check_input_feasibility": (
        C.c_int,
        [_VP, C.POINTER(BcSoa), C.c_int64, _VP, C.c_double, C.c_double, C.c_double, C.c_double, _VP, C.c_uint32, _VP],
    ),
    "rqcd_plan_best": (C.c_int, [_VP, C.POINTER(BcSoa), C.c_int64, C.POINTER(PlanParams), C.c_uint32, C.c_int64, _VP,
                                 C.POINTER(Summary)]),
    "rqcd_solve_cubic": (C.c_int, [_VP, _VP, C.c_int64, C.c_uint32, _VP, _VP]),
    "rqcd_solve_quartic": (C.c_int, [_VP, _VP, C.c_int64, C.c_uint32, _VP, _VP]),
    "rqcd_solve_conditioning": (C.c_int, [_VP, C.c_int32, _VP, C.c_int64, C.c_uint32, _VP, _VP]),
    "rqcd_merge_summaries": (C.c_int, [C.POINTER(Summary), C.c_int32, C.POINTER(Summary)]),
    "rqcd_shard_range": (None, [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "rqcd_group_create": (C.c_int, [C.POINTER(_VP), C.POINTER(C.c_int32), C.c_int32]),
    "rqcd_group_destroy": (None, [_VP]),
    "rqcd_group_size": (C.c_int32, [_VP]),
    "rqcd_group_ctx": (_VP, [_VP, C.c_int32]),
    "rqcd_group_last_error": (C.c_char_p, [_VP]),
    "rqcd_group_peer_slots": (C.c_int, [_VP]),
    "rqcd_group_set_obstacles": (C.c_int, [_VP, C.POINTER(Obstacle), C.c_int32]),
    "rqcd_group_generate_check": (C.c_int, [_VP, C.POINTER(BcSoa), C.c_int64, C.c_double, C.c_uint32, C.POINTER(Out)]),
    "rqcd_group_generate_check_resident": (C.c_int, [_VP, C.POINTER(BcSoa), C.POINTER(C.c_int64), C.c_double, C.c_uint32,
                                                     C.c_int64, C.POINTER(Out), C.POINTER(Summary)]),
    "rqcd_group_read_summary": (C.c_int, [_VP, C.POINTER(Summary)]),
    "rqcd_comm_unique_id": (C.c_int, [_VP]),
    "rqcd_comm_init": (C.c_int, [_VP, _VP, C.c_int32, C.c_int32]),
    "rqcd_comm_finalize": (C.c_int, [_VP]),
    "rqcd_comm_world": (C.c_int32, [_VP]),
    "rqcd_exchange_summaries": (C.c_int, [_VP, _VP]),
    "rqcd_read_merged_summary": (C.c_int, [_VP, C.POINTER(Summary)]),
    "rqcd_synth_bc": (C.c_int, [_VP, C.c_uint64, C.c_int64, C.c_int64, _VP, C.c_int64]),
    "rqcd_measure_fp64_peak": (C.c_int, [_VP, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

def make_obstacles(specs):
    """Build a ctypes array of rqcd_obstacle from dicts:
    {type:'sphere', center, radius} | {type:'prism', center, side, quat}, optionally with
    motion (18,), motion_t_start, motion_t_end for a translating obstacle."""
    arr = (Obstacle * max(len(specs), 1))()
    for o, s in zip(arr, specs):
        o.type = SPHERE if s["type"] == "sphere" else RECT_PRISM
        o.center[:] = [float(x) for x in s["center"]]
        o.radius = float(s.get("radius", 0.0))
        o.side[:] = [float(x) for x in s.get("side", (0.0, 0.0, 0.0))]
        o.quat[:] = [float(x) for x in s.get("quat", (1.0, 0.0, 0.0, 0.0))]
        if "motion" in s:
            o.moving = 1
            o.motion[:] = [float(x) for x in list(s["motion"])]
            o.motion_t_start = float(s["motion_t_start"])
            o.motion_t_end = float(s["motion_t_end"])
    return arr


_lib = None


def load() -> C.CDLL:
    """Load librqcd.so (built in-tree by __graft_entry__.build()).  Fails loudly when absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA library is the only implementation of this path "
            "(no CPU fallback). Build it with `python -c 'import __graft_entry__ as g; g.build()'`."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
