"""GPU: the parts of the C-ABI boundary around the kernels -- the chunked streaming host pipeline (pinned input, progress
word), rows shared by groups of candidates (rqcd_bc_soa::shared_rows), device groups (rqcd_group_*: one process, several
GPUs) and the in-library NCCL exchange (rqcd_comm_*). Each is checked against the plain single-context call and the CPU
oracle on the same seeded inputs; verdicts, first hits, counts and the best candidate bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest

from rapidquadcoptercollisiondetection_b200 import abi, workloads as W
from rapidquadcoptercollisiondetection_b200.api import Context, Group, comm_unique_id, merge_summaries, shared_row_mask
from tests.util import assert_close

pytestmark = pytest.mark.gpu
MIN_T = 0.002
FOREST_SHARED = (0, 1, 2, 3, 4, 5, 6, 7, 8, 12, 13, 14, 15, 16, 17)  # all but goal position (9..11) and Tf (18)


def _pinned(a):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


def _fresh_ctx(env):
    """A context created under a changed environment (the tuning knobs are read by rqcd_create)."""
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return Context(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("pool", [0, 1])
def test_streamed_pinned_input_equals_resident_input(pool, oracle):
    """Page-locked input in several chunks (RQCD_HOST_CHUNK small): the kernel starts after the first chunk and waits on
    the progress word for the rest. n is not a multiple of 256; both check kernels; results equal the pageable
    (single-copy) call and the oracle."""
    specs = W.static_obstacles(21, 12, 12)
    obs = abi.make_obstacles(specs)
    n = 100_003
    bc = W.synth_bc(17, 0, n)
    keep, bc_pin = _pinned(bc)
    with _fresh_ctx({"RQCD_HOST_CHUNK": 4096, "RQCD_POOL": pool}) as c:
        c.set_obstacles(obs, len(specs))
        streamed = c.generate_check(bc_pin, MIN_T, matrix=False)
        name = c.last_kernel_name()
        plain = c.generate_check(bc.copy(), MIN_T, matrix=False)
    assert ("k_pool" in name) == bool(pool)
    o = oracle.generate_check(bc, obs, len(specs), MIN_T, matrix=False, nthreads=8)
    for r in (streamed, plain):
        np.testing.assert_array_equal(r["verdict"], o["verdict"])
        np.testing.assert_array_equal(r["first_obstacle"], o["first_obstacle"])
        assert_close(r["cost"], o["cost"], what="cost")
        assert r["summary"]["pair_counts"] == o["summary"]["pair_counts"]
        assert r["summary"]["best_index"] == o["summary"]["best_index"]
    assert streamed["summary"] == plain["summary"]
    del keep


def test_streamed_pinned_pairwise(oracle):
    n = 70_001
    bc = W.synth_bc(18, 0, n)
    u = W.synth_bc(19, 0, n)
    sph = np.empty((4, n))
    sph[0:3] = u[3:6]
    sph[3] = 0.1 + (u[6] + 4.0) / 8.0 * 1.4
    k1, bc_pin = _pinned(bc)
    k2, sph_pin = _pinned(sph)
    with _fresh_ctx({"RQCD_HOST_CHUNK": 8192}) as c:
        r = c.generate_check_pairwise(bc_pin, sph_pin, MIN_T)
    o = oracle.generate_check_pairwise(bc, sph, MIN_T, nthreads=8)
    np.testing.assert_array_equal(r["verdict"], o["verdict"])
    assert r["summary"]["traj_counts"] == o["summary"]["traj_counts"]
    del k1, k2


def _forest_batch(n, group, seed):
    """A Forest-style batch (ForestPerformanceEval.cpp:175-193): one initial state per `group` candidates, goal position
    and duration per candidate, goal velocity / acceleration zero. Returns (full 19 x n matrix, compact matrix)."""
    ng = (n + group - 1) // group
    per = W.synth_bc(seed, 0, n)
    grp = W.synth_bc(seed + 1, 0, ng)
    compact = np.zeros((abi.BC_ROWS, n))
    compact[9:12] = per[9:12]
    compact[18] = per[18]
    for r in FOREST_SHARED:
        compact[r, :ng] = 0.0 if r >= 12 else grp[r]
    full = compact.copy()
    gi = np.arange(n) // group
    for r in FOREST_SHARED:
        full[r] = compact[r, gi]
    return full, compact


@pytest.mark.parametrize("pinned", [False, True])
def test_shared_rows_equal_the_expanded_matrix(pinned, oracle):
    """rqcd_bc_soa::shared_rows: 4 per-candidate rows + 15 rows with one value per group of 100 give the results of the
    expanded 19 x n matrix, through the host pipeline (pageable and streamed) and with device pointers; a sub-range
    selected with `first` (not on a group boundary) equals the slice of the whole."""
    import torch
    specs = W.forest_obstacles()
    obs = abi.make_obstacles(specs)
    n, G = 50_000 + 37, 100
    full, compact = _forest_batch(n, G, 31)
    mask = shared_row_mask(FOREST_SHARED)
    o = oracle.generate_check(full, obs, len(specs), MIN_T, flags=abi.F_EARLY_EXIT, matrix=False, nthreads=8)
    keep = None
    src = compact
    if pinned:
        keep, src = _pinned(compact)
    with _fresh_ctx({"RQCD_HOST_CHUNK": 4096}) as c:
        c.set_obstacles(obs, len(specs))
        r = c.generate_check(src, MIN_T, flags=abi.F_EARLY_EXIT, n=n, shared_rows=mask, group_size=G)
        np.testing.assert_array_equal(r["verdict"], o["verdict"])
        np.testing.assert_array_equal(r["first_obstacle"], o["first_obstacle"])
        assert_close(r["cost"], o["cost"], what="cost")
        assert_close(r["coeffs"], o["coeffs"], what="coeffs")
        assert r["summary"]["pair_counts"] == o["summary"]["pair_counts"]
        assert r["summary"]["best_index"] == o["summary"]["best_index"]
        # a sub-range that starts inside a group
        first, n2 = 12_345, 20_011
        r2 = c.generate_check(src, MIN_T, flags=abi.F_EARLY_EXIT, n=n2, shared_rows=mask, group_size=G, first=first,
                              index_base=first)
        np.testing.assert_array_equal(r2["verdict"], o["verdict"][first:first + n2])
        assert_close(r2["cost"], o["cost"][first:first + n2], what="cost")
        # device pointers: the kernel reads the compact matrix in place
        d = torch.from_numpy(compact).cuda()
        ver = torch.zeros(n2, dtype=torch.uint8, device="cuda")
        s = c.generate_check_dev(d.data_ptr(), n, n2, MIN_T, flags=abi.F_EARLY_EXIT, index_base=first, d_verdict=ver.data_ptr(),
                                 want_summary=True, shared_rows=mask, group_size=G, first=first)
        np.testing.assert_array_equal(ver.cpu().numpy(), o["verdict"][first:first + n2])
        assert s == r2["summary"]
        # the other entry points that take boundary conditions
        g = c.generate(src, n=n, shared_rows=mask, group_size=G)
        assert_close(g["cost"], o["cost"], what="cost (rqcd_generate)")
        feas_full = c.input_feasibility(full)
        feas = c.input_feasibility(src, n=n, shared_rows=mask, group_size=G)
        np.testing.assert_array_equal(feas, feas_full)
    del keep


def test_bad_shared_rows_are_refused(ctx):
    bc = W.synth_bc(1, 0, 64)
    lib = ctx.lib
    out = abi.Out()
    for inp in (abi.BcSoa(bc.ctypes.data, 64, abi.GOAL_ALL, 1, 0, 0),       # shared rows without a group size
                abi.BcSoa(bc.ctypes.data, 64, abi.GOAL_ALL, 1 << 19, 4, 0),  # a row that does not exist
                abi.BcSoa(bc.ctypes.data, 64, abi.GOAL_ALL, 0, 0, -1),       # negative first
                abi.BcSoa(bc.ctypes.data, 64, abi.GOAL_ALL, 0, 0, 1)):       # first + n beyond the stride
        assert lib.rqcd_generate_check(ctx.h, C.byref(inp), 64, MIN_T, 0, 0, C.byref(out)) == abi.ERR_INVALID_ARG


@pytest.mark.parametrize("host_slots", [0, 1])
def test_group_of_shards_equals_one_context(host_slots, oracle):
    """rqcd_group_*: candidates sharded by index over the group's contexts (here several shards on the devices present;
    with one GPU all on cuda:0), summaries published into the slot table (peer-mapped device memory or mapped host
    memory) and merged on the first device: outputs and the merged record equal the single-context call and the oracle,
    for any shard count, including a non-empty best candidate."""
    import torch
    ndev = torch.cuda.device_count()
    specs = W.static_obstacles(6, 4, 4)  # sparse enough that some candidates stay collision free
    obs = abi.make_obstacles(specs)
    n = 30_011
    bc = W.synth_bc(23, 0, n)
    o = oracle.generate_check(bc, obs, len(specs), MIN_T, nthreads=8)
    assert o["summary"]["best_index"] >= 0, "the workload must exercise the best-candidate merge"
    old = os.environ.get("RQCD_GROUP_HOST_SLOTS")
    os.environ["RQCD_GROUP_HOST_SLOTS"] = str(host_slots)
    try:
        for shards in (1, 2, 3, 5):
            devices = [k % ndev for k in range(shards)]
            with Group(devices) as g:
                assert g.size() == shards and g.peer_slots() == (not host_slots)
                g.set_obstacles(obs, len(specs))
                before = g.launch_count()
                r = g.generate_check(bc, MIN_T)
                assert g.launch_count() == before + 3 * shards + 1  # per shard: check, finalize, publish; one merge
                np.testing.assert_array_equal(r["verdict_matrix"], o["verdict_matrix"])
                np.testing.assert_array_equal(r["verdict"], o["verdict"])
                np.testing.assert_array_equal(r["first_obstacle"], o["first_obstacle"])
                assert_close(r["cost"], o["cost"], what="cost")
                assert_close(r["coeffs"], o["coeffs"], what="coeffs")
                assert r["summary"]["traj_counts"] == o["summary"]["traj_counts"]
                assert r["summary"]["pair_counts"] == o["summary"]["pair_counts"]
                assert r["summary"]["best_index"] == o["summary"]["best_index"]
                assert r["summary"]["best_cost"] == r["cost"][r["summary"]["best_index"]]
                # resident shards: device pointers per shard, enqueue only, one read of the merged record
                d_bc, ns, keep = [], [], []
                for k in range(shards):
                    lo, hi = n * k // shards, n * (k + 1) // shards
                    t = torch.from_numpy(np.ascontiguousarray(bc[:, lo:hi])).to(f"cuda:{devices[k]}")
                    keep.append(t)
                    d_bc.append(t.data_ptr())
                    ns.append(hi - lo)
                st = g.generate_check_resident(d_bc, ns, ns, MIN_T, want_summary=False)
                assert st == abi.OK
                assert g.read_summary() == r["summary"]
    finally:
        if old is None:
            os.environ.pop("RQCD_GROUP_HOST_SLOTS", None)
        else:
            os.environ["RQCD_GROUP_HOST_SLOTS"] = old


def test_group_with_shared_rows(oracle):
    """Shards of one compact Forest-style matrix: shard boundaries fall inside groups."""
    import torch
    ndev = torch.cuda.device_count()
    specs = W.forest_obstacles()
    obs = abi.make_obstacles(specs)
    n, G = 10_007, 100
    full, compact = _forest_batch(n, G, 41)
    o = oracle.generate_check(full, obs, len(specs), MIN_T, flags=abi.F_EARLY_EXIT, matrix=False, nthreads=8)
    with Group([k % ndev for k in range(3)]) as g:
        g.set_obstacles(obs, len(specs))
        r = g.generate_check(compact, MIN_T, flags=abi.F_EARLY_EXIT, n=n, shared_rows=shared_row_mask(FOREST_SHARED), group_size=G)
    np.testing.assert_array_equal(r["verdict"], o["verdict"])
    assert r["summary"]["pair_counts"] == o["summary"]["pair_counts"]
    assert r["summary"]["best_index"] == o["summary"]["best_index"]


def test_exchange_without_host_round_trip(ctx, oracle):
    """rqcd_exchange_summaries: with no communicator (world 1) the merged record is the local one; with a one-rank NCCL
    communicator the all-gather runs on the context's stream. Per-step records land in a device ring that is read once."""
    import torch
    specs = W.static_obstacles(6, 4, 4)
    obs = abi.make_obstacles(specs)
    ctx.set_obstacles(obs, len(specs))
    n = 8192
    bc = W.synth_bc(29, 0, n)
    o = oracle.generate_check(bc, obs, len(specs), MIN_T, matrix=False, nthreads=8)
    d = torch.from_numpy(bc).cuda()
    ring = torch.zeros((4, 11), dtype=torch.int64, device="cuda")
    assert ctx.comm_world() == 1
    for step in range(2):
        ctx.generate_check_dev(d.data_ptr(), n, n, MIN_T)
        ctx.exchange_summaries(ring[step].data_ptr())
    ctx.comm_init(comm_unique_id(), 0, 1)
    try:
        for step in range(2, 4):
            ctx.generate_check_dev(d.data_ptr(), n, n, MIN_T)
            ctx.exchange_summaries(ring[step].data_ptr())
        merged = ctx.read_merged_summary()
    finally:
        ctx.comm_finalize()
    assert merged["pair_counts"] == o["summary"]["pair_counts"] and merged["best_index"] == o["summary"]["best_index"]
    rec = ring.cpu().numpy()
    for step in range(4):
        assert list(rec[step, 4:8]) == o["summary"]["pair_counts"]
        assert int(rec[step, 10]) == o["summary"]["best_index"]
    assert merge_summaries([merged]) == merged


@pytest.mark.parametrize("shape", ["table", "pairwise", "early_exit"])
def test_page_locked_output_buffers_are_written_directly(shape, oracle):
    """Host-pointer calls whose verdict / first-hit / cost buffers are page-locked get their results by direct stores of the
    kernel (no staging copy after the launch); pageable buffers are staged. Both must hold the same bytes, and the
    oracle's; RQCD_DIRECT_OUT=0 switches the direct path off."""
    import torch
    n = 77_777
    bc = W.synth_bc(31, 0, n)
    keep_in, bc_pin = _pinned(bc)
    specs = W.static_obstacles(33, 10, 10) if shape != "early_exit" else W.forest_obstacles()
    obs = abi.make_obstacles(specs)
    u = W.synth_bc(32, 0, n)
    sph = np.empty((4, n))
    sph[0:3] = u[3:6]
    sph[3] = 0.1 + (u[6] + 4.0) / 8.0 * 1.4
    flags = abi.F_EARLY_EXIT if shape == "early_exit" else 0

    def call(ctx, pinned_out):
        mk = (lambda dt: torch.empty(n, dtype=dt).pin_memory().numpy()) if pinned_out else (lambda dt: torch.empty(n, dtype=dt).numpy())
        v, f, c = mk(torch.uint8), mk(torch.int32), mk(torch.float64)
        v[:], f[:], c[:] = 77, -7, -1.0
        summ = abi.Summary()
        inp = abi.BcSoa(bc_pin.ctypes.data, n, abi.GOAL_ALL, 0, 0, 0)
        if shape == "pairwise":
            out = abi.Out(verdict=v.ctypes.data, cost=c.ctypes.data, summary=C.pointer(summ))
            st = ctx.lib.rqcd_generate_check_pairwise(ctx.h, C.byref(inp), sph.ctypes.data, n, n, MIN_T, 0, 0, C.byref(out))
        else:
            out = abi.Out(verdict=v.ctypes.data, first_obstacle=f.ctypes.data, cost=c.ctypes.data, summary=C.pointer(summ))
            st = ctx.lib.rqcd_generate_check(ctx.h, C.byref(inp), n, MIN_T, flags, 0, C.byref(out))
        assert st == abi.OK
        return v.copy(), f.copy(), c.copy(), summ.as_dict()

    with Context(0) as ctx:
        if shape != "pairwise":
            ctx.set_obstacles(obs, len(specs))
        direct = call(ctx, True)
        staged = call(ctx, False)
    with _fresh_ctx({"RQCD_DIRECT_OUT": 0}) as c0:
        if shape != "pairwise":
            c0.set_obstacles(obs, len(specs))
        off = call(c0, True)
    if shape == "pairwise":
        o = oracle.generate_check_pairwise(bc, sph, MIN_T, nthreads=8)
    else:
        o = oracle.generate_check(bc, obs, len(specs), MIN_T, flags=flags, matrix=False, nthreads=8)
    for got in (direct, staged, off):
        np.testing.assert_array_equal(got[0], o["verdict"])
        if shape != "pairwise":
            np.testing.assert_array_equal(got[1], o["first_obstacle"])
        assert_close(got[2], o["cost"], what="cost")
        assert got[3]["traj_counts"] == o["summary"]["traj_counts"]
    np.testing.assert_array_equal(direct[2], staged[2])
