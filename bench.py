#!/usr/bin/env python
"""bench.py -- trajectory-obstacle checks/s of the fused generate+check path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c4|c5|c1|c2] [--impl reference]

A step = one pass of the hot path (rqcd_generate_check through the C ABI, librqcd.so) over one batch of
synthetic candidates. Workloads are the configurations of BASELINE.json / SURVEY.md section 8(d):
  c3 (default) 2^20 primitives x 64 static obstacles (32 spheres + 32 prisms), all pairs
  c4           2^18 primitives x 32 moving obstacles
  c5           2^26 primitives x 256 static obstacles, sharded by candidate index over the ranks (strong scaling)
  c1           PerformanceEval shape: trajectory i vs its own sphere i (pairwise), 2^22 unfiltered trials
  c2           ForestPerformanceEval: 10^4 x 100 candidates x 5 prisms, early exit
With N > 1 (torchrun, one rank per GPU) every rank holds the whole obstacle set and its own candidate shard;
c3/c4/c1/c2 keep the per-rank batch fixed (weak scaling), c5 splits 2^26 (strong). The only exchange is the
all-gather of the 88-byte per-shard summaries (verdict counts + best candidate), merged by
rqcd_merge_summaries.

`value` is timed with CUDA events on the launching stream with inputs resident in HBM; `e2e` is the same call
with pinned HOST buffers (H2D of the boundary conditions and D2H of verdicts/first hits/costs inside the
timed region). `--impl reference` times the reference's own CPU implementation (oracle/_ref/libref.so, the
unmodified reference classes; the oracle port if that file is absent) on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# The driver reads ONE JSON line on stdout. Native libraries (NCCL prints its version banner to fd 1) must not get
# there: fd 1 is pointed at stderr for the whole run and the JSON line is written to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

MIN_T = 0.002
METRIC = "trajectory_obstacle_checks_per_sec"
UNIT = "checks/s"

WORKLOADS = {
    # name: (per-rank candidates, obstacles, description, scaling)
    "c3": dict(n=1 << 20, m=64, seed=1, obs_seed=2, desc="2^20 primitives x 64 static obstacles (32 spheres + 32 prisms), all pairs", scaling="weak"),
    "c4": dict(n=1 << 18, m=32, seed=3, obs_seed=4, desc="2^18 primitives x 32 moving non-rotating obstacles, all pairs", scaling="weak"),
    "c5": dict(n=1 << 26, m=256, seed=5, obs_seed=6, desc="2^26 primitives x 256 static obstacles (128 spheres + 128 prisms) sharded by candidate index", scaling="strong"),
    "c1": dict(n=1 << 22, m=1, seed=7, obs_seed=8, desc="PerformanceEval shape: 2^22 unfiltered trials, trajectory i vs its own sphere i", scaling="weak"),
    "c2": dict(n=1_000_000, m=5, seed=0, obs_seed=0, desc="ForestPerformanceEval: 10^4 batches x 100 candidates x 5 prisms, early exit", scaling="weak"),
}


# ---------------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi while the timed region runs
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.06)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self) -> dict:
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                power.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


# ---------------------------------------------------------------------------------------------------------
# the algorithmic flop model of SURVEY.md section 8(d), fed by the oracle's work counters
# ---------------------------------------------------------------------------------------------------------
def flops_per_check(stats: dict, m: int) -> float:
    checks = max(stats["checks"], 1)
    fs, fp = stats["sphere_checks"] / checks, stats["prism_checks"] / checks
    f_in = 9 * fs + 18 * fp
    f_tan = 18 * fs + 48 * fp
    S, St, P = stats["sections"] / checks, stats["tangent_sections"] / checks, stats["test_points"] / checks
    solves = max(stats["trig"] + stats["cardano"], 1)
    f_q = 46 + (20 * stats["trig"] + 13 * stats["cardano"]) / solves + 5 * stats["root_pairs"] / solves
    per_check = 2 * (60 + f_in) + S * (63 + f_in) + St * (f_tan + 34 + f_q) + P * 68
    return per_check + 210.0 / max(m, 1)  # + generation (0.21 kflop per trajectory) amortised over its obstacles


# ---------------------------------------------------------------------------------------------------------
def build_obstacles(name: str, W, abi, gen_coeffs):
    w = WORKLOADS[name]
    if name in ("c3", "c5"):
        specs = W.static_obstacles(w["obs_seed"], w["m"] // 2, w["m"] // 2)
    elif name == "c4":
        specs = W.moving_obstacles(w["obs_seed"], w["m"], gen_coeffs)
    elif name == "c2":
        specs = W.forest_obstacles()
    else:
        specs = []
    return specs, abi.make_obstacles(specs)


def host_inputs(name: str, W, base: int, n: int):
    """Boundary conditions (19 x n) [+ spheres 4 x n for c1] for global indices [base, base+n)."""
    w = WORKLOADS[name]
    if name == "c2":
        # ForestPerformanceEval's distributions (ForestPerformanceEval.cpp:77-102, 175-193), counter-based
        u = W.synth_bc(w["seed"] + 100, base, n)  # U(-4,4) rows re-mapped below
        unit = lambda r: (u[r] + 4.0) / 8.0  # noqa: E731
        bc = np.zeros((19, n))
        bc[0] = -2.5
        batch = (np.arange(base, base + n) // 100).astype(np.int64)
        ub = W.synth_bc(w["seed"] + 101, 0, int(batch.max()) + 1)
        unit_b = lambda r: ((ub[r] + 4.0) / 8.0)[batch]  # noqa: E731
        bc[3] = 2.0 + 6.0 * unit_b(3)
        bc[4] = -2.0 + 4.0 * unit_b(4)
        bc[5] = -2.0 + 4.0 * unit_b(5)
        bc[6] = 4.0 + 6.0 * unit_b(6)
        bc[7] = -2.0 + 4.0 * unit_b(7)
        bc[8] = -2.0 + 4.0 * unit_b(8)
        for r in (9, 10, 11):
            bc[r] = -2.5 + 5.0 * unit(r)
        bc[18] = 0.5 + 1.5 * unit(12)
        return bc, None
    bc = W.synth_bc(w["seed"], base, n)
    sph = None
    if name == "c1":
        u = W.synth_bc(w["obs_seed"], base, n)
        sph = np.empty((4, n))
        sph[0:3] = u[3:6]                        # centre U(-4,4)^3
        sph[3] = 0.1 + (u[6] + 4.0) / 8.0 * 1.4  # radius U[0.1,1.5]
    return bc, sph


def cpu_checker():
    from oracle import binding as B
    if B.have_reference():
        return B.Reference(), "reference"
    return B.Oracle(), "port"


def run_cpu(impl, name, obs, m, bc, sph, threads):
    from rapidquadcoptercollisiondetection_b200 import abi
    t0 = time.perf_counter()
    if name == "c1":
        r = impl.generate_check_pairwise(bc, sph, MIN_T, nthreads=threads)
    else:
        flags = abi.F_EARLY_EXIT if name == "c2" else 0
        r = impl.generate_check(bc, obs, m, MIN_T, flags=flags, matrix=False, nthreads=threads)
    dt = time.perf_counter() - t0
    return r, dt


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from rapidquadcoptercollisiondetection_b200 import abi, workloads as W
    impl, kind = cpu_checker()
    name = args.workload
    w = WORKLOADS[name]
    threads = os.cpu_count() or 1
    gen = lambda b: impl.generate(b)["coeffs"]  # noqa: E731
    specs, obs = build_obstacles(name, W, abi, gen)
    m = max(len(specs), 1)
    # bounded sample of the same workload: ~1.6e7 pair checks per step
    n_s = max(1024, min(w["n"], (1 << 24) // m))
    bc, sph = host_inputs(name, W, 0, n_s)
    times, pairs = [], 0
    for it in range(args.warmup + args.steps):
        r, dt = run_cpu(impl, name, obs, len(specs), bc, sph, threads)
        if it >= args.warmup:
            times.append(dt)
            pairs += r["summary"]["pairs_checked"]
    total = sum(times)
    value = pairs / total
    sample = f"first {n_s} candidates of {name} x {len(specs) if name != 'c1' else 1} obstacles per step"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True,
        "scaling": w["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{name}: {w['desc']}", "sample": sample, "min_time_section": MIN_T},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    from rapidquadcoptercollisiondetection_b200 import abi, sharding, workloads as W
    from rapidquadcoptercollisiondetection_b200.api import Context

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the path has no CPU implementation in this repo")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    name = args.workload
    w = WORKLOADS[name]
    n_global = w["n"] if w["scaling"] == "strong" else w["n"] * world
    lo, hi = sharding.shard_range(n_global, rank, world)
    n = hi - lo

    ctx = Context(local)
    stream = torch.cuda.Stream(dev)  # the one stream everything is launched and timed on
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    specs, obs = build_obstacles(name, W, abi, lambda b: ctx.generate(b)["coeffs"])
    m = len(specs)
    if name != "c1":
        ctx.set_obstacles(obs, m)
    pairs_per_traj = 1 if name == "c1" else m
    flags = abi.F_EARLY_EXIT if name == "c2" else 0

    # ---- inputs resident in HBM -------------------------------------------------------------------------
    d_bc = torch.empty((abi.BC_ROWS, n), dtype=torch.float64, device=dev)
    d_sph = None
    if name in ("c2", "c1"):
        bc_h, sph_h = host_inputs(name, W, lo, n)
        d_bc.copy_(torch.from_numpy(bc_h))
        if sph_h is not None:
            d_sph = torch.from_numpy(sph_h).to(dev)
    else:
        ctx.synth_bc(w["seed"], lo, n, d_bc.data_ptr(), n)  # generated on the device from (seed, global index)
    d_verdict = torch.empty(n, dtype=torch.uint8, device=dev)
    d_first = torch.empty(n, dtype=torch.int32, device=dev)
    d_cost = torch.empty(n, dtype=torch.float64, device=dev)
    torch.cuda.synchronize(dev)

    gather_buf = torch.zeros((world, 11), dtype=torch.int64, device=dev) if world > 1 else None

    def step_device():
        if name == "c1":
            s = ctx.generate_check_pairwise_dev(d_bc.data_ptr(), n, d_sph.data_ptr(), n, n, MIN_T, index_base=lo,
                                                d_verdict=d_verdict.data_ptr(), d_cost=d_cost.data_ptr(), want_summary=True)
        else:
            s = ctx.generate_check_dev(d_bc.data_ptr(), n, n, MIN_T, flags=flags, index_base=lo,
                                       d_verdict=d_verdict.data_ptr(), d_first=d_first.data_ptr(),
                                       d_cost=d_cost.data_ptr(), want_summary=True)
        return exchange(s)

    def exchange(s):
        """The one exchange step: all-gather of the per-shard summary records over NCCL, merged on every rank."""
        return sharding.exchange_summaries(s, dev, gather_buf)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing ---------------------------------------------------------------------------
    for _ in range(args.warmup):
        summary = step_device()
    barrier()
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        ev0.record(stream)
        pairs_total = 0
        for _ in range(args.steps):
            summary = step_device()
            pairs_total += summary["pairs_checked"]
        ev1.record(stream)
        barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = pairs_total / (ms * 1e-3)
    clocks = clk.summary()

    # ---- end to end: pinned host buffers through the host-pointer C-ABI call --------------------------------
    e2e = None
    if not args.no_e2e:
        # the whole shard when it is at most 2^22 candidates (c3: all 2^20); a 2^22 prefix of it otherwise (c5)
        ne = min(n, 1 << 22)
        if name in ("c1", "c2"):
            bc_host = torch.from_numpy(np.ascontiguousarray(bc_h[:, :ne])).pin_memory()
            sph_host = torch.from_numpy(np.ascontiguousarray(sph_h[:, :ne])).pin_memory() if sph_h is not None else None
        else:
            bc_host = torch.empty((abi.BC_ROWS, ne), dtype=torch.float64).pin_memory()
            bc_host.copy_(d_bc[:, :ne])
            sph_host = None
        v_host = torch.empty(ne, dtype=torch.uint8).pin_memory()
        f_host = torch.empty(ne, dtype=torch.int32).pin_memory()
        c_host = torch.empty(ne, dtype=torch.float64).pin_memory()
        import ctypes as C
        summ = abi.Summary()
        lib = ctx.lib

        def step_host():
            inp = abi.BcSoa(bc_host.data_ptr(), ne, abi.GOAL_ALL, 0)
            if name == "c1":
                out = abi.Out(verdict=v_host.data_ptr(), cost=c_host.data_ptr(), summary=C.pointer(summ))
                st = lib.rqcd_generate_check_pairwise(ctx.h, C.byref(inp), sph_host.data_ptr(), ne, ne, MIN_T, 0, lo, C.byref(out))
            else:
                out = abi.Out(verdict=v_host.data_ptr(), first_obstacle=f_host.data_ptr(), cost=c_host.data_ptr(),
                              summary=C.pointer(summ))
                st = lib.rqcd_generate_check(ctx.h, C.byref(inp), ne, MIN_T, flags, lo, C.byref(out))
            assert st == abi.OK, st
            return exchange(summ.as_dict())

        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(3):
            step_host()
        barrier()
        ev0.record(stream)
        e_pairs = 0
        for _ in range(e2e_steps):
            e_pairs += step_host()["pairs_checked"]
        ev1.record(stream)
        barrier()
        e_ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([e_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t.item())
        h2d = 19 * ne * 8 + (4 * ne * 8 if name == "c1" else 0)
        d2h = ne * (1 + 8) + (0 if name == "c1" else 4 * ne) + 88
        e2e = {"value": e_pairs / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e_ms / e2e_steps, "steps": e2e_steps, "candidates_per_rank_per_step": ne}

    # ---- rank 0: CPU baseline on a bounded sample, flop model, roofline ---------------------------------------
    if rank == 0:
        from oracle import binding as B
        orc = B.Oracle()
        n_s = max(1024, min(n, (1 << 22) // max(pairs_per_traj, 1)))
        bc_s, sph_s = host_inputs(name, W, lo, n_s)
        if name == "c1":
            o = orc.generate_check_pairwise(bc_s, sph_s, MIN_T, nthreads=os.cpu_count() or 1, stats=True)
        else:
            o = orc.generate_check(bc_s, obs, m, MIN_T, flags=flags, matrix=False, nthreads=os.cpu_count() or 1, stats=True)
        fpc = flops_per_check(o["stats"], pairs_per_traj)
        # parity spot check of the resident results against the oracle sample (the checker, not the product)
        got = d_verdict[:n_s].cpu().numpy()
        parity = int((got != o["verdict"]).sum())

        cpu_baseline = None
        if not args.no_cpu_baseline:
            impl, kind = cpu_checker()
            threads = os.cpu_count() or 1
            n_c = max(1024, min(n, (1 << 25) // max(pairs_per_traj, 1)))  # ~1 s on 16 threads = ~16 core-seconds
            bc_c, sph_c = host_inputs(name, W, lo, n_c)
            run_cpu(impl, name, obs, m, bc_c[:, :1024].copy(), None if sph_c is None else sph_c[:, :1024].copy(), threads)
            r, dt = run_cpu(impl, name, obs, m, bc_c, sph_c, threads)
            cpu_baseline = {"value": r["summary"]["pairs_checked"] / dt, "unit": UNIT, "cores": threads, "kind": kind,
                            "sample": f"first {n_c} candidates of rank 0's shard x {pairs_per_traj} obstacles, {dt:.2f} s wall"}

        peaks = ctx.measure_fp64_peak(200)
        kernel_s = (ms * 1e-3) / args.steps  # k_check dominates the step (k_finalize + memset ~ microseconds)
        pairs_per_launch = pairs_total / args.steps / world
        achieved = fpc * pairs_per_launch / kernel_s / 1e12
        peak_tf = peaks["dmul_dadd_flops"] / 1e12
        measured = {}
        try:
            measured = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bytes_per_launch = n * (19 * 8 + 1 + 4 + 8)
        traffic = None
        try:  # DRAM bytes of one launch from the committed ncu capture of this workload (not measured in this run)
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[name]["dram_bytes_per_launch"]
        except Exception:
            pass
        roofline = {
            "bound": "fp64", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
            "traffic": traffic,
            "peak_source": "measured in this run by rqcd_measure_fp64_peak: dependent-free DMUL+DADD chains (the "
                           "ceiling of the -fmad=false exact mode); dfma_peak is the FMA-counted figure",
            "dfma_peak": peaks["dfma_flops"] / 1e12, "frac_of_dfma_peak": achieved / (peaks["dfma_flops"] / 1e12),
            "flops_per_check": fpc, "kernel_ms": kernel_s * 1e3,
            "hbm": {"achieved_gbs": bytes_per_launch / kernel_s / 1e9, "peak_gbs": measured.get("hbm_gbs"),
                    "algorithmic_bytes_per_launch": bytes_per_launch},
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{name}: {w['desc']}", "candidates_per_rank": n, "candidates_total": n_global,
                       "obstacles": pairs_per_traj, "min_time_section": MIN_T, "mode": "exact (no FMA contraction)",
                       "l2": "inputs (19 x n doubles per rank) exceed the 126 MB L2" if 19 * n * 8 > 126e6 else
                             "inputs fit in L2; the kernel is FP64-bound (0.16 KB per candidate), no flush",
                       "parallelism": f"candidate-index shards x{world}, obstacle set replicated"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "summary": summary, "parity_mismatches_vs_oracle_sample": parity, "oracle_sample": n_s,
        }
        emit(line)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
