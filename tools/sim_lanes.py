#!/usr/bin/env python
"""tools/sim_lanes.py [n_traj=4096] [c3|c5] -- lane-scheduling model of the check kernels on the C3 (or C5) workload, fed
with the REAL number of CollisionCheckSection frames of every (trajectory, obstacle) pair (counted by the CPU oracle).

It answers one question for the designs discussed in DESIGN.md section 5: what fraction of the lane slots of a warp
carries a frame? Policies:
  free     : k_check -- a lane walks its trajectory's obstacles, one frame per trip; a warp refills when `refill`
             lanes are idle (measured on the GPU: 90 %)
  uni      : the first-frame experiment -- 32 trajectories per warp, first frames in obstacle order at full lanes
             (phase A), extra frames from per-lane stacks when >= `trigger` lanes have some (measured: phase B 29 %)
  uni-pool : the same with the parked pairs of the warp in ONE pool served to any free lane (a pair's frames stay
             sequential: hi subtree before lo)
CPU only; the oracle is used as the frame counter."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import binding as B
from rapidquadcoptercollisiondetection_b200 import abi, workloads as W


def frames_per_pair(n, workload="c3"):
    orc = B.Oracle()
    if workload == "c5":
        specs = W.static_obstacles(6, 128, 128)  # the C5 table (bench.py: obs_seed 6, candidates seed 5)
        bc = W.synth_bc(5, 0, n)
    else:
        specs = W.static_obstacles(2, 32, 32)  # the C3 table (bench.py: obs_seed 2, candidates seed 1)
        bc = W.synth_bc(1, 0, n)
    obs = abi.make_obstacles(specs)
    g = orc.generate(bc)
    co = np.ascontiguousarray(g["coeffs"].T)  # n x 18
    fn = orc._f("check_one", C.c_int)
    out = np.zeros((n, len(specs)), dtype=np.int32)
    st = B.Stats()
    for i in range(n):
        c18 = co[i]
        tf = float(bc[abi.BC_TF, i])
        for k in range(len(specs)):
            C.memset(C.byref(st), 0, C.sizeof(st))
            fn(c18.ctypes.data_as(C.c_void_p), C.c_double(0.0), C.c_double(tf), C.byref(obs[k]), C.c_double(0.002), C.byref(st))
            out[i, k] = st.sections
    return out


def sim_free(fr, refill):
    """k_check: returns (lane-slot utilisation, trips per pair)."""
    n, m = fr.shape
    nxt = 0
    lanes = [None] * 32  # [traj, k, frames left in the current pair]
    used = trips = 0
    while True:
        idle = [j for j in range(32) if lanes[j] is None]
        if nxt < n and (len(idle) >= refill or len(idle) == 32):
            for j in idle:
                if nxt < n:
                    lanes[j] = [nxt, 0, max(int(fr[nxt, 0]), 0)]
                    nxt += 1
        if all(l is None for l in lanes):
            break
        trips += 1
        for j in range(32):
            l = lanes[j]
            if l is None:
                continue
            # pairs without a section (end point inside) cost no frame: skip them
            while l[2] == 0 and l[1] < m - 1:
                l[1] += 1
                l[2] = int(fr[l[0], l[1]])
            if l[2] == 0:
                lanes[j] = None
                continue
            used += 1
            l[2] -= 1
            if l[2] == 0:
                if l[1] == m - 1:
                    lanes[j] = None
                else:
                    l[1] += 1
                    l[2] = int(fr[l[0], l[1]])
    return used / (32.0 * trips), trips / (n * m / 32.0)


def sim_uni(fr, trigger, pool):
    """first frames in obstacle order; extra frames from per-lane stacks (pool=False) or a warp pool (pool=True).
    Returns (phase A trips, phase B trips, phase B utilisation) per batch average."""
    n, m = fr.shape
    a_trips = b_trips = b_used = 0
    for b0 in range(0, n - 31, 32):
        extra = np.maximum(fr[b0:b0 + 32] - 1, 0)
        stacks = [[] for _ in range(32)]  # remaining frames of parked pairs
        poolq = []
        busy = [0] * 32  # pool mode: frames left of the pair a lane is working on
        k = 0
        while True:
            if pool:
                pending = len(poolq) + sum(1 for x in busy if x > 0)
            else:
                pending = sum(1 for s in stacks if s)
            if k < m and pending < trigger:
                a_trips += 1
                for j in range(32):
                    e = int(extra[j, k])
                    if e > 0:
                        (poolq if pool else stacks[j]).append(e)
                k += 1
            elif pending > 0:
                b_trips += 1
                if pool:
                    for j in range(32):
                        if busy[j] == 0 and poolq:
                            busy[j] = poolq.pop()
                        if busy[j] > 0:
                            busy[j] -= 1
                            b_used += 1
                else:
                    for s in stacks:
                        if s:
                            s[-1] -= 1
                            b_used += 1
                            if s[-1] == 0:
                                s.pop()
            else:
                break
    nb = (n // 32)
    return a_trips / nb, b_trips / nb, b_used / (32.0 * max(b_trips, 1))


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    wl = sys.argv[2] if len(sys.argv) > 2 else "c3"
    fr = frames_per_pair(n, wl)
    print(f"{wl} sample: {n} trajectories x {fr.shape[1]} obstacles, {fr.mean():.3f} frames per pair, "
          f"{(fr > 1).mean() * 100:.1f} % of the pairs need more than one, max {fr.max()}")
    per_traj = np.maximum(fr - 1, 0).sum(1)
    print(f"extra frames per trajectory: mean {per_traj.mean():.1f}, median {np.median(per_traj):.0f}, "
          f"90th pct {np.percentile(per_traj, 90):.0f}, max {per_traj.max()}")
    for refill in (4, 6, 8):
        u, t = sim_free(fr, refill)
        print(f"free     refill {refill}: lane slots used {u * 100:.1f} %")
    for trig in (12, 20, 28):
        a, b, u = sim_uni(fr, trig, False)
        print(f"uni      trigger {trig:2d}: {a:.0f} phase A + {b:.1f} phase B trips per batch, phase B lanes used {u * 100:.1f} %")
    for trig in (32, 64, 96):
        a, b, u = sim_uni(fr, trig, True)
        print(f"uni-pool trigger {trig:2d}: {a:.0f} phase A + {b:.1f} phase B trips per batch, phase B lanes used {u * 100:.1f} %")
