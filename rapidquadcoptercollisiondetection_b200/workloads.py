"""Synthetic inputs for the configurations of BASELINE.json / SURVEY.md section 8(d) (host-side numpy).

This is input synthesis, not the path: it only draws boundary conditions and obstacle parameters. The
trajectory generator shared by host and device is the counter-based one of rqcd_synth_bc (splitmix64 of
(seed, row, global index)); `synth_bc` below is its numpy form so that a host test can regenerate any index.
"""
from __future__ import annotations

import math

import numpy as np

from . import abi

_M64 = (1 << 64) - 1


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(_M64)
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(_M64)
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & np.uint64(_M64)
    return x ^ (x >> np.uint64(31))


def synth_bc(seed: int, index_base: int, n: int) -> np.ndarray:
    """19 x n boundary conditions, identical to rqcd_synth_bc / orc_synth_bc.
    Ranges of test/PerformanceEval.cpp:59-65: pos0 = 0, vel0/acc0/posf/velf/accf U(-4,4), Tf U(0.2,4)."""
    return synth_bc_at(seed, np.arange(index_base, index_base + n, dtype=np.uint64))


def synth_bc_at(seed: int, indices) -> np.ndarray:
    """synth_bc for an arbitrary list of global candidate indices (e.g. every 1024-th of a shard)."""
    with np.errstate(over="ignore"):
        g = np.asarray(indices).astype(np.uint64)
        bc = np.empty((abi.BC_ROWS, len(g)))
        for r in range(abi.BC_ROWS):
            k = _splitmix64(np.array([np.uint64(seed) ^ (np.uint64(r) * np.uint64(0xD6E8FEB86659FD93))], dtype=np.uint64))
            h = _splitmix64(k + g)
            u = (h >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
            if r < abi.BC_V0:
                bc[r] = 0.0
            elif r == abi.BC_TF:
                bc[r] = 0.2 + u * (4.0 - 0.2)
            else:
                bc[r] = -4.0 + u * 8.0
    return bc


def quat_from_rotation_vector(v) -> list:
    """Unit quaternion (w, x, y, z) of the rotation vector v (angle = |v|)."""
    v = np.asarray(v, dtype=np.float64)
    ang = float(np.linalg.norm(v))
    if ang < 1e-12:
        return [1.0, 0.0, 0.0, 0.0]
    ax = v / ang
    s = np.sin(ang / 2)
    return [float(np.cos(ang / 2)), float(ax[0] * s), float(ax[1] * s), float(ax[2] * s)]


def static_obstacles(seed: int, n_spheres: int, n_prisms: int, extent: float = 4.0) -> list:
    """spheres (centre U(+-extent)^3, r U[0.1,1.5]) then prisms (centre U(+-extent)^3,
    sides U[0.2,3]^3, rotation vector U(+-pi)^3)."""
    rng = np.random.default_rng(seed)
    specs = []
    for _ in range(n_spheres):
        specs.append({"type": "sphere", "center": rng.uniform(-extent, extent, 3).tolist(),
                      "radius": float(rng.uniform(0.1, 1.5))})
    for _ in range(n_prisms):
        specs.append({"type": "prism", "center": rng.uniform(-extent, extent, 3).tolist(),
                      "side": rng.uniform(0.2, 3.0, 3).tolist(),
                      "quat": quat_from_rotation_vector(rng.uniform(-np.pi, np.pi, 3))})
    return specs


def moving_obstacles(seed: int, m: int, generate) -> list:
    """obstacle k translates along a quintic generated from the C1 state distribution (its own
    Tf in [0.2,4], start 0) and carries a sphere (even k) or a fixed-orientation prism (odd k) at the origin of
    the moving frame. `generate(bc) -> 18 x m coefficients` is the caller's generator (GPU path or checker)."""
    rng = np.random.default_rng(seed)
    bc = np.zeros((abi.BC_ROWS, m))
    bc[abi.BC_P0:abi.BC_P0 + 3] = rng.uniform(-4, 4, (3, m))  # where the obstacle starts
    bc[abi.BC_V0:abi.BC_TF] = rng.uniform(-4, 4, (15, m))
    bc[abi.BC_TF] = rng.uniform(0.2, 4.0, m)
    coeffs = np.asarray(generate(bc))
    specs = []
    for k in range(m):
        s = {"center": [0.0, 0.0, 0.0], "motion": coeffs[:, k].tolist(), "motion_t_start": 0.0,
             "motion_t_end": float(bc[abi.BC_TF, k])}
        if k % 2 == 0:
            s.update(type="sphere", radius=float(rng.uniform(0.1, 1.5)))
        else:
            s.update(type="prism", side=rng.uniform(0.2, 3.0, 3).tolist(),
                     quat=quat_from_rotation_vector(rng.uniform(-np.pi, np.pi, 3)))
        specs.append(s)
    return specs


def specs_to_arrays(specs) -> dict:
    """Flat numpy form of an obstacle list (for .npz fixtures)."""
    m = len(specs)
    out = {"type": np.zeros(m, np.int32), "moving": np.zeros(m, np.int32), "center": np.zeros((m, 3)),
           "radius": np.zeros(m), "side": np.zeros((m, 3)), "quat": np.tile([1.0, 0, 0, 0], (m, 1)),
           "motion": np.zeros((m, 18)), "mt0": np.zeros(m), "mt1": np.zeros(m)}
    for k, s in enumerate(specs):
        out["type"][k] = 0 if s["type"] == "sphere" else 1
        out["center"][k] = s["center"]
        out["radius"][k] = s.get("radius", 0.0)
        out["side"][k] = s.get("side", (0, 0, 0))
        out["quat"][k] = s.get("quat", (1, 0, 0, 0))
        if "motion" in s:
            out["moving"][k] = 1
            out["motion"][k] = np.asarray(s["motion"]).reshape(18)
            out["mt0"][k] = s["motion_t_start"]
            out["mt1"][k] = s["motion_t_end"]
    return out


def arrays_to_specs(a, prefix: str = "") -> list:
    g = lambda k: a[prefix + k]  # noqa: E731
    specs = []
    for k in range(len(g("type"))):
        s = {"type": "sphere" if g("type")[k] == 0 else "prism", "center": g("center")[k].tolist()}
        if g("type")[k] == 0:
            s["radius"] = float(g("radius")[k])
        else:
            s["side"] = g("side")[k].tolist()
            s["quat"] = g("quat")[k].tolist()
        if g("moving")[k]:
            s["motion"] = g("motion")[k].tolist()
            s["motion_t_start"] = float(g("mt0")[k])
            s["motion_t_end"] = float(g("mt1")[k])
        specs.append(s)
    return specs


# ---- the reference drivers' own input streams: std::mt19937(0) + libstdc++ uniform_real_distribution<double> --------
def mt19937_canonical(n: int, seed: int = 0) -> np.ndarray:
    """n draws of std::generate_canonical<double, 53>(std::mt19937(seed)) as libstdc++ computes it: two 32-bit outputs per
    value, low word first, (g1 + g2 * 2^32) / 2^64 in double arithmetic (a result that rounds to 1 becomes
    nextafter(1, 0)). numpy's MT19937 with the legacy seeding is init_genrand(seed), the engine's own seeding."""
    bg = np.random.MT19937()
    bg._legacy_seeding(seed)
    raw = bg.random_raw(2 * n).astype(np.float64)
    ret = (raw[0::2] + raw[1::2] * 4294967296.0) / 18446744073709551616.0
    ret[ret >= 1.0] = np.nextafter(1.0, 0.0)
    return ret


def c1_trials(n_trials: int):
    """The first n_trials trials of test/PerformanceEval.cpp:117-148, accepted or not: 20 draws each in the order
    vel0, acc0, posf, velf, accf (3 each, U(-4, 4); g++ evaluates the three constructor arguments right to left, so
    the first draw of a vector is its z), Tf U(0.2, 4), obstacle position (3), radius U(0.1, 1.5); pos0 = 0.
    Returns (bc 19 x n, spheres 4 x n). The driver keeps the trials that pass CheckInputFeasibility(5, 30, 20, 0.002)
    until it has 10^7 of them: that takes 15 571 728 trials (tests/golden/golden_counts.json: c1_trials)."""
    u = mt19937_canonical(20 * n_trials).reshape(n_trials, 20)
    bc = np.zeros((abi.BC_ROWS, n_trials))
    sph = np.empty((4, n_trials))
    sym = lambda x: x * 8.0 + -4.0  # noqa: E731  uniform_real_distribution: canonical * (b - a) + a
    for v, row in enumerate((abi.BC_V0, abi.BC_A0, abi.BC_PF, abi.BC_VF, abi.BC_AF)):
        for k in range(3):
            bc[row + 2 - k] = sym(u[:, 3 * v + k])
    bc[abi.BC_TF] = u[:, 15] * (4.0 - 0.2) + 0.2
    for k in range(3):
        sph[2 - k] = sym(u[:, 16 + k])
    sph[3] = u[:, 19] * (1.5 - 0.1) + 0.1
    return bc, sph


def c2_candidates(n_batches: int, per_batch: int = 100) -> np.ndarray:
    """test/ForestPerformanceEval.cpp:167-193: per batch vel0 = (U[2,8], U+-2, U+-2), acc0 = (U[4,10], U+-2, U+-2) (z drawn
    first), pos0 = (-2.5, 0, 0); per candidate posf U(+-2.5)^3 (z first) and Tf U[0.5, 2]; goal velocity = acceleration = 0.
    Returns bc 19 x (n_batches * per_batch), candidates of a batch contiguous."""
    per = 6 + 4 * per_batch
    u = mt19937_canonical(per * n_batches).reshape(n_batches, per)
    n = n_batches * per_batch
    bc = np.zeros((abi.BC_ROWS, n))
    bc[abi.BC_P0] = -2.5
    rep = lambda x: np.repeat(x, per_batch)  # noqa: E731
    bc[abi.BC_V0 + 2] = rep(u[:, 0] * 4.0 + -2.0)
    bc[abi.BC_V0 + 1] = rep(u[:, 1] * 4.0 + -2.0)
    bc[abi.BC_V0 + 0] = rep(u[:, 2] * 6.0 + 2.0)
    bc[abi.BC_A0 + 2] = rep(u[:, 3] * 4.0 + -2.0)
    bc[abi.BC_A0 + 1] = rep(u[:, 4] * 4.0 + -2.0)
    bc[abi.BC_A0 + 0] = rep(u[:, 5] * 6.0 + 4.0)
    c = u[:, 6:].reshape(n_batches, per_batch, 4)
    for k in range(3):
        bc[abi.BC_PF + 2 - k] = (c[:, :, k] * 5.0 + -2.5).reshape(n)
    bc[abi.BC_TF] = (c[:, :, 3] * 1.5 + 0.5).reshape(n)
    return bc


def forest_obstacles() -> list:
    """The five trees of test/ForestPerformanceEval.cpp:130-136 (two tilted +-45 deg about x)."""
    pos = [(-1.75, 1.5, 0), (0.5, -1.5, 0), (1.5, 0.5, 0), (-1.0, -1.0, 0), (0.0, 0.8, -0.3)]
    specs = []
    for i, p in enumerate(pos):
        q = [1.0, 0.0, 0.0, 0.0]
        if i >= 3:
            ang = (45 if i == 3 else -45) * math.pi / 180  # libm, as Rotation::FromAxisAngle (Rotation.hpp:92-97)
            q = [math.cos(ang * 0.5), math.sin(ang * 0.5), 0.0, 0.0]
        specs.append({"type": "prism", "center": list(p), "side": [0.5, 0.5, 5.0], "quat": q})
    return specs
