#!/usr/bin/env python
"""tools/c5_expected.py [log2_n] -- the merged summary of BASELINE config 5 (2^26 candidates x 256 static obstacles) from the
CPU ORACLE alone, chunk by chunk, merged with the library's host rule (rqcd_merge_summaries). Written to
tests/golden/c5_expected.json; bench.py's c5 block and the multi-GPU runs compare their merged record with it
(verdict counts and the best candidate must be identical at 1/2/4/8 GPUs). ~1.7e10 pair checks: minutes to an hour of CPU."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import binding as B  # noqa: E402
from rapidquadcoptercollisiondetection_b200 import abi, workloads as W  # noqa: E402
from rapidquadcoptercollisiondetection_b200.api import merge_summaries  # noqa: E402


def main():
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    n, m, seed, obs_seed = 1 << log2n, 256, 5, 6
    specs = W.static_obstacles(obs_seed, m // 2, m // 2)
    obs = abi.make_obstacles(specs)
    orc = B.Oracle()
    chunk = 1 << 18
    parts, t0 = [], time.time()
    for lo in range(0, n, chunk):
        bc = W.synth_bc(seed, lo, min(chunk, n - lo))
        r = orc.generate_check(bc, obs, m, 0.002, index_base=lo, matrix=False, nthreads=threads)
        parts.append(r["summary"])
        if (lo // chunk) % 16 == 0:
            print(f"{lo + chunk}/{n} candidates, {time.time() - t0:.0f} s", flush=True)
    s = merge_summaries(parts)
    s["source"] = f"oracle/oracle.c on the CPU, 2^{log2n} candidates (seed {seed}) x {m} obstacles (seed {obs_seed}), min_t 0.002"
    out = os.path.join(ROOT, "tests", "golden", "c5_expected.json" if log2n == 26 else f"c5_expected_{log2n}.json")
    json.dump(s, open(out, "w"), indent=1)
    print(s)


if __name__ == "__main__":
    main()
