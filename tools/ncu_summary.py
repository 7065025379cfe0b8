#!/usr/bin/env python
"""tools/ncu_summary.py REPORT.ncu-rep [out.md] -- the counters this repo's kernels are judged on, from an
`ncu --set full` capture: duration, warp execution efficiency, FP64 pipe utilisation, issue slots, DRAM
traffic, registers/occupancy, stall mix, and the per-block SASS hot spots."""
import collections
import csv
import io
import subprocess
import sys


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args, "--csv"], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    out = []
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw"))))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        out.append(f"## {d.get('Kernel Name', '?')}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        keys = [
            "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed.avg.per_cycle_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__bytes_read.sum.per_second", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
            "sm__cycles_elapsed.avg.per_second",
        ]
        for k in keys:
            if k in d:
                out.append(f"- {k} = {d[k]} {u[k]}")
        st = {k[len("smsp__pcsamp_warps_issue_stalled_"):]: float(v) for k, v in d.items()
              if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and v}
        tot = sum(st.values()) or 1
        out.append("- stall mix (pc samples): " + ", ".join(f"{k} {v / tot * 100:.1f}%" for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]))
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source"))))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows[hi + 1:]:
        try:
            data.append((r[ix["Source"]], float(r[ix["# Samples"]]), float(r[ix["Instructions Executed"]]),
                         float(r[ix["Thread Instructions Executed"]]), float(r[ix["stall_no_inst"]])))
        except Exception:
            pass
    ti = sum(d[2] for d in data) or 1
    ts = sum(d[1] for d in data) or 1
    out.append(f"\n### SASS: {len(data)} instructions, {ti:.3g} warp-instructions executed, avg active threads {sum(d[3] for d in data) / ti:.2f}")
    ops = collections.Counter()
    for d in data:
        t = d[0].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] += d[2]
    out.append("- opcode mix: " + ", ".join(f"{k} {v / ti * 100:.1f}%" for k, v in ops.most_common(14)))
    out.append("- by 250-instruction block: offset, share of executed instructions, avg threads, share of samples, no-instruction stall share")
    for i in range(0, len(data), 250):
        b = data[i:i + 250]
        ie = sum(d[2] for d in b)
        if ie / ti < 0.002:
            continue
        out.append(f"  - {i:5d}: inst {ie / ti * 100:5.1f}%  thr {sum(d[3] for d in b) / max(ie, 1):5.1f}  samples {sum(d[1] for d in b) / ts * 100:5.1f}%  no_inst {sum(d[4] for d in b) / max(sum(d[1] for d in b), 1) * 100:3.0f}%")
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
