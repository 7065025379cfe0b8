#!/usr/bin/env python
"""tools/sass_regions.py REPORT.ncu-rep LIB.so KERNEL_SUBSTR [kernel_file.cu] -- attribute the executed
warp-instructions (and stall samples) of an ncu capture to the OUTERMOST source line inside the kernel's own file
(nvdisasm -gi inline chains of the SAME library build), i.e. to the statement of the kernel body that caused them,
however deeply the device functions it calls were inlined. Also prints the static size of the hot code: the number
of SASS instructions that were executed at least once per `HOT` (default 1e-4) of all executed instructions --
the working set the instruction caches have to hold (L1.5 = 32 KB = 2048 instructions on Blackwell).
Instructions are matched by their order inside the kernel."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, lib, sub = sys.argv[1], sys.argv[2], sys.argv[3]
    kfile = sys.argv[4] if len(sys.argv) > 4 else "rqcd_pool.cu"
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    insts = []
    for cubin in sorted(os.listdir(tmp)):
        if not cubin.endswith(".cubin") or "-" in cubin:
            continue
        dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
        cur, inner, inside = ("?", 0), ("?", 0), False
        pending = []
        for ln in dis:
            if ln.startswith(".text.") and ln.endswith(":"):
                inside = sub in ln
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
            if m:
                pending.append((os.path.basename(m.group(1)), int(m.group(2))))
                continue
            m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
            if m:
                if pending:
                    inner = pending[0]
                    outer = [x for x in pending if x[0] == kfile]
                    cur = outer[-1] if outer else pending[-1]
                    pending = []
                insts.append((cur, inner, m.group(2).strip()))
        if insts:
            break
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
    ix = {h: i for i, h in enumerate(rows[hi])}
    prof = []
    for r in rows[hi + 1:]:
        try:
            prof.append((r[ix["Source"]], float(r[ix["Instructions Executed"]]), float(r[ix["Thread Instructions Executed"]]),
                         float(r[ix["# Samples"]]), float(r[ix["stall_no_inst"]])))
        except Exception:
            pass
    if len(prof) != len(insts):
        print(f"warning: {len(prof)} profiled vs {len(insts)} disassembled instructions -- different builds?", file=sys.stderr)
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0.0, 0.0, 0])  # inst, fp64, thread inst, samples, no_inst, static
    tot = sum(p[1] for p in prof) or 1
    hot_thr = float(os.environ.get("HOT", "1e-4")) * tot / max(len(prof), 1) * 10
    hot = 0
    for (loc, _inner, _), (src, ie, te, sm, ni) in zip(insts, prof):
        t = src.split()
        op = t[1] if t[0].startswith("@") else t[0]
        a = agg[loc]
        a[0] += ie
        a[1] += ie if op[0] == "D" or op.startswith("MUFU") else 0.0
        a[2] += te
        a[3] += sm
        a[4] += ni
        a[5] += 1
        hot += ie > hot_thr
    ts = sum(a[3] for a in agg.values()) or 1
    print(f"{len(prof)} SASS instructions, {tot:.4g} executed warp-instructions; hot static code: {hot} instructions = {hot * 16 / 1024:.1f} KB")
    srcs = {}
    print(f"{'file:line':24s} {'static':>6s} {'inst%':>6s} {'fp64%':>6s} {'thr':>5s} {'smpl%':>6s} {'noinst':>6s}  source")
    for loc, a in sorted(agg.items(), key=lambda kv: (kv[0][0] != kfile, kv[0][1])):
        if a[0] / tot < 0.002:
            continue
        f = loc[0]
        if f not in srcs:
            pth = os.path.join("rapidquadcoptercollisiondetection_b200/csrc", f)
            srcs[f] = open(pth, errors="replace").read().splitlines() if os.path.exists(pth) else []
        text = srcs[f][loc[1] - 1].strip()[:80] if 0 < loc[1] <= len(srcs[f]) else ""
        print(f"{f + ':' + str(loc[1]):24s} {a[5]:6d} {a[0] / tot * 100:6.2f} {a[1] / tot * 100:6.2f} {a[2] / max(a[0], 1):5.1f} "
              f"{a[3] / ts * 100:6.2f} {a[4] / max(a[3], 1) * 100:5.0f}%  {text}")


if __name__ == "__main__":
    main()
