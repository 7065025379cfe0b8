#!/usr/bin/env python
"""tools/sass_by_line.py REPORT.ncu-rep LIB.so [KERNEL_SUBSTR[&SUBSTR..]] -- attribute the executed warp-instructions of an ncu
capture to source lines (nvdisasm -g line info of the SAME library build), split into FP64 / other, with the
average number of active threads. Instructions are matched by their order inside the kernel."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, lib = sys.argv[1], sys.argv[2]
    sub = sys.argv[3] if len(sys.argv) > 3 else "k_checkILi0ELb0ELb0"
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    dis = []  # every cubin of the library (one per kernel file and build); the kernel is found by its name
    for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
        dis += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
    # instructions of the kernel, each with the innermost (file, line)
    insts, cur, inside = [], ("?", 0), False
    for ln in dis:
        if ln.startswith(".text.") and ln.endswith(":"):
            # the first kernel whose name holds every '&'-separated part of KERNEL_SUBSTR, not its out-of-line helpers
            inside = all(part in ln for part in sub.split("&")) and "$" not in ln and not insts
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            insts.append((cur, m.group(2).strip()))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
    ix = {h: i for i, h in enumerate(rows[hi])}
    prof = []
    for r in rows[hi + 1:]:
        try:
            prof.append((r[ix["Source"]], float(r[ix["Instructions Executed"]]), float(r[ix["Thread Instructions Executed"]]),
                         float(r[ix["# Samples"]])))
        except Exception:
            pass
    if len(prof) != len(insts):
        print(f"warning: {len(prof)} profiled vs {len(insts)} disassembled instructions -- different builds?", file=sys.stderr)
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0.0])  # inst, fp64 inst, thread inst, samples
    for (loc, _), (src, ie, te, sm) in zip(insts, prof):
        t = src.split()
        op = t[1] if t[0].startswith("@") else t[0]
        a = agg[loc]
        a[0] += ie
        a[1] += ie if op[0] == "D" or op.startswith("MUFU") else 0.0
        a[2] += te
        a[3] += sm
    tot = sum(a[0] for a in agg.values()) or 1
    ts = sum(a[3] for a in agg.values()) or 1
    srcs = {}
    print(f"{'file:line':28s} {'inst%':>6s} {'fp64%':>6s} {'other%':>6s} {'thr':>5s} {'smpl%':>6s}  source")
    for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("TOP", "60"))]:
        f = loc[0]
        if f not in srcs:
            for root in ("rapidquadcoptercollisiondetection_b200/csrc", "/usr/local/cuda/include", "/usr/local/cuda/include/crt"):
                pth = os.path.join(root, f)
                if os.path.exists(pth):
                    srcs[f] = open(pth, errors="replace").read().splitlines()
                    break
            else:
                srcs[f] = []
        text = srcs[f][loc[1] - 1].strip()[:70] if 0 < loc[1] <= len(srcs[f]) else ""
        print(f"{f + ':' + str(loc[1]):28s} {a[0] / tot * 100:6.2f} {a[1] / tot * 100:6.2f} {(a[0] - a[1]) / tot * 100:6.2f} "
              f"{a[2] / max(a[0], 1):5.1f} {a[3] / ts * 100:6.2f}  {text}")


if __name__ == "__main__":
    main()
