#!/usr/bin/env python
"""tools/sass_histogram.py LIB.so [KERNEL_SUBSTR] -- static evidence from the SHIPPED cubin (cuobjdump -sass): per kernel
whose name contains KERNEL_SUBSTR (default k_pool) the instruction count, the opcode histogram, and the TMA / mbarrier
lines (UBLKCP = cp.async.bulk global->shared, SYNCS.* = mbarrier ops) with their addresses."""
import collections
import re
import subprocess
import sys


def main():
    lib = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else "k_pool"
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    arch = re.findall(r"arch = (sm_\w+)", txt)
    print(f"# {lib}: cubin arch {sorted(set(arch))}")
    cur, hist, tma, n = None, None, None, 0

    def flush():
        if cur is None or sub not in cur:
            return
        name = subprocess.run(["c++filt", cur], capture_output=True, text=True).stdout.strip()
        print(f"\n## {name}\n- {n} SASS instructions")
        top = ", ".join(f"{k} {v}" for k, v in hist.most_common(28))
        print(f"- opcodes: {top}")
        f64 = sum(v for k, v in hist.items() if k in ("DMUL", "DADD", "DFMA"))
        print(f"- DMUL+DADD+DFMA = {f64} ({100.0 * f64 / max(n, 1):.1f} %); tensor-core opcodes (HMMA/UTCMMA): "
              f"{sum(v for k, v in hist.items() if 'MMA' in k)}")
        for line in tma:
            print(f"- {line}")

    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush()
            cur, hist, tma, n = m.group(1), collections.Counter(), [], 0
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(2).split(".")[0]
            hist[op] += 1
            n += 1
            if op in ("UBLKCP", "SYNCS", "UTMALDG"):
                tma.append(f"/*{m.group(1)}*/ " + line.split("*/", 1)[1].split(";")[0].strip())
    flush()


if __name__ == "__main__":
    main()
