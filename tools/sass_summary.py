#!/usr/bin/env python
"""Per-kernel SASS summary of libcfk.so (cuobjdump -sass): instruction count and the mnemonics that show how data moves.
    python tools/sass_summary.py [centroflye_b200/libcfk.so] > profiles/rNN_sass_summary.txt"""
import collections
import re
import subprocess
import sys

WATCH = ["UBLKCP", "UTMALDG", "SYNCS", "CCTL", "LDGSTS", "LDG", "STG", "LDS", "STS", "ATOMS", "ATOMG", "ATOM", "REDG", "BAR",
         "SHFL", "VOTE", "MATCH", "REDUX", "IMAD", "LOP3", "SHF", "POPC", "BREV", "FLO", "LDL", "STL", "HMMA", "UTC"]


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else "centroflye_b200/libcfk.so"
    text = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for ln in text.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0]
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)", ln)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            base = op.split(".")[0]
            for w in WATCH:
                if base == w or (w in ("UTC",) and base.startswith(w)):
                    cur[w] += 1
            if base == "CCTL" or "PREFETCH" in op:
                cur["prefetch/CCTL"] += 1
    print(f"# {so}: SASS instruction counts per kernel (static; sm_100a).  UBLKCP = cp.async.bulk (TMA engine), SYNCS = mbarrier,")
    print("# ATOMS / ATOMG / REDG = shared / global atomics (REDG: no return value), LDL / STL = local-memory (spill) traffic, HMMA / UTC* = tensor cores (none: no contraction here)")
    cols = ["total", "UBLKCP", "SYNCS", "LDG", "STG", "LDS", "STS", "ATOMS", "ATOMG", "REDG", "BAR", "SHFL", "VOTE", "MATCH", "REDUX",
            "POPC", "BREV", "CCTL", "LDL", "STL", "HMMA", "UTC"]
    print(f"{'kernel':46s} " + " ".join(f"{c:>6s}" for c in cols))
    for name, c in kernels.items():
        print(f"{name[:46]:46s} " + " ".join(f"{c.get(col, 0):6d}" for col in cols))


if __name__ == "__main__":
    main()
