#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel: joins `ncu --page source --csv` (SASS view, one row per
instruction) with the line table of the cubin (`nvdisasm -g`), because the CUDA-C view of the report carries no
metrics here.

    python tools/ncu_lines.py gpurun_out/prof_full.ncu-rep pair_sketch_kernel [libcfk.so] [top_n]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(so_path, kernel):
    """[(line_no, inlined_chain)] per instruction of `kernel`, in address order."""
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so_path)], cwd=tmp, check=True,
                       capture_output=True)
        text = ""  # one cubin per translation unit: the kernel lives in one of them
        for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
            text += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True,
                                   text=True).stdout
    out, inside, cur, src_file = [], False, None, None
    for ln in text.splitlines():
        if ln.startswith("//--------------------- .text."):
            inside = kernel in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "(.*?)", line (\d+)(.*)', ln)
        if m:
            chain = re.findall(r'line (\d+)', m.group(3))
            cur = (int(m.group(2)), tuple(int(c) for c in chain))
            src_file = src_file or m.group(1)
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            out.append(cur)
    return out, src_file


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    so = sys.argv[3] if len(sys.argv) > 3 else "centroflye_b200/libcfk.so"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    cols = rows[hdr]
    ie, ts, smp = cols.index("Instructions Executed"), cols.index("Thread Instructions Executed"), cols.index("# Samples")
    wf = cols.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in cols else None
    body = [r for r in rows[hdr + 1:] if len(r) == len(cols)]
    if "Kernel Name" in rows[0][0] and len([r for r in rows if r and r[0] == "Kernel Name"]) > 1:
        # several launches matched: keep the first
        n_first = next((i for i, r in enumerate(rows[hdr + 1:]) if r and r[0] == "Kernel Name"), len(body))
        body = body[:n_first]
    lines, src_file = sass_lines(so, kernel)
    if len(lines) != len(body):
        print(f"# warning: {len(body)} profiled instructions vs {len(lines)} in the cubin (rebuilt since?)")
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    tot = [0, 0, 0, 0]
    for r, li in zip(body, lines):
        key = li[0] if li else -1
        vals = [int(r[ie] or 0), int(r[ts] or 0), int(r[smp] or 0), int(r[wf] or 0) if wf is not None else 0]
        for j, v in enumerate(vals):
            agg[key][j] += v
            tot[j] += v
    src_path = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", os.path.basename(src_file or "cfk.cu"))
    src = open(src_path).read().splitlines()
    print(f"# {kernel}: {tot[0]:.3e} warp-instructions, {tot[1]:.3e} thread-instructions, {tot[2]} samples, "
          f"{tot[3]:.3e} shared wavefronts")
    print(f"# {'line':>5} {'inst%':>6} {'smp%':>6} {'thr/inst':>8} {'smem wf%':>8}  source")
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        s = src[key - 1].strip()[:110] if 0 < key <= len(src) else "?"
        print(f"  {key:5d} {100 * v[0] / max(tot[0], 1):6.2f} {100 * v[2] / max(tot[2], 1):6.2f} "
              f"{v[1] / max(v[0], 1):8.1f} {100 * v[3] / max(tot[3], 1):8.2f}  {s}")


if __name__ == "__main__":
    main()
