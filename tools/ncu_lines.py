#!/usr/bin/env python
"""Aggregate ncu warp-stall samples per CUDA source line.

    python tools/ncu_lines.py <report.ncu-rep> <library.so> <kernel-substring> [top N]

ncu's CSV source page lists SASS instructions in order; `nvdisasm -g` on the cubin of the same build lists the
same instructions with `//## File ..., line N` annotations.  The two are joined by instruction order.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(so, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
    out = []
    for cub in sorted(os.listdir(tmp)):
        txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
        in_k = False
        cur = ("?", 0)
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                in_k = kernel in m.group(1)
                continue
            if ln.startswith("\t.section") or ln.startswith(".section"):
                in_k = False
            if not in_k:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                if "inlined at" not in ln or True:
                    cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                out.append((int(m.group(1), 16), m.group(2).strip(), cur))
    return out


def main():
    rep, so, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    si = hdr.index("# Samples")
    ii = hdr.index("Instructions Executed")
    wi = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
    wx = hdr.index("L1 Wavefronts Shared Excessive") if "L1 Wavefronts Shared Excessive" in hdr else None
    inst = [r for r in rows[hdr_i + 1:] if r and r[0].startswith("0x")]
    sass = sass_lines(so, kernel)
    if len(sass) != len(inst):
        print(f"warning: {len(sass)} SASS instructions in the cubin vs {len(inst)} in the report", file=sys.stderr)
    agg = {}
    total = 0
    for (addr, text, loc), r in zip(sass, inst):
        s = int(r[si] or 0)
        e = int(r[ii] or 0)
        a = agg.setdefault(loc, [0, 0, 0, 0])
        a[0] += s
        a[1] += e
        if wi is not None:
            a[2] += int(r[wi] or 0)
            a[3] += int(r[wx] or 0)
        total += s
    srcs = {}
    by = 2 if os.environ.get("NCU_LINES_SORT") == "smem" else 0
    print(f"total samples {total}, shared-memory wavefronts {sum(a[2] for a in agg.values())} (excessive {sum(a[3] for a in agg.values())})")
    for loc, (s, e, w, x) in sorted(agg.items(), key=lambda kv: -kv[1][by])[:top]:
        f, n = loc
        if f not in srcs:
            for root in (os.path.join(os.path.dirname(os.path.abspath(so)), "csrc"), "."):
                p = os.path.join(root, f)
                if os.path.exists(p):
                    srcs[f] = open(p).read().splitlines()
                    break
            else:
                srcs[f] = []
        code = srcs[f][n - 1].strip() if 0 < n <= len(srcs[f]) else ""
        print(f"{100.0 * s / max(total, 1):6.2f}%  {s:7d} smp {e:10d} inst {w:10d} wf {x:9d} xs  {f}:{n}  {code[:100]}")


if __name__ == "__main__":
    main()
