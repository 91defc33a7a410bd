#!/usr/bin/env python
"""Basic-block cost profile of one kernel from an ncu report (--import-source on).

    python tools/ncu_blocks.py <report.ncu-rep> [min_share_percent]

Consecutive SASS instructions with the same execution count are one block; prints, in address order, every block whose
executed warp instructions are at least min_share_percent (default 0.5) of the kernel's: share, samples share, count,
instructions, opcode mix."""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
si, ii, so = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
inst = [r for r in rows[hi + 1:] if r and r[0].startswith("0x")]
tot_e = sum(int(r[ii] or 0) for r in inst)
tot_s = sum(int(r[si] or 0) for r in inst)
blocks = []
for r in inst:
    e, s, text = int(r[ii] or 0), int(r[si] or 0), r[so]
    op = text.split()[0] if not text.startswith("@") else text.split()[1]
    op = op.split(".")[0]
    if blocks and blocks[-1][0] == e:
        b = blocks[-1]
        b[1] += 1; b[2] += s; b[3][op] += 1
    else:
        blocks.append([e, 1, s, collections.Counter({op: 1}), r[0]])
print(f"total warp instructions {tot_e}, samples {tot_s}, SASS instructions {len(inst)}")
for e, n, s, ops, addr in blocks:
    share = 100.0 * e * n / max(tot_e, 1)
    if share >= min_share or 100.0 * s / max(tot_s, 1) >= 2 * min_share:
        mix = " ".join(f"{k}{v}" for k, v in ops.most_common(8))
        print(f"{addr[-5:]} inst {share:5.2f}%  smp {100.0 * s / max(tot_s, 1):5.2f}%  x{e:9d}  n={n:4d}  {mix}")
