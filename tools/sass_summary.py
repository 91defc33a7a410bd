#!/usr/bin/env python
"""sass_summary.py <library.so> <kernel-substring> -- what the compiler made of a kernel: instruction histogram by mnemonic
and the lines that show the data-movement instructions (UBLKCP = cp.async.bulk / TMA, SYNCS = mbarrier operations,
LDGSTS = cp.async, LDG / STG / LDS / STS widths).  Written to stdout; profiles/r2_sass_*.txt are its outputs."""
import collections
import re
import subprocess
import sys


def main():
    so, kernel = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    cur, lines = None, []
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            continue
        if cur and kernel in cur and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", ln):
            lines.append((cur, ln))
    if not lines:
        sys.exit(f"no kernel matching {kernel!r} in {so}")
    name = lines[0][0]
    ops = collections.Counter()
    shown = []
    for _, ln in lines:
        body = re.sub(r"/\*.*?\*/", "", ln).strip().rstrip(";")
        body = re.sub(r"^@!?U?P\d\s+", "", body)
        op = body.split()[0] if body else "?"
        ops[op.split(".")[0]] += 1
        if re.match(r"(UBLKCP|SYNCS|LDGSTS|UTMA|ELECT|FENCE|MEMBAR|ACQBULK|CCTL)", op):
            shown.append(ln.rstrip())
    print(f"kernel   {name}")
    print(f"library  {so}")
    print(f"SASS instructions: {len(lines)}")
    print("\n-- by mnemonic")
    for op, n in ops.most_common(40):
        print(f"{n:7d}  {op}")
    detail = collections.Counter()
    for _, ln in lines:
        body = re.sub(r"/\*.*?\*/", "", ln).strip()
        body = re.sub(r"^@!?U?P\d\s+", "", body)
        op = body.split()[0] if body else "?"
        if op.split(".")[0] in ("LDG", "STG", "LDS", "STS", "LDGSTS", "UBLKCP", "SYNCS", "ATOMG", "RED", "LD", "ST"):
            detail[op] += 1
    print("\n-- memory instructions by full opcode")
    for op, n in sorted(detail.items()):
        print(f"{n:7d}  {op}")
    print("\n-- bulk copy / mbarrier / cp.async / fence lines")
    for ln in shown:
        print(ln)


if __name__ == "__main__":
    main()
