#!/usr/bin/env python
"""Condense `nvcc -Xptxas=-v` output (stdin or a file) into one line per kernel: name, registers, spills, smem."""
import re
import subprocess
import sys

text = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
rows = []
cur = None
for line in text.splitlines():
    m = re.search(r"Compiling entry function '([^']+)'", line)
    if m:
        cur = {"name": m.group(1)}
        rows.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        cur["stack"], cur["st"], cur["ld"] = m.groups()
    m = re.search(r"Used (\d+) registers", line)
    if m:
        cur["regs"] = m.group(1)
        m2 = re.search(r"(\d+) bytes smem", line)
        cur["smem"] = m2.group(1) if m2 else "0"
names = subprocess.run(["c++filt"], input="\n".join(r["name"] for r in rows), capture_output=True, text=True).stdout.splitlines()
for r, n in zip(rows, names):
    n = re.sub(r"\(.*", "", n).replace("void ", "")
    print("%-64s regs %3s  stack %4s  spill st/ld %4s/%4s  static smem %s" % (
        n, r.get("regs", "?"), r.get("stack", "?"), r.get("st", "?"), r.get("ld", "?"), r.get("smem", "?")))
