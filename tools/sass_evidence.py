#!/usr/bin/env python
"""Per kernel of the shipped library: counts of the Blackwell data-movement SASS instructions (cuobjdump -sass) and their
first occurrences.   python tools/sass_evidence.py > profiles/rNN_sass_tma_tmem.txt"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "powerfit_b200", "libpowerfit_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kernels, cur = [], None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = [m.group(1), []]
        kernels.append(cur)
    elif cur is not None and "/*" in line and ";" in line:
        cur[1].append(line.rstrip())
names = subprocess.run(["c++filt"], input="\n".join(k[0] for k in kernels), capture_output=True, text=True).stdout.splitlines()
KEYS = ["UTMALDG", "SYNCS", "LDTM", "STTM", "UTCATOMSWS", "LDGSTS", "FFMA2", "FADD2", "FMUL2", "LDS", "STS", "LDG", "STG", "BAR.SYNC"]
print("# SASS evidence (cuobjdump -sass powerfit_b200/libpowerfit_b200.so): per kernel, counts of the Blackwell data-movement")
print("# instructions (UTMALDG = cp.async.bulk.tensor load, SYNCS = mbarrier ops, LDTM/STTM = tcgen05.ld/st, UTCATOMSWS = tcgen05.alloc/dealloc,")
print("# LDGSTS = per-thread cp.async, FFMA2/FADD2/FMUL2 = packed f32x2) and the first occurrences.\n")
for (mangled, lines), name in zip(kernels, names):
    name = re.sub(r"\(.*", "", name)
    if not re.search(r"fused_|cls_", name):
        continue
    counts = {k: sum(1 for l in lines if re.search(r"\b" + re.escape(k) + r"\b", l.split("*/", 1)[1].split()[0] if False else l)) for k in KEYS}
    counts = {k: sum(1 for l in lines if re.search(r"\s" + re.escape(k) + r"[.\s]", l)) for k in KEYS}
    print("== " + name)
    print("   " + "  ".join("%s:%d" % (k, v) for k, v in counts.items() if v))
    for k in ("UTCATOMSWS", "SYNCS", "UTMALDG", "STTM", "LDTM"):
        first = next((l for l in lines if re.search(r"\s" + k + r"[.\s]", l)), None)
        if first:
            print("   " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", first.strip()))
    print()
