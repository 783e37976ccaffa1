#!/usr/bin/env python
"""Top stall sites of one kernel in an ncu report (SASS view): python tools/ncu_hotlines.py rep regex [n]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# several kernels may follow each other; take the first block
hdr = rows[1]
body = []
for r in rows[2:]:
    if len(r) != len(hdr) or r[0] == "Kernel Name":
        break
    body.append(r)
si = hdr.index("# Samples")
tot = sum(int(r[si]) for r in body)
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot, "instructions", len(body))
order = sorted(range(len(body)), key=lambda i: -int(body[i][si]))[:n]
for i in sorted(order):
    r = body[i]
    top = sorted(((int(r[j]), hdr[j][6:]) for j in stalls), reverse=True)[:2]
    print("%5d %5.1f%%  %-60s %s" % (i, 100.0 * int(r[si]) / tot, r[1].strip()[:60], " ".join("%s:%d" % (b, a) for a, b in top if a)))
