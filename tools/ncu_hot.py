#!/usr/bin/env python
"""Top stall sites of one kernel from an ncu report's source page (SASS view).
   python tools/ncu_hot.py rep.ncu-rep kernel_regex [N]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# first kernel only
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] == "Kernel Name":
        break
    body.append(r)
ci = {k: i for i, k in enumerate(hdr)}
S = ci["# Samples"]
tot = sum(int(r[S] or 0) for r in body)
stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(int(r[ci[k]] or 0) for r in body) for k in stall_cols}
print("total samples", tot)
print("by reason:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]))
order = sorted(range(len(body)), key=lambda i: -int(body[i][S] or 0))[:n]
for i in sorted(order):
    r = body[i]
    top = sorted(((int(r[ci[k]] or 0), k[6:]) for k in stall_cols), reverse=True)[:2]
    print("%5d %5.2f%%  %-70s %s" % (i, 100.0 * int(r[S] or 0) / tot, r[ci["Source"]][:70], " ".join("%s=%d" % (k, v) for v, k in top)))
