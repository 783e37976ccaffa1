"""Share of stall samples and instructions per barrier-delimited phase of one kernel in an ncu report:
   python tools/ncu_phases.py rep.ncu-rep kernel_regex"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; body=[]
for r in rows[2:]:
    if len(r) != len(hdr) or r[0] == "Kernel Name": break
    body.append(r)
si = hdr.index("# Samples"); ie = hdr.index("Instructions Executed")
tot = sum(int(r[si]) for r in body); toti = sum(int(r[ie]) for r in body)
seg=0; acc=0; acci=0; start=0
for i,r in enumerate(body):
    acc += int(r[si]); acci += int(r[ie])
    if "BAR.SYNC" in r[1] or i == len(body)-1:
        print("phase %d: sass %d-%d  samples %.1f%%  instr %.1f%%" % (seg, start, i, 100.*acc/tot, 100.*acci/toti))
        seg += 1; acc = 0; acci = 0; start = i+1
print("total samples", tot, "warp instr", toti)
