"""Top source lines (warp instructions executed, stall samples, active lanes) of one kernel from an ncu report.
usage: python tools/src_hot.py report.ncu-rep kernel-regex [N] [launch-skip]   (launch-skip: take ONE launch, the n-th of the report)"""
import csv, subprocess, sys, collections
rep, kre = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
import re
SKIP = ["--launch-skip", sys.argv[4], "--launch-count", "1"] if len(sys.argv) > 4 else []
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + SKIP,
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None; func = None; take = False
agg = collections.defaultdict(lambda: [0, 0, 0, ""])
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        take = re.search(kre, r[1]) is not None
        if take: func = r[1]
        continue
    if not take: continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0]:
        ie = hdr.index("Instructions Executed"); te = hdr.index("Thread Instructions Executed"); sm = hdr.index("# Samples")
        try:
            a = agg[(cur_file, int(r[0]))]
            a[0] += int(r[ie]); a[1] += int(r[te]); a[2] += int(r[sm]); a[3] = r[1].strip()
        except ValueError:
            pass
print(func)
tot = sum(a[0] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print(f"total warp-instr {tot}, samples {ts}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:N]:
    print(f"{f}:{ln:<5d} {100*a[0]/max(tot,1):5.1f}% instr  {100*a[2]/max(ts,1):5.1f}% samples  lanes {a[1]/max(a[0],1):5.1f}  | {a[3][:90]}")
