"""Dynamic SASS profile of one launch from an ncu report: warp instructions executed and stall samples per opcode
and per source line (every SASS instruction counted ONCE, attributed to the innermost source line ncu lists for it).
usage: python tools/sass_dyn.py report.ncu-rep kernel-regex launch-skip [N]"""
import collections
import csv
import re
import subprocess
import sys

rep, kre, skip = sys.argv[1], sys.argv[2], sys.argv[3]
N = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
by_op = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
lines = []
for r in rows:
    if not r:
        continue
    if r[0] in ("Address", "Line No") or (hdr is None and "Instructions Executed" in r):
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    try:
        ie = int(r[hdr.index("Instructions Executed")])
        te = int(r[hdr.index("Thread Instructions Executed")])
        sm = int(r[hdr.index("# Samples")])
    except ValueError:
        continue
    src = r[hdr.index("Source")]
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2).split(".")[0] if m else "?"
    by_op[op][0] += ie
    by_op[op][1] += te
    by_op[op][2] += sm
    tot[0] += ie
    tot[1] += te
    tot[2] += sm
    lines.append((ie, te, sm, src))
print(f"total warp-instr {tot[0]}  thread-instr {tot[1]}  lanes {tot[1]/max(tot[0],1):.1f}  samples {tot[2]}")
for op, a in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:N]:
    print(f"{op:10s} {100*a[0]/max(tot[0],1):5.1f}% instr  {100*a[2]/max(tot[2],1):5.1f}% samples  lanes {a[1]/max(a[0],1):5.1f}")
