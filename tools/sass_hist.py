"""Per-opcode histogram of executed warp instructions from an ncu report's SASS page.
usage: python tools/sass_hist.py report.ncu-rep [kernel-substring]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kern = None; hdr = None; data = collections.OrderedDict()
for r in rows:
    if r and r[0] == "Kernel Name":
        kern = r[1]; data[kern] = []; hdr = None; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if kern and hdr and len(r) == len(hdr):
        data[kern].append(dict(zip(hdr, r)))
for kern, ins in data.items():
    if want not in kern: continue
    tot = sum(int(i["Instructions Executed"]) for i in ins)
    thr = sum(int(i["Thread Instructions Executed"]) for i in ins)
    print(f"== {kern[:110]}\n   warp-instr {tot}  avg active threads {thr/max(tot,1):.2f}  static SASS {len(ins)}")
    by = collections.Counter(); byt = collections.Counter(); smp = collections.Counter()
    for i in ins:
        op = i["Source"].split()
        op = [t for t in op if not t.startswith("@")][0].split(".")[0]
        by[op] += int(i["Instructions Executed"]); byt[op] += int(i["Thread Instructions Executed"]); smp[op] += int(i["# Samples"])
    ts = sum(smp.values())
    for op, c in by.most_common(22):
        print(f"   {op:10s} {c:12d} {100*c/tot:5.1f}%  lanes {byt[op]/max(c,1):5.1f}  samples {100*smp[op]/max(ts,1):5.1f}%")
