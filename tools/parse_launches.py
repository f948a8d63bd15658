"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: last call only."""
import csv, sys
def load(f):
    rows = []
    with open(f) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            val = float(r["Metric Value"].replace(",", ""))
            unit = r["Metric Unit"]
            val = val / 1000 if unit == "ns" else (val * 1000 if unit == "ms" else val)
            rows.append((r["Kernel Name"], val, r["Grid Size"], r["Block Size"]))
    return rows
if __name__ == "__main__":
    f = sys.argv[1]
    calls = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    rows = load(f)
    per = len(rows) // calls
    tot = 0
    for n, v, g, b in rows[-per:]:
        n = n.replace("unnamed>::", "").replace("void ", "")
        print(f"   {n[:58]:58s} {v:9.1f} us  grid {g} block {b}")
        tot += v
    print(f"   total of last call: {tot:.1f} us over {per} launches")
