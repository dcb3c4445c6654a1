"""Top stall lines of an `ncu --page source --csv` dump (SASS view): python tools/ncu_source_top.py X_source.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hi]
C = {h: i for i, h in enumerate(H)}
data = [r for r in rows[hi + 1:] if len(r) > C["# Samples"]]
tot = sum(int(r[C["# Samples"]] or 0) for r in data)
stalls = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
print(f"{len(data)} SASS instructions, {tot} samples")
agg = {s: sum(int(r[C[s]] or 0) for r in data) for s in stalls}
print("stall mix:", ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(data)), key=lambda i: -int(data[i][C["# Samples"]] or 0))[:n]
for i in sorted(order):
    r = data[i]
    s = int(r[C["# Samples"]] or 0)
    top = sorted(((int(r[C[k]] or 0), k[6:]) for k in stalls), reverse=True)[:2]
    print(f"{i:5d} {100 * s / max(tot, 1):5.1f}%  x{r[C['Instructions Executed']]:>9s}  {r[C['Source']].strip()[:70]:70s} {top[0][1]}:{top[0][0]} {top[1][1]}:{top[1][0]}")
