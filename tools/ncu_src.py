"""Summarise an ncu source-page CSV: top stalled SASS lines + totals per stall reason.
   ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > f.csv ; python tools/ncu_src.py f.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
# find header row
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot)
agg = {s: sum(f(r, s) for r in data) for s in stalls}
print("by reason:", ", ".join(f"{k[6:]}={100*v/max(tot,1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.005 * tot))
print(f"{'samples%':>8} {'exec':>9}  source / dominant stalls")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:top]:
    dom = sorted(((f(r, s), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{100*f(r,'# Samples')/max(tot,1):8.2f} {int(f(r,'Instructions Executed')):9d}  {r[ix['Source']][:90]:90s} {dom[0][1]}:{int(dom[0][0])} {dom[1][1]}:{int(dom[1][0])}")
