"""Opcode histogram + stall summary + hottest SASS lines from an ncu source-page CSV."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
items = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 16
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]; hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
end = his[1] if len(his) > 1 else len(rows)
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, 'Instructions Executed') for r in data); ts = sum(f(r, '# Samples') for r in data)
print("SASS lines", len(data), "warp-instr per item", round(tot / items, 1), "samples", ts)
ops = collections.Counter(); st = collections.Counter()
for r in data:
    src = r[ix['Source']].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
    op = m.group(2).split('.')[0] if m else src[:10]
    ops[op] += f(r, 'Instructions Executed'); st[op] += f(r, '# Samples')
print("  ".join(f"{op}:{c/items:.0f}({100*st[op]/ts:.0f}%)" for op, c in ops.most_common(18)))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s_: sum(f(r, s_) for r in data) for s_ in stalls}
print(", ".join(f"{k[6:]}={100*v/ts:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * ts))
for r in sorted(data, key=lambda r: -f(r, '# Samples'))[:top]:
    dom = sorted(((f(r, s_), s_[6:]) for s_ in stalls), reverse=True)[:2]
    print(f"{100*f(r,'# Samples')/ts:6.2f}% exec={int(f(r,'Instructions Executed')):8d} {r[ix['Source']][:64]:64s} {dom[0][1]}:{int(dom[0][0])} {dom[1][1]}:{int(dom[1][0])}")
