"""Top stalled SASS instructions of one kernel from `ncu -i X.ncu-rep --page source --csv` output (first kernel block).
usage: ncu_src_top.py src.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hi = his[0]
end = his[1] - 1 if len(his) > 1 else len(rows)
hdr = rows[hi]; data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[ix['# Samples']]) for r in data)
print('total samples', tot, 'instructions', len(data))
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:n]:
    st = sorted(((s, int(r[ix[s]])) for s in stalls if int(r[ix[s]]) > 0), key=lambda kv: -kv[1])[:3]
    print(r[ix['# Samples']].rjust(6), r[ix['Address']][-5:], r[ix['Source']].strip()[:72].ljust(72), st)
