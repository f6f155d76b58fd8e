"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name over the
last N launches (one training step).  usage: summarize_launches.py launches.csv [launches_per_step]"""
import csv, sys, re, collections
path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    rows.append((int(r["ID"]), r["Kernel Name"], ns))
n = int(sys.argv[2]) if len(sys.argv) > 2 else len(rows)
rows = rows[-n:]
tot = sum(r[2] for r in rows)
agg = collections.defaultdict(lambda: [0, 0.0])
for _, name, ns in rows:
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void ", "", short)
    short = short.replace("pub::<unnamed>::", "").replace("pub::(anonymous namespace)::", "")
    agg[short][0] += 1; agg[short][1] += ns
print(f"launches {len(rows)}  total {tot/1e6:.3f} ms (serialised, cold-cache: compare SHARES)")
for name, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{ns/1e6:9.3f} ms  {100*ns/tot:5.1f}%  x{c:<5d} {name[:110]}")
