"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list:
time (and DRAM traffic, when captured) per kernel name over ONE training step = the launches between two consecutive
adamw_kernel launches (or, with --per-step N, the N launches ending at the one adamw launch captured).
usage: summarize_launches.py launches.csv [--per-step N] [--json out.json]"""
import csv, sys, re, collections, json
path = sys.argv[1]
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
per = collections.OrderedDict()          # launch id -> {name, ns, rd, wr}
UNIT = {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in csv.DictReader(lines):
    m = r.get("Metric Name")
    if m not in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"):
        continue
    d = per.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "ns": 0.0, "rd": None, "wr": None})
    v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r.get("Metric Unit", ""), 1)
    if m == "gpu__time_duration.sum": d["ns"] = v
    elif m == "dram__bytes_read.sum": d["rd"] = v
    else: d["wr"] = v
rows = list(per.values())
ad = [i for i, d in enumerate(rows) if "adamw_kernel" in d["name"]]
if len(ad) >= 2:
    rows = rows[ad[-2] + 1: ad[-1] + 1]        # exactly one step
elif len(ad) == 1 and "--per-step" in sys.argv:  # window ending at the optimizer launch
    n = int(sys.argv[sys.argv.index("--per-step") + 1])
    head = rows[ad[0] + 1:][:n]                     # the beginning of the NEXT step (steps are structurally identical)
    need = n - len(head)                            # ... completed by the end of this one
    rows = head + rows[max(0, ad[0] - need + 1): ad[0] + 1]
def short(n):
    n = n.replace("pub::<unnamed>::", "").replace("pub::(anonymous namespace)::", "")
    return re.sub(r"\(.*", "", re.sub(r"^void ", "", n))
tot = sum(d["ns"] for d in rows)
has_dram = any(d["rd"] is not None for d in rows)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for d in rows:
    a = agg[short(d["name"])]
    a[0] += 1; a[1] += d["ns"]; a[2] += (d["rd"] or 0) + (d["wr"] or 0)
print(f"launches {len(rows)}  total {tot/1e6:.3f} ms (serialised, cold-cache: compare SHARES)"
      + (f"  dram traffic {sum(a[2] for a in agg.values())/1e9:.2f} GB" if has_dram else ""))
for name, (c, ns, by) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    extra = f"  {by/1e6:9.1f} MB  {by/ns:6.2f} GB/ms" if has_dram and ns else ""
    print(f"{ns/1e6:9.3f} ms  {100*ns/tot:5.1f}%  x{c:<5d} {name[:70]:70s}{extra}")
fam = [v for k, v in agg.items() if re.search(r"conv_tc_kernel|conv_halo_kernel|wgrad_tc_kernel", k)]
out = {"launches": len(rows), "total_ms": tot / 1e6,
       "conv_family": {"launches": sum(v[0] for v in fam), "ms": sum(v[1] for v in fam) / 1e6,
                       "dram_bytes": sum(v[2] for v in fam) if has_dram else None,
                       "share_of_step": sum(v[1] for v in fam) / tot if tot else None}}
print(json.dumps(out))
if "--json" in sys.argv:
    json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
