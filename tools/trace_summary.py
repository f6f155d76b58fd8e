"""Summarise a chrome trace written by tools/step_profile.py: time per kernel name (and per grid with -g)."""
import json, sys, re, collections
d = json.load(open(sys.argv[1]))
by_grid = "-g" in sys.argv
pat = [a for a in sys.argv[2:] if a != "-g"]
ev = [e for e in d["traceEvents"] if e.get("cat") == "kernel"]
def short(n):
    n = n.replace("pub::(anonymous namespace)::", "").replace("pub::<unnamed>::", "")
    n = re.sub(r"^void ", "", n)
    return re.sub(r"\(.*", "", n)
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    n = short(e["name"])
    if pat and not any(p in n for p in pat):
        continue
    k = (n, tuple(e["args"].get("grid", []))) if by_grid else n
    agg[k][0] += 1; agg[k][1] += e["dur"]
tot = sum(v[1] for v in agg.values())
t0 = min(e["ts"] for e in ev); t1 = max(e["ts"] + e["dur"] for e in ev)
print(f"span {(t1 - t0) / 1e3:.3f} ms, kernel sum {tot / 1e3:.3f} ms, launches {sum(v[0] for v in agg.values())}")
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{us / 1e3:8.3f} ms {100 * us / tot:5.1f}%  x{c:<4d} avg {us / c:8.1f} us  {str(k)[:110]}")
