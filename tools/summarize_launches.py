"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel (template arguments kept)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0       # optional: number of training steps the list covers
lines = [l for l in open(path, newline="") if not l.startswith("==")]
tot = collections.defaultdict(lambda: [0.0, 0])
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    m = re.match(r"([\w:]+)(<[^(]*>)?\(", name)
    short = (m.group(1) + (m.group(2) or "") if m else name)[:72]
    v = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r.get("Metric Unit", "ns"), 1e-3)
    tot[short][0] += v
    tot[short][1] += 1
total = sum(v[0] for v in tot.values())
print(f"total {total / 1e3:.3f} ms over {sum(v[1] for v in tot.values())} launches ({total / 1e3 / steps:.3f} ms per step over {steps:g} steps)")
for name, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{us / total * 100:6.2f}%  {us / 1e3 / steps:8.3f} ms/step  {n:5d} launches  {us / n:8.2f} us/launch  {name}")
