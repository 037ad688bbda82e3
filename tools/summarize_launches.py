"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name."""
import csv, sys, re, collections
path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = collections.defaultdict(lambda: [0.0, 0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].replace("(anonymous namespace)::", "")
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"<.*", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
    tot[name][0] += v; tot[name][1] += 1
total = sum(v[0] for v in tot.values())
print(f"total {total/1e3:.3f} ms over {sum(v[1] for v in tot.values())} launches")
for name, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{us/total*100:6.2f}%  {us/1e3:9.3f} ms  {n:6d} launches  {us/n:9.2f} us/launch  {name}")
