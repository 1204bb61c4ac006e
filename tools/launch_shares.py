"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total and share.
usage: python tools/launch_shares.py profiles/r1_launches_bench.csv [--skip N] [--md]"""
import argparse, csv, collections, re
ap = argparse.ArgumentParser(); ap.add_argument("csv"); ap.add_argument("--skip", type=int, default=0); ap.add_argument("--md", action="store_true")
a = ap.parse_args()
rows = []
with open(a.csv) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((int(r["ID"]), r["Kernel Name"], float(r["Metric Value"]) / (1e3 if r["Metric Unit"] == "ns" else 1.0)))
rows = rows[a.skip:]
def short(n):
    n = n.replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    n = re.sub(r"\(.*", "", n); n = re.sub(r"(?<=[A-Za-z0-9_])<.*", "", n)
    return n.split("::")[-1][:48]
agg = collections.OrderedDict()
for _, n, us in rows:
    k = short(n); c = agg.setdefault(k, [0, 0.0]); c[0] += 1; c[1] += us
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot/1e3:.3f} ms total device time (serialised, cold)")
fmt = "| {:48s} | {:>6} | {:>10} | {:>6} |" if a.md else "{:48s} {:>6} {:>10} {:>6}"
print(fmt.format("kernel", "n", "total us", "share"))
if a.md: print("|---|---|---|---|")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(fmt.format(k, n, f"{us:.1f}", f"{100*us/tot:.1f}%"))
