"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the
last bench step (from its gray kernel to the end)."""
import collections, csv, re, sys

path = sys.argv[1]
rows = []
with open(path) as fh:
    lines = [l for l in fh if not l.startswith("==")]
for row in csv.DictReader(lines):
    if row.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((int(row["ID"]), row["Kernel Name"], float(row["Metric Value"].replace(",", ""))))
short = lambda n: re.sub(r"\(.*", "", n).split("::")[-1]
idx = [i for i, r in enumerate(rows) if "RgbLuma" in r[1]]
last = rows[idx[-1]:]
agg = collections.OrderedDict()
for _, n, t in last:
    a = agg.setdefault(short(n), [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(v[1] for v in agg.values())
print(f"# {path}: last step = {len(last)} launches, {tot / 1e3:.1f} us of kernel time (ncu: serialised, cold cache)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:42s} n={v[0]:4d} total={v[1] / 1e3:10.1f} us share={v[1] / tot * 100:5.1f}%")
