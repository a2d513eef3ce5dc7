"""Sums gpu__time_duration per kernel from an `ncu --csv --log-file` launch list. usage: launch_split.py file.csv"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.Counter(); cnt = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", "")); u = r[ix["Metric Unit"]]
    v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
    k = r[ix["Kernel Name"]][:70]; agg[k] += v; cnt[k] += 1
t = sum(agg.values())
for k, v in agg.most_common():
    print(f"{k:72s} n={cnt[k]:3d} {v:10.1f} us {100 * v / t:5.1f}%")
