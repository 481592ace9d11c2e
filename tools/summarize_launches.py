"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (dev tool).
usage: summarize_launches.py launches.csv [first_launch_id] [n_launches]"""
import collections, csv, re, sys
path = sys.argv[1]
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
count = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
lines = [l for l in open(path) if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines)]
rows = [r for r in rows if first <= int(r["ID"]) < first + count]
agg, tot = collections.OrderedDict(), 0.0
for r in rows:
    v = float(r["Metric Value"].replace(",", ""))
    v = v / 1000 if r["Metric Unit"] == "ns" else (v * 1000 if r["Metric Unit"] == "ms" else v)
    name = re.sub(r"^void (owl::)?", "", r["Kernel Name"])
    name = re.sub(r"\(CUtensorMap_st.*", "", name)
    name = re.sub(r"\((const |float|int|long|void|__half|unsigned).*", "", name).replace("(bool)", "").replace("(int)", "")
    k = name[:100]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print(f"{len(rows)} launches, {tot:.1f} us total (ncu serialised, cold caches: compare shares)")
print(f"{'total us':>10} {'n':>4} {'avg us':>8} {'share':>6}  kernel")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} {n:4d} {t / n:8.1f} {100 * t / tot:5.1f}%  {k}")
