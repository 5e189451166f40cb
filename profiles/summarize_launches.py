"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for row in csv.DictReader(lines):
    try: v = float(row['Metric Value'].replace(',', ''))
    except Exception: continue
    u = row['Metric Unit']
    v = v / 1e6 if u in ('ns', 'nsecond') else v / 1e3 if u in ('us', 'usecond') else v * 1e3 if u == 'second' else v
    name = re.sub(r'^void ', '', row['Kernel Name']); name = re.sub(r'<unnamed>::', '', name)
    short = re.sub(r'\(.*', '', name)[:78]
    agg[short][0] += 1; agg[short][1] += v; tot += v
print(f"total {tot:.2f} ms over {sum(a[0] for a in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{t:9.2f} ms {100*t/tot:5.1f}% {n:6d}  {k}")
