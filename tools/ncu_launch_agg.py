"""Aggregate an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,...` launch list by kernel + grid."""
import collections
import csv
import io
import sys

txt = open(sys.argv[1]).read()
r = csv.DictReader(io.StringIO(txt[txt.index('"ID"'):]))
agg = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
for row in r:
    key = (row['Kernel Name'][:70], row['Grid Size'])
    v = float(row['Metric Value'].replace(',', ''))
    agg[key][row['Metric Name']] += v
    if row['Metric Name'] == 'gpu__time_duration.sum':
        cnt[key] += 1
tot = 0.0
for k, d in sorted(agg.items(), key=lambda kv: -kv[1]['gpu__time_duration.sum']):
    n = cnt[k]
    t = d['gpu__time_duration.sum'] / n / 1e3
    rd = d.get('dram__bytes_read.sum', 0.0) / n
    bw = rd / (t * 1e-6) / 1e9 if t else 0
    print(f"{k[0]:70s} {k[1]:16s} n={n:4d} avg {t:9.2f} us  dram read {rd / 1e6:9.2f} MB -> {bw:7.1f} GB/s  total {t * n:9.1f} us")
    tot += t * n
print(f"sum {tot:.1f} us")
