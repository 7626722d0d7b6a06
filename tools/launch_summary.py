"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum[,...] --csv) by kernel: launches, total ms, share.
  python tools/launch_summary.py gpurun_out/<tag>_launches.csv > profiles/<tag>_launch_summary_bench.csv"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
r = csv.reader(lines)
hdr = next(r)
ik, im, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for row in r:
    if len(row) <= iv or not row[im].startswith('gpu__time_duration'):
        continue
    v = float(row[iv].replace(',', ''))
    ms = v * {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0}[row[iu]]
    k = row[ik].split('(')[0][:70]
    agg[k][0] += 1
    agg[k][1] += ms
tot = sum(v[1] for v in agg.values())
print('kernel,launches,total_ms,share')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f'"{k}",{v[0]},{v[1]:.3f},{v[1] / tot:.4f}')
