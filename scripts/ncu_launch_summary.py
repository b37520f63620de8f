"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = None
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr = r; rows = rows[i + 1:]; break
ik, iv, im, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name'), hdr.index('Metric Unit')
agg = collections.defaultdict(list)
scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3}
for r in rows:
    if r[im] == 'gpu__time_duration.sum':
        agg[r[ik][:70]].append(float(r[iv].replace(',', '')) * scale.get(r[iu], 1.0))
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':70s} {'n':>5s} {'total_us':>11s} {'avg_us':>9s} {'share':>6s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:70s} {len(v):5d} {sum(v):11.1f} {sum(v)/len(v):9.1f} {100*sum(v)/tot:5.1f}%")
