"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time of the LAST encode in the log."""
import csv, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = [r for r in csv.DictReader(lines) if r['Metric Name'] == 'gpu__time_duration.sum']
names = [(r['Kernel Name'].split('(')[0].replace('void ', ''), float(r['Metric Value'].replace(',', ''))) for r in rows]
# the last encode starts at the last histogram launch (preceded by the plane conversion for RGB)
start = max(i for i, (n, _) in enumerate(names) if 'k_hist' in n)
if start and 'k_to_planes' in names[start - 1][0]:
    start -= 1
tot = 0
for n, v in names[start:]:
    print(f"{n:28s} {v / 1000:9.1f} us")
    tot += v
print(f"{'total':28s} {tot / 1000:9.1f} us")
