"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count and total time
over the last 1/parts of the launches (one iteration of a probe that repeats its work `parts` times)."""
import csv, collections, sys
path = sys.argv[1]; parts = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
rows = rows[1:]
it = rows[-(len(rows) // parts):]
agg = collections.OrderedDict()
for r in it:
    v = float(r[vi].replace(',', ''))
    if r[ui] == 'ns': v /= 1e3
    a = agg.setdefault(r[ki][:70], [0, 0.0]); a[0] += 1; a[1] += v
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]): print(f"{a[1]:10.1f} us {a[0]:5d}  {k}")
print(f"{sum(a[1] for a in agg.values()):10.1f} us total, {len(it)} launches")
