"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: kernel time shares of the last full step."""
import collections, csv, re, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
r = csv.reader(lines); hdr = next(r)
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
rows = []
for row in r:
    if len(row) <= vi: continue
    v = float(row[vi].replace(',', '')); unit = row[ui]
    if unit == 'ns': v /= 1e3
    elif unit == 'ms': v *= 1e3
    name = re.sub(r'\(.*', '', row[ki]); name = re.sub(r'void |<unnamed>::|at::native::', '', name)[:64]
    rows.append((name, v))
idx = [i for i, (nm, _) in enumerate(rows) if nm.startswith('sgd_kernel')]
step = rows[idx[-2] + 1: idx[-1] + 1]
tot = collections.defaultdict(lambda: [0, 0.0])
for name, v in step:
    tot[name][0] += 1; tot[name][1] += v
s = sum(v[1] for v in tot.values())
print('one train step: %d launches, %.2f ms summed kernel time (serialised, cold cache: compare SHARES)' % (len(step), s / 1e3))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print('%7.2f ms %5.1f%% x%4d  %s' % (v[1] / 1e3, 100 * v[1] / s, v[0], k))
