"""Summarise an `ncu --metrics gpu__time_duration.sum --csv [--nvtx]` launch list:
per NVTX range (if present) the kernels by total time."""
import collections, csv, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
lines = [l for l in open(path) if not l.startswith('==')]
rd = csv.DictReader(lines)
nv = [f for f in rd.fieldnames if 'Push/Pop' in f]
groups = collections.OrderedDict()
for row in rd:
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v = v / 1000 if u == 'ns' else (v * 1000 if u == 'ms' else v)
    rng = row[nv[0]].split(':')[1] + ':' + ':'.join(row[nv[0]].split(':')[-3:]) if nv and row[nv[0]] else '-'
    if nv and row[nv[0]]:
        rng = row[nv[0]]
    k = row['Kernel Name']
    for pre in ('void ', 'at::native::', '(anonymous namespace)::'):
        k = k.replace(pre, '')
    k = k[:80]
    g = groups.setdefault(rng, collections.OrderedDict())
    a = g.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
for rng, g in groups.items():
    tot = sum(t for _, t in g.values())
    n = sum(c for c, _ in g.values())
    print(f'== {rng[-60:]}: {n} launches, {tot:.1f} us')
    for k, (c, t) in sorted(g.items(), key=lambda x: -x[1][1])[:top]:
        print(f'   {c:5d} {t:10.1f} us {t / c:9.1f}  {k}')
