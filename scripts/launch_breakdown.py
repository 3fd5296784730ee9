"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list into a per-kernel table for one forward pass."""
import collections, csv, re, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
seq = []
for row in csv.DictReader(lines):
    try: v = float(row['Metric Value'].replace(',', ''))
    except Exception: continue
    u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    nm = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('cmpc::', '')
    seq.append((nm, v))
idx = [i for i, (n, _) in enumerate(seq) if n.startswith('words_prepare')]
s, e = idx[0], idx[1]
agg = collections.OrderedDict(); tot = 0
for n, v in seq[s:e]:
    if 'at::' in n: n = 'torch fill/copy (plumbing)'
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print(f"one forward pass: {e - s} launches, {tot:.1f} us summed (ncu: cold-cache, serialised)")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t / tot * 100:6.2f}%  {t:9.1f} us  n={c:3d}  avg {t / c:8.1f} us  {k[:70]}")
