"""Summarise an ncu launch-list CSV: the last complete step's kernels and their share."""
import csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = [(r['Kernel Name'], float(r['Metric Value'].replace(',', '')) / 1000.0, r.get('Grid Size', '')) for r in csv.DictReader(lines)]
idx = [i for i, r in enumerate(rows) if 'step_begin' in r[0]]
a, b = idx[-2], idx[-1]
step = rows[a:b]
tot = sum(v for n, v, g in step if 'FillFunctor' not in n)
print(f"# kernels in one step: {len([1 for n,_,_ in step if 'FillFunctor' not in n])}, sum of durations {tot:.1f} us (cold cache, serialised)")
for n, v, g in step:
    if 'FillFunctor' in n: continue
    short = n.split('(')[0].replace('b200nn::', '').replace('void ', '')
    print(f"{v:8.2f} us {100*v/tot:5.1f}%  {short[:70]:70s} grid={g}")
