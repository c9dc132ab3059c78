#!/usr/bin/env python
"""Per-chunk table from an ncu launch list CSV (scripts/gpu_profile_g.sh): us per pass, threads/inst, issue, Minst."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; L = collections.OrderedDict()
for r in rows[hi + 1:]:
    d = dict(zip(hdr, r))
    L.setdefault(int(d['ID']), {'k': d['Kernel Name'].split('::')[-1][:14]})[d['Metric Name']] = float(d['Metric Value'].replace(',', ''))
chunk, cur = [], []
for i in sorted(L):
    cur.append(L[i])
    if L[i]['k'].startswith('resolve'): chunk.append(cur); cur = []
T = 'gpu__time_duration.sum'; tot = collections.Counter()
for c, ch in enumerate(chunk):
    tr = [x for x in ch if x['k'].startswith('trace')]; sh = [x for x in ch if x['k'].startswith('shadow')]
    for x in ch: tot[x['k'][:6]] += x[T]
    print(c, 'trace', [round(x[T] / 1e3) for x in tr[:4]], 'shadow', [round(x[T] / 1e3) for x in sh[:4]],
          'thr/inst %.1f %.1f' % (tr[0]['smsp__thread_inst_executed_per_inst_executed.ratio'], sh[0]['smsp__thread_inst_executed_per_inst_executed.ratio']),
          'issue %.0f %.0f' % (tr[0]['smsp__issue_active.avg.pct_of_peak_sustained_active'], sh[0]['smsp__issue_active.avg.pct_of_peak_sustained_active']),
          'Minst %d %d' % (tr[0]['smsp__inst_executed.sum'] / 1e6, sh[0]['smsp__inst_executed.sum'] / 1e6))
print({k: round(v / 1e6, 2) for k, v in tot.items()}, 'ms')
