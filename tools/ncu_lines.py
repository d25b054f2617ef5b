#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line (and per file)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ia = isamp = None
lines = []
fname = ''
for r in rows:
    if 'Instructions Executed' in r:
        ia = r.index('Instructions Executed'); isamp = r.index('# Samples'); continue
    if ia is None or len(r) <= ia:
        if len(r) >= 2 and ('.cu' in r[1] or '.h' in r[1]): fname = r[1].split('/')[-1]
        continue
    if r[0] != '' and r[ia].isdigit():
        lines.append((int(r[ia]), int(r[isamp]) if r[isamp].isdigit() else 0, fname, int(r[0]), r[1].strip()[:120]))
tot = sum(l[0] for l in lines); tots = max(1, sum(l[1] for l in lines))
print('total warp-inst', tot, 'samples', tots)
if len(sys.argv) > 3:   # ranges: file:lo-hi=label,...
    cats = collections.OrderedDict()
    spec = []
    for item in sys.argv[3].split(','):
        rng, label = item.split('=')
        f, lh = rng.split(':'); lo, hi = lh.split('-')
        spec.append((f, int(lo), int(hi), label))
    for l in lines:
        lab = 'other'
        for f, lo, hi, label in spec:
            if l[2] == f and lo <= l[3] <= hi: lab = label; break
        c = cats.setdefault(lab, [0, 0]); c[0] += l[0]; c[1] += l[1]
    for k, v in cats.items(): print(f"{k:14s} {100*v[0]/tot:5.1f}% inst {100*v[1]/tots:5.1f}% samples")
lines.sort(reverse=True)
for l in lines[:top]:
    print(f"{100*l[0]/tot:5.1f}% inst {100*l[1]/tots:5.1f}% samp {l[2]}:{l[3]:>4} {l[4]}")
