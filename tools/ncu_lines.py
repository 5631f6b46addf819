#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source sass,cuda` dump by CUDA source line: executed warp
instructions, stall samples and shared-memory wavefronts per line.  Usage: ncu_lines.py dump.csv [top_n]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
agg = defaultdict(lambda: [0, 0, 0, 0, ''])
cur_file = ''
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r and r[0] == 'Line No':
        hdr = {k: i for i, k in enumerate(r)}
        # two "Source" columns: first = CUDA source text, second = SASS
        continue
    if hdr is None or len(r) < 10:
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    key = (cur_file, line)
    a = agg[key]
    def g(name):
        try:
            return int(float(r[hdr[name]]))
        except (KeyError, ValueError, IndexError):
            return 0
    a[0] += g('Instructions Executed')
    a[1] += g('# Samples')
    a[2] += g('L1 Wavefronts Shared')
    a[3] += g('L1 Wavefronts Shared Excessive')
    a[4] = r[1][:90]
tot_i = sum(a[0] for a in agg.values()) or 1
tot_s = sum(a[1] for a in agg.values()) or 1
print('total warp instructions %d, samples %d' % (tot_i, tot_s))
print('%-18s %6s %7s %7s %9s %9s  %s' % ('file:line', 'inst%', 'samp%', 'inst', 'smem_wf', 'excess', 'source'))
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print('%-18s %6.2f %7.2f %7d %9d %9d  %s' % ('%s:%d' % (f[:12], l), 100.0 * a[0] / tot_i, 100.0 * a[1] / tot_s,
                                                  a[0], a[2], a[3], a[4].strip()))
