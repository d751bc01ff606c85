#!/usr/bin/env python
"""Opcode mix of an `ncu --page source --csv` dump, normalised per evaluated (position x warp)."""
import csv, collections, sys
path, npos = sys.argv[1], float(sys.argv[2])
rows = list(csv.reader(open(path)))
hdr, data = rows[1], rows[2:]
ie, isrc = hdr.index('Instructions Executed'), hdr.index('Source')
byop = collections.Counter(); tot = 0
for r in data:
    n = int(r[ie]); s = r[isrc].strip().split()
    op = s[1] if s[0].startswith('@') else s[0]
    byop[op.split('.')[0]] += n; tot += n
P = npos / 32
print(f'total warp-instructions {tot}  per position-warp {tot / P:.1f}')
for op, n in byop.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 22):
    print(f'{op:14s} {n / P:8.2f}')
