#!/usr/bin/env python
"""Print the SASS of given source lines with per-instruction executed counts.
    python tools/ncu_sass_lines.py src.csv all.sass <kernel> file:line [file:line ...]"""
import csv, re, sys
src_csv, sass, kern = sys.argv[1:4]
want = set()
for a in sys.argv[4:]:
    f, l = a.rsplit(':', 1)
    want.add((f, int(l)))
line_of = {}
cur, on = None, False
for ln in open(sass, errors='replace'):
    if ln.startswith('.text.'):
        on = ln.strip().rstrip(':') == '.text.' + kern
        continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m: line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, isrc, iex, ith, ismp = (hdr.index(k) for k in ('Address', 'Source', 'Instructions Executed', 'Thread Instructions Executed', '# Samples'))
base = int(rows[2][ia], 16)
for r in rows[2:]:
    if len(r) < len(hdr): continue
    off = int(r[ia], 16) - base
    if line_of.get(off) in want:
        print(f'{off:06x} {line_of[off][1]:>4} {int(r[iex]):>9} {int(r[ith])/max(int(r[iex]),1):5.1f} {int(r[ismp]):>5}  {r[isrc]}')
