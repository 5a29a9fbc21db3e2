#!/usr/bin/env python
"""Stall samples / executed instructions of K1 aggregated by code region (markers are
comment strings searched in the current wfm_sample.cu, so line numbers follow edits).
    python tools/ncu_regions.py src.csv all.sass <kernel> [n_tiles]"""
import csv, re, sys, collections
src_csv, sass, kern = sys.argv[1:4]
n_tiles = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
cu = open(__file__.rsplit('/', 2)[0] + '/waveforms_b200/csrc/wfm_sample.cu').read().split('\n')
def find(marker, start=0):
    for i in range(start, len(cu)):
        if marker in cu[i]: return i + 1
    raise KeyError(marker)
marks = [('eval: head', 'Val<U> eval_unit('), ('eval: sincos rows', 'for (int i = 0; i < n_sc; ++i)'), ('eval: rot rows', 'for (int j = 0; j < n_child; ++j)'),
         ('eval: generic rows', '// -- every other basis function'), ('eval: terms', '  // -- terms'), ('eval: end', '// ---- pre-pass (once per program)')]
k0 = find('sample_kernel(const __grid_constant__')
marks2 = [('kernel: setup', 'sample_kernel(const __grid_constant__'), ('tile: prefetch + packet wait', '  for (; t < tile_end; advance()) {'),
          ('tile: header', 'const PacketHeader* __restrict__ h'), ('tile: wait store read', "// the previous tile's bulk store must have finished"),
          ('tile: fill', '// base fill'), ('tile: patches', '// ---- flat segments with their own value'),
          ('tile: unit loop', "// ---- the tile's ACTIVE samples"), ('tile: store', '    // ---- store ---'), ('kernel: end', '// complex128 output')]
bounds = [(n, find(m)) for n, m in marks] + [(n, find(m, k0 - 1)) for n, m in marks2]
def region(f, l):
    if f == 'wfm_sample.cu':
        name = 'helpers (asm wrappers, abscissa, fill_tile)'
        for n, b in bounds:
            if l >= b: name = n
        for (n, b), (n2, b2) in zip(bounds, bounds[1:]):
            if b <= l < b2: return n
        return name
    if f == 'wfm_math.cuh': return 'math: sincos_n' if l >= 94 else ('math: mul/add/sub' if l <= 12 else 'math: sincos scalar')
    return f
line_of = {}; cur = None; on = False
for ln in open(sass, errors='replace'):
    if ln.startswith('.text.'):
        on = ln.strip().rstrip(':') == '.text.' + kern; continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m: line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv))); hdr = rows[1]
ia, iex, ismp = (hdr.index(k) for k in ('Address', 'Instructions Executed', '# Samples'))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
base = int(rows[2][ia], 16)
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()]); T = [0, 0]
for r in rows[2:]:
    if len(r) < len(hdr): continue
    k = line_of.get(int(r[ia], 16) - base) or ('?', 0)
    a = agg[region(*k)]; a[0] += int(r[iex]); a[1] += int(r[ismp]); T[0] += int(r[iex]); T[1] += int(r[ismp])
    for i in stall_cols:
        v = int(r[i] or 0)
        if v: a[2][hdr[i][6:]] += v
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    st = ' '.join(f'{n}={v}' for n, v in a[2].most_common(4))
    print(f'{k:44s} samples {100*a[1]/T[1]:5.1f}%  inst {100*a[0]/T[0]:5.1f}% ({a[0]/n_tiles:7.1f}/tile)  {st}')
