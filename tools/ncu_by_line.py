#!/usr/bin/env python
"""Join an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) with
nvdisasm -g -c line info of the same cubin and aggregate executed instructions
and stall samples per CUDA source line.

    python tools/ncu_by_line.py src.csv all.sass <mangled kernel name> [top]
"""
import csv, re, sys, collections

src_csv, sass, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# line table of the kernel
line_of = {}
cur, on = None, False
for ln in open(sass, errors='replace'):
    if ln.startswith('.text.'):
        on = ln.strip().rstrip(':') == '.text.' + kern
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, isrc = hdr.index('Address'), hdr.index('Source')
iex, ith = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
ismp = hdr.index('# Samples')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
base = int(rows[2][ia], 16)
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
tot = [0, 0, 0]
opc = collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    off = int(r[ia], 16) - base
    key = line_of.get(off, (('?', 0), ''))[0]
    ex, th, sm = int(r[iex]), int(r[ith]), int(r[ismp])
    a = agg[key]
    a[0] += ex; a[1] += th; a[2] += sm
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            a[3][hdr[i]] += v
    tot[0] += ex; tot[1] += th; tot[2] += sm
    opc[r[isrc].split()[0] if not r[isrc].startswith('@') else r[isrc].split()[1]] += ex
print(f'total warp-instructions {tot[0]}  thread-instr {tot[1]}  samples {tot[2]}')
print('--- by source line (sorted by executed warp instructions)')
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ' '.join(f'{k[6:]}={v}' for k, v in a[3].most_common(3))
    print(f'{str(key[0]):>22}:{key[1]:<5} inst {a[0]:>10} {100*a[0]/tot[0]:5.1f}%  samples {a[2]:>7} {100*a[2]/max(tot[2],1):5.1f}%  {st}')
print('--- by opcode')
for k, v in opc.most_common(25):
    print(f'{k:>14} {v:>10} {100*v/tot[0]:5.1f}%')
