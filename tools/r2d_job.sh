#!/bin/bash
o=gpurun_out
python -m pytest tests/test_pairs.py -m gpu -q 2>&1 | tail -3
python tools/bench_configs.py --only cfg3,cfg3v > $o/r2d_cfg3_pair.json 2> $o/r2d_cfg3_pair.err; python -c "
import json; r=json.load(open('$o/r2d_cfg3_pair.json'))
for k in ('cfg3','cfg3v'): print(k, {x:r[k][x] for x in r[k] if x not in ('layout',)}, r[k]['layout'])
print(r['latency'])"
python tools/bench_configs.py --only cfg3 --no-pair > $o/r2d_cfg3_nopair.json 2>> $o/r2d_cfg3_pair.err; python -c "
import json; r=json.load(open('$o/r2d_cfg3_nopair.json'))
for k in ('cfg3',): print(k, {x:r[k][x] for x in r[k] if x not in ('layout',)}, r[k]['layout'])"
python bench.py --no-cpu --no-e2e > $o/r2d_bench.json 2> $o/r2d_bench.err; python -c "
import json; r=json.load(open('$o/r2d_bench.json')); print(r['value'], r['roofline']['frac'], r['kernel_layout'])"
