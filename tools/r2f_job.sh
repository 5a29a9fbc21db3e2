#!/bin/bash
o=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --no-cpu --no-e2e --steps 100 > $o/r2f_bench.json 2> $o/r2f_bench.err; python -c "
import json; r=json.load(open('$o/r2f_bench.json')); print('cfg2', r['value'], r['roofline']['frac'])"
python tools/bench_configs.py --only cfg3,cfg4,cfg5 > $o/r2f_cfgs.json 2> $o/r2f_cfgs.err; python -c "
import json; r=json.load(open('$o/r2f_cfgs.json'))
for k in ('cfg3','cfg4','cfg5'): print(k, r[k]['GSa/s'], r[k]['roofline_frac'], r[k]['layout']['tile_samples'])
print(r['latency'])"
