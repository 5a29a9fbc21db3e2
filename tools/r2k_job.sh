#!/bin/bash
o=gpurun_out
python -m pytest tests/test_builder.py -m gpu -q 2>&1 | tail -2
python tools/bench_configs.py --only cfg3,cfg3v > $o/r2k_cfg3.json 2> $o/r2k_cfg3.err; python -c "
import json; r=json.load(open('$o/r2k_cfg3.json'))
for k in ('cfg3','cfg3v'): print(k, r[k]['GSa/s'], r[k]['ms'], r[k]['layout'])"
WFM_K1_UNIT=2 python tools/bench_configs.py --only cfg3 2>/dev/null | python -c "
import json,sys; r=json.load(sys.stdin); print('unit2', r['cfg3']['GSa/s'], r['cfg3']['layout'])"
ncu --set full --clock-control none --import-source on -k regex:sample_dense -s 2 -c 1 -o $o/r2k_dense python tools/bench_configs.py --only cfg3 --reps 2 > /dev/null 2>&1
