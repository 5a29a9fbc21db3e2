#!/bin/bash
o=gpurun_out
python bench.py --no-cpu --no-e2e --steps 50 > $o/r2e_bench.json 2> $o/r2e_bench.err; python -c "
import json; r=json.load(open('$o/r2e_bench.json')); print(r['value'], r['roofline']['frac'])"
ncu --set full --clock-control none --import-source on -k regex:sample_kernel -s 3 -c 1 -o $o/r2e_k1 \
    python bench.py --steps 2 --warmup 3 --frames 16 --no-cpu --no-e2e > /dev/null 2>&1
ls -la $o/r2e*
