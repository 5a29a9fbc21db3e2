#!/bin/bash
# A/B one environment knob of the library on one box:  tools/ab_env.sh VAR v1 v2 ...   (two rounds, bench.py K1 only)
var=$1; shift
for round in 1 2; do
  for v in "$@"; do
    env $var=$v timeout 300 python bench.py --no-cpu --no-e2e --steps 50 2>/dev/null | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$var=$v', round(d['value'],1), round(d['roofline']['frac'],4))"
  done
done
