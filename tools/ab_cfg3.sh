#!/bin/bash
# A/B of K1 variants on the dense RB batch (cfg3) and the other configs: tools/ab_cfg3.sh <lib> [<lib> ...]
for round in 1 2; do
  for lib in "$@"; do
    WFM_LIB=$lib timeout 600 python tools/bench_configs.py --only cfg3,cfg3v,cfg5 --reps 5 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', {k: round(v['GSa/s'],1) for k, v in d.items() if isinstance(v, dict) and 'GSa/s' in v})"
  done
done
