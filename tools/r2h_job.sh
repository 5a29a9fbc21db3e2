#!/bin/bash
for g in 0 2 4 7 8 11 16 22; do
  WFM_FFT_GROUP=$g timeout 300 python tools/bench_dsp.py --reps 5 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('group $g', ' | '.join('%s %.3f' % (k[:14], v['ms']) for k, v in d['stages'].items() if k.startswith(('K3'))))"
done
python -m pytest tests/test_gpu_fft.py tests/test_gpu_dsp.py -q -m gpu -x 2>&1 | tail -2
