#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fft.py tests/test_gpu_dsp.py -q -m gpu -x 2>&1 | tail -4
python tools/bench_dsp.py --reps 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(' | '.join('%s %.3f' % (k[:14], v['ms']) for k, v in d['stages'].items()))"
python tools/bench_dsp.py --reps 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(' | '.join('%s %.3f' % (k[:14], v['ms']) for k, v in d['stages'].items()))"
