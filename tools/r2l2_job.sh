python -m pytest tests/test_gpu_fft.py tests/test_gpu_dsp.py -q -m gpu -x 2>&1 | tail -2
for r in 1 2; do
for v in 0 1; do
  if [ $v = 1 ]; then export WFM_FFT_NO_TMA=1; else unset WFM_FFT_NO_TMA; fi
  python tools/bench_dsp.py --reps 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('no_tma=$v', ' | '.join('%s %.3f' % (k[:14], v['ms']) for k, v in d['stages'].items() if k.startswith('K3')))"
done; done
