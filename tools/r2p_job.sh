#!/bin/bash
python -m pytest tests/test_gpu_fft.py -m gpu -q 2>&1 | tail -2
for lib in libwfmb200 libwfm_nopad libwfm_pad4; do
WFM_LIB=/root/repo/waveforms_b200/csrc/$lib.so python tools/bench_dsp.py --reps 5 2>/dev/null | python -c "
import json,sys; r=json.load(sys.stdin)
print('$lib', [(k[:12], round(v['ms'],3)) for k,v in r['stages'].items() if k.startswith('K3')])"
done
