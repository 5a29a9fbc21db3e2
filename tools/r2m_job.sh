#!/bin/bash
python tools/bench_configs.py --only cfg3,cfg3v 2>/dev/null | python -c "
import json,sys; r=json.load(sys.stdin); print('dw12', r['cfg3']['GSa/s'], r['cfg3v']['GSa/s'], r['cfg3']['layout']['tile_samples'])"
for W in 16 14 10; do
WFM_LIB=/root/repo/waveforms_b200/csrc/libwfm_dw$W.so python tools/bench_configs.py --only cfg3,cfg3v 2>/dev/null | python -c "
import json,sys; r=json.load(sys.stdin); print('dw$W', r['cfg3']['GSa/s'], r['cfg3v']['GSa/s'], r['cfg3']['layout']['tile_samples'])"
done
python -m pytest tests/test_pairs.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -2
