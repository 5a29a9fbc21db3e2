#!/bin/bash
L=/root/repo/waveforms_b200/csrc/libwfm_s4.so
for mode in "" "WFM_K1_SPARSE4=1" "WFM_K1_UNIT=2" ; do
env WFM_LIB=$L $mode python bench.py --no-cpu --no-e2e --no-extras --steps 100 2>/dev/null | python -c "
import json,sys; r=json.load(sys.stdin); print('12w/1cta [$mode] cfg2', round(r['value'],1), round(r['roofline']['frac'],4), r['kernel_layout']['tile_samples'], r['kernel_layout']['samples_per_lane_unit'])"
env WFM_LIB=$L $mode python bench.py --no-cpu --no-e2e --no-extras --steps 100 --dtype f32 2>/dev/null | python -c "
import json,sys; r=json.load(sys.stdin); print('12w/1cta [$mode] cfg2 f32', round(r['value'],1), round(r['roofline']['frac'],4))"
done
env WFM_LIB=$L WFM_K1_SPARSE4=1 python tools/bench_configs.py --only cfg4,cfg5 2>/dev/null | python -c "
import json,sys; r=json.load(sys.stdin); print('s4 cfg4/5', r['cfg4']['GSa/s'], r['cfg5']['GSa/s'])"
env WFM_LIB=$L WFM_K1_SPARSE4=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden_fp64 or fp32" 2>&1 | tail -2
