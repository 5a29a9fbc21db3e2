#!/bin/bash
# compute-sanitizer over the K3 tests (TMA row pass, compile-time plans, parked twiddles), the staged create of small
# programs and the device template expansion.  Writes gpurun_out/<tag>_sanitizer_fft.txt.
tag=${1:-run}
o=gpurun_out
mkdir -p $o
: > $o/${tag}_sanitizer_fft.txt
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool" >> $o/${tag}_sanitizer_fft.txt
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_fft.py tests/test_gpu_dsp.py tests/test_builder.py tests/test_gpu_parity.py -m gpu -q -x \
      -k "fft or reflection or predistort or kernel or compact or fast_create or four_step or two_level or padded or lfilter" 2>&1 \
    | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|hazard|Error" | tail -12 >> $o/${tag}_sanitizer_fft.txt
done
cat $o/${tag}_sanitizer_fft.txt
