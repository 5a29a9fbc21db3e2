#!/bin/bash
# compute-sanitizer over a cross-section of the GPU tests (pairs, dense kernel, every unit size, joint IIR scan with
# cp.async tiles, prepared FFT responses, overlapping builder channels).  Writes gpurun_out/<tag>_sanitizer.txt.
tag=${1:-run}
o=gpurun_out
mkdir -p $o
SEL='tests/test_pairs.py tests/test_gpu_dsp.py tests/test_builder.py'
K='unit_sizes or pair or scan or reflection_response or prepared_response or overlapping or sweep_batch or cold or wide'
: > $o/${tag}_sanitizer.txt
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool" >> $o/${tag}_sanitizer.txt
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest $SEL tests/test_gpu_parity.py tests/test_gpu_fft.py -m gpu -q -x -k "$K" 2>&1 \
    | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|hazard|Error" | tail -12 >> $o/${tag}_sanitizer.txt
done
cat $o/${tag}_sanitizer.txt
