#!/bin/bash
o=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'sosfilt_scan' -c 2 -o $o/r2i_scan_new python tools/bench_dsp.py --reps 1 > /dev/null 2>&1
WFM_IIR_OLD_SCAN=1 ncu --set full --clock-control none --import-source on -k regex:'sosfilt_scan' -c 2 -o $o/r2i_scan_old python tools/bench_dsp.py --reps 1 > /dev/null 2>&1
ls -la $o/r2i*
