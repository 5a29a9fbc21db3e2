#!/bin/bash
o=gpurun_out; tag=r2g
ncu --set full --clock-control none --import-source on -k regex:'fft_' -s 3 -c 3 -o $o/${tag}_dsp python tools/bench_dsp.py --reps 1 > /dev/null 2>&1
ncu -i $o/${tag}_dsp.ncu-rep --page details > $o/${tag}_dsp_ncu_details.txt 2>/dev/null
ncu -i $o/${tag}_dsp.ncu-rep --page source --csv > $o/${tag}_dsp_src.csv 2>/dev/null
rm -f $o/${tag}_dsp.ncu-rep
ls -la $o | grep ${tag}_
