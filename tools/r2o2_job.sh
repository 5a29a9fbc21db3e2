o=gpurun_out; tag=r2o
ncu --set full --clock-control none --import-source on -k regex:'fft_cols' -s 2 -c 1 -o $o/${tag}_cols python tools/bench_dsp.py --reps 1 > /dev/null 2>&1
ncu -i $o/${tag}_cols.ncu-rep --page details > $o/${tag}_cols_details.txt 2>/dev/null
ncu -i $o/${tag}_cols.ncu-rep --page source --csv > $o/${tag}_cols_src.csv 2>/dev/null
rm -f $o/${tag}_cols.ncu-rep
ls -la $o | grep ${tag}_
