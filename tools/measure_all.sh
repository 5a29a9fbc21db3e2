#!/bin/bash
# One B200 box: the whole evidence set of a state of the repo, written to gpurun_out/<tag>_*.
#   gpurun --timeout 1500 -- 'tools/measure_all.sh r1o'
tag=${1:-run}
o=gpurun_out
mkdir -p $o
python -m pytest tests -m gpu -q -x > $o/${tag}_pytest.txt 2>&1; tail -3 $o/${tag}_pytest.txt
python bench.py > $o/${tag}_bench_cfg2.json 2> $o/${tag}_bench.err; tail -c 1500 $o/${tag}_bench_cfg2.json
python bench.py --impl reference --steps 5 --warmup 3 > $o/${tag}_bench_cfg2_reference_arm.json 2>> $o/${tag}_bench.err
python bench.py --dtype f32 --no-cpu --no-e2e > $o/${tag}_bench_cfg2_f32.json 2>> $o/${tag}_bench.err
python tools/bench_configs.py --only cfg3,cfg3v,cfg4,cfg5 > $o/${tag}_k1_other_configs.json 2> $o/${tag}_configs.err
# config 3 at FULL size on one GPU, every channel distinct (vectorised builder): 8192 outputs, 8.2 M pulses
python tools/bench_configs.py --only cfg3v --cfg3v-channels 4096 > $o/${tag}_cfg3_full_builder.json 2>> $o/${tag}_configs.err
python tools/bench_dsp.py --cpu > $o/${tag}_dsp_cfg4.json 2> $o/${tag}_dsp.err
# launch lists (cold-cache, serialised: shares of the step, not absolute times)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches_cfg2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $o/${tag}_dsp_launches.csv \
    python tools/bench_dsp.py --reps 1 > /dev/null 2>&1
# full captures: K1 (one launch of the bench kernel), K3 (the three FFT passes), K2 scan
ncu --set full --clock-control none --import-source on -k regex:sample_kernel -s 3 -c 1 -o $o/${tag}_k1 \
    python bench.py --steps 2 --warmup 3 --frames 16 --no-cpu --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'fft_|sosfilt_scan' -s 4 -c 4 -o $o/${tag}_dsp \
    python tools/bench_dsp.py --reps 1 > /dev/null 2>&1
ls -la $o | grep ${tag}_
