tag=r1p; o=gpurun_out
python -m pytest tests -m gpu -q -x > $o/${tag}_pytest.txt 2>&1; tail -2 $o/${tag}_pytest.txt
python tools/bench_dsp.py --cpu > $o/${tag}_dsp_cfg4.json 2> $o/${tag}_dsp.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $o/${tag}_dsp_launches.csv python tools/bench_dsp.py --reps 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'fft_' -s 5 -c 3 -o $o/${tag}_dsp python tools/bench_dsp.py --reps 1 > /dev/null 2>&1
grep -A1 '"K' $o/${tag}_dsp_cfg4.json | grep -v "^--" | paste - - | cut -c1-110
