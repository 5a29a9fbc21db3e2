ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'fft_' -s 3 -c 3 --csv --log-file gpurun_out/r2m_rows.csv python tools/bench_dsp.py --reps 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2m_rows.csv')))
i=[k for k,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[i]
for r in rows[i+1:]:
    print(r[h.index('Kernel Name')][:40], r[h.index('Metric Name')], r[h.index('Metric Value')])
PY
python -m pytest tests/test_gpu_fft.py tests/test_gpu_dsp.py -q -m gpu -x 2>&1 | tail -2
python tools/bench_dsp.py --reps 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(' | '.join('%s %.3f' % (k[:14], v['ms']) for k, v in d['stages'].items() if k.startswith('K3')))"
