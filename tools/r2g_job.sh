#!/bin/bash
o=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4
( time python bench.py > $o/r2g_bench.json 2> $o/r2g_bench.err ) 2>&1 | grep real
tail -5 $o/r2g_bench.err
python - <<'PY'
import json
r=json.load(open('gpurun_out/r2g_bench.json'))
print('value',r['value'],'frac',r['roofline']['frac'],'e2e',r['e2e']['value'],r['e2e']['frac_of_copy_ceiling'], r['e2e'].get('fp32'))
print('parity',r['parity']); print('calib',r['calibration']); print('fp32',r.get('fp32')); print('extras_s', r.get('extras_s'), r.get('extras_error'))
for k,v in (r.get('configs') or {}).items():
    print(k, round(v['GSa/s'],1), v['roofline']['bound'], round(v['roofline']['frac'],3), v['roofline'].get('fp64_pipe',{}).get('frac'), v.get('parity'), (v.get('cpu_baseline') or {}).get('value'))
p=r.get('pipeline_cfg4')
if p:
    print('pipeline', p['pipeline_ms'], p['GSa/s'], p['roofline']['frac'])
    for k,v in p['stages'].items(): print('  ',k, round(v['ms'],3), round(v['roofline']['frac'],3))
    print(p['parity']); print(p['iir_error']); print(p['cpu_baseline'])
PY
