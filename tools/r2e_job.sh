#!/bin/bash
# compact IR: GPU tests + cfg3 bench line + sanitizer on the expansion
set -x
mkdir -p gpurun_out
python -m pytest tests/test_builder.py tests/test_abi.py -q -m gpu -x > gpurun_out/r2e_pytest_builder.txt 2>&1; tail -5 gpurun_out/r2e_pytest_builder.txt
python bench.py --steps 6 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -c 600 gpurun_out/r2e_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2e_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'])
print(json.dumps(d['configs']['cfg3'].get('compact'), indent=1))
print({k: d['configs'][k]['GSa/s'] for k in d['configs']})
PY
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_builder.py -q -m gpu -k "compact" > gpurun_out/r2e_sanitizer.txt 2>&1; tail -4 gpurun_out/r2e_sanitizer.txt
python -m pytest tests -q -m gpu -x > gpurun_out/r2e_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r2e_pytest_gpu.txt
