#!/bin/bash
# round-2 first look: calibrations + ncu of K1 on dense cfg3
o=gpurun_out
python - > $o/r2a_calib.json 2> $o/r2a_calib.err <<'PY'
import json, torch
from waveforms_b200 import engine
torch.cuda.init()
r = {'fp64': engine.calibrate_fp64(5), 'd2h_1GiB': engine.calibrate_copy(1 << 30, 'd2h', 4), 'h2d_1GiB': engine.calibrate_copy(1 << 30, 'h2d', 4),
     'd2h_4GB': engine.calibrate_copy(4096000000, 'd2h', 3)}
print(json.dumps(r, indent=1))
PY
cat $o/r2a_calib.json
ncu --set full --clock-control none --import-source on -k regex:sample_kernel -s 2 -c 1 -o $o/r2a_k1_cfg3 \
    python tools/bench_configs.py --only cfg3 --reps 2 > $o/r2a_cfg3_under_ncu.json 2> $o/r2a_cfg3_ncu.err
ls -la $o | grep r2a_
