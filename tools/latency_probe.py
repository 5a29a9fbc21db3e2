#!/usr/bin/env python
"""cfg1 latency by API call: wfm_program_create / wfm_sample_host / wfm_program_destroy of the README pair."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests' / 'golden'))
import torch, bench, cases
from waveforms_b200 import engine
from waveforms_b200.batch import channel_grid
from waveforms_b200.lowering import lower
ns = bench.b200_namespace()
x, y = cases._readme(ns)
for w in (x, y):
    w.start, w.stop, w.sample_rate = -1e-6, 9e-6, 1e9
batch = lower([channel_grid(x), channel_grid(y)]).pin()
host = torch.empty(batch.total_samples, dtype=torch.float64, pin_memory=True).numpy()
dev = torch.cuda.current_device()
T = {'create': [], 'sample_host': [], 'destroy': []}
for i in range(300):
    t0 = time.perf_counter(); p = engine.Program(batch, dev)
    t1 = time.perf_counter(); p.sample_host(out=host)
    t2 = time.perf_counter(); p.close()
    t3 = time.perf_counter()
    if i >= 50:
        T['create'].append(t1 - t0); T['sample_host'].append(t2 - t1); T['destroy'].append(t3 - t2)
print({k: round(float(np.median(v)) * 1e6, 1) for k, v in T.items()}, 'us (median)', {k: round(float(np.min(v)) * 1e6, 1) for k, v in T.items()}, 'min')
