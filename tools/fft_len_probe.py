#!/usr/bin/env python
"""K3 fft_filter time against the transform length (256 signals): which padded 7-smooth length is cheapest?"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from waveforms_b200 import dsp
def timed(f, reps=5):
    f(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
rng = np.random.default_rng(0)
for n in [int(v) for v in sys.argv[1:]] or [400000, 403200, 405000, 409600, 406250, 403368, 404250, 408240, 409500, 410000, 412160]:
    x = torch.from_numpy(rng.standard_normal((256, n))).cuda()
    H = np.fft.fft(rng.standard_normal(n) * np.exp(-np.arange(n) / 50.0))
    out = torch.empty_like(x)
    ms = timed(lambda: dsp.fft_filter_device(x, H, out=out))
    ref = np.fft.ifft(np.fft.fft(x[3].cpu().numpy()) * H).real
    err = np.max(np.abs(out[3].cpu().numpy() - ref)) / np.max(np.abs(ref))
    print(n, '%.3f ms' % ms, '%.3f ns/sample' % (ms * 1e6 / (256 * n)), 'err %.1e' % err, flush=True)
