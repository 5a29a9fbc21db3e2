#!/usr/bin/env python
"""cfg4 (BASELINE.json configs[3]): flux-line predistortion at full size —
256 channels x 200 us @ 2 GSa/s (400 000 samples each): K1 sampling, K2 sample-time IIR
(exp-decay sos, exact and scan modes), K2b lfilter (predistort's IIR), K3 FFT correction
(correct_reflection, n = 400 000) and the kernel convolution of predistort (padded to a
7-smooth length).  CUDA events, warm, per stage; algorithmic bytes per SURVEY §8d
(IIR 16 B/sample, FFT filter 16 B/sample floor).  Prints one JSON object.

    python tools/bench_dsp.py [--channels 256] [--reps 5] [--cpu]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests' / 'golden'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--channels', type=int, default=256)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--cpu', action='store_true', help='time scipy on ONE channel for each stage')
    args = ap.parse_args()
    import torch
    import bench
    import cases
    from waveforms_b200 import distortion as D, dsp, engine
    from waveforms_b200.batch import channel_grid
    from waveforms_b200.lowering import lower

    ns = bench.b200_namespace()
    rate, t_end, n = 2e9, 200e-6, 400000
    rng = np.random.default_rng(20260004)
    chans = [cases.flux_channel(ns, rng, 20, t_end, rate)[0] for _ in range(args.channels)]
    batch = lower([channel_grid(w) for w in chans])
    prog = engine.Program(batch, 0)
    sig = prog.sample_device(dtype=engine.WFM_F64)
    stride = int(batch.waves['out_off'][1]) if args.channels > 1 else n
    sig2 = sig[:args.channels * stride].view(args.channels, stride)[:, :n]
    sos = D.exp_decay_filter([-0.03, 0.02], [0.1e-6, 0.3e-6], rate, inv=True, output='sos')
    ba = [D.exp_decay_filter(a, t, rate) for a, t in [(-0.03, 0.1e-6), (0.02, 0.3e-6)]]
    b, a = D.combine_filters(ba)
    freq = np.fft.fftfreq(n, 1 / rate)
    Hinv = 1 / D.reflection_filter(freq, 0.05, 13.3e-9)
    ker = D.zDistortKernel(1 / rate, [(0.1e-6, -0.03), (0.3e-6, 0.02)])
    Hk, L = D._centered_kernel_response(ker, n)
    pad = torch.zeros(args.channels, L, dtype=torch.float64, device='cuda')

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.reps + 1)]
        ev[0].record()
        for k in range(args.reps):
            fn()
            ev[k + 1].record()
        torch.cuda.synchronize()
        return min(ev[k].elapsed_time(ev[k + 1]) for k in range(args.reps))

    work = sig2.contiguous().clone()
    out = torch.empty_like(work)
    stages = {}

    def add(name, ms, samples, bytes_per_sample, note=''):
        stages[name] = {'ms': ms, 'GSa/s': samples / ms / 1e6, 'algorithmic_GB/s': samples * bytes_per_sample / ms / 1e6,
                        'bytes_per_sample': bytes_per_sample, 'note': note}

    tot = args.channels * n
    add('K1 sample (20 erf-edged squares / channel)', timed(lambda: prog.sample_device(dtype=engine.WFM_F64, out=sig)), tot, 8,
        'write-only')
    add('K2 sosfilt exact (%d section(s), 1 warp / signal)' % len(sos), timed(lambda: dsp.sosfilt_device(sos, work, out=out, mode='exact')),
        tot, 16, 'bit-identical to scipy.signal.sosfilt')
    add('K2 sosfilt scan (%d section(s), 1 CTA / signal)' % len(sos), timed(lambda: dsp.sosfilt_device(sos, work, out=out, mode='scan')), tot,
        16, 'block-parallel associative scan')
    w2 = work.clone()
    add('K2b lfilter exact (order %d)' % (max(len(a), len(b)) - 1), timed(lambda: dsp.lfilter_device(b, a, w2)), tot, 16,
        'bit-identical to scipy.signal.lfilter')
    w3 = work.clone()
    add('K2b lfilter scan (order %d)' % (max(len(a), len(b)) - 1), timed(lambda: dsp.lfilter_device(b, a, w3, mode='scan')), tot, 16,
        "block-parallel: predistort(iir_mode='scan')")
    add('K3 fft_filter n=400000 (correct_reflection)', timed(lambda: dsp.fft_filter_device(work, Hinv, out=out)), tot, 16,
        'four-step Stockham, H applied between the passes')
    pad[:, :n] = work
    add('K3 fft_filter n=%d (predistort kernel conv, K=%d)' % (L, len(ker)), timed(lambda: dsp.fft_filter_device(pad, Hk, out=pad)),
        args.channels * L, 16, 'linear convolution on a 7-smooth circular grid')
    res = {'workload': 'cfg4: %d flux channels x 400000 samples (200 us @ 2 GSa/s), fp64' % args.channels, 'stages': stages,
           'hbm_peak_GB/s': bench.measured_peak()[0]}
    if args.cpu:
        from scipy.signal import lfilter, sosfilt
        x = work[0].cpu().numpy()
        cpu = {}
        for name, fn in [('scipy.signal.sosfilt', lambda: sosfilt(sos, x)), ('scipy.signal.lfilter', lambda: lfilter(b, a, x)),
                         ('np.fft correct_reflection', lambda: np.fft.ifft(np.fft.fft(x) * Hinv).real)]:
            fn()
            t0 = time.perf_counter()
            for _ in range(3):
                fn()
            cpu[name] = {'ms_per_channel': (time.perf_counter() - t0) / 3 * 1e3}
        res['cpu_one_core_one_channel'] = cpu
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
