#!/usr/bin/env python
"""K1 on the other BASELINE.json configs at (per-GPU) full size, device-resident:
  cfg3  RB batch: 4096 channels x depth-1000 back-to-back DRAG cosPulses, I and Q (all samples active)
  cfg4  flux channels: 256 x 400 000 samples, 20 erf-edged squares each
  cfg5  sweep: 12 500 waveforms (one GPU's share of 100 000) x 20 000 samples of drag_sin / drag_sinx
A few channels are built through the drop-in API and lowered; the batch is `replicate`d to
full size with distinct amplitudes.  CUDA events, warm.  One JSON object on stdout.

    python tools/bench_configs.py [--reps 5] [--only cfg3,cfg5]
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests' / 'golden'))
import cases  # noqa: E402


def build(cfg, ns, pair=True):
    from waveforms_b200.batch import channel_grid
    from waveforms_b200.lowering import find_pairs, lower, replicate
    rng = np.random.default_rng(20260000 + int(cfg[3]))
    if cfg == 'cfg3':
        chans = []
        for ch in range(4):
            for which in (0, 1):
                chans.append(cases.rb_channel(ns, np.random.default_rng(20260003 + ch), 1000, ch, which=which)[0])
        items = [channel_grid(w) for w in chans]  # I, Q, I, Q ...: adjacent channels of one mixing() call
        base, copies = lower(find_pairs(items) if pair else items), 4096 // 4
    elif cfg == 'cfg4':
        chans = [cases.flux_channel(ns, rng, 20, 200e-6, 2e9)[0] for _ in range(8)]
        base, copies = lower([channel_grid(w) for w in chans]), 256 // 8
    else:
        chans = []
        for k in range(10):
            mk = ns.drag_sinx if k == 9 else ns.drag_sin
            kw = dict(block_freq=(-250e6, 180e6)) if k == 9 else dict(block_freq=(-250e6, ))
            w = rng.uniform(0.1, 1) * mk(rng.uniform(50e6, 150e6), 30e-9, plateau=0, delta=1e6, phase=rng.uniform(0, 6),
                                          t0=100e-9, **kw)
            w.start, w.stop, w.sample_rate = 0.0, 4e-6, 5e9
            chans.append(w)
        base, copies = lower([channel_grid(w) for w in chans]), 12500 // 10
    scale = 2.0 ** -(np.arange(copies) % 4)  # exact scalings: replica c == replica 0 * 2^-k bit for bit
    return chans, base, replicate(base, copies, amp_scale=scale), scale


def build_cfg3_vectorised(ns, n_ch, depth=1000, pair=True):
    """cfg3 with EVERY channel distinct (no replication), built from parameter arrays by
    waveforms_b200.builder: channel ch -> outputs 2*ch (I) and 2*ch+1 (Q), the same random
    Clifford-like sequence on both.  Returns (LoweredBatch, host seconds, spot-check closure)."""
    import time
    from waveforms_b200.builder import PulseTemplate, pulse_train_batch
    amps, phases = (0.5, 1.0), (0, np.pi / 2, np.pi, 3 * np.pi / 2)

    def fn(f8, a, p, which):
        if which is None:  # both outputs of the mixing() call: an I/Q pair template
            return lambda t0: ns.mixing(amps[a] * ns.cosPulse(20e-9) >> t0, freq=-20e6 * (1 + f8), phase=phases[p],
                                        DRAGScaling=4e-10)
        return lambda t0: ns.mixing(amps[a] * ns.cosPulse(20e-9) >> t0, freq=-20e6 * (1 + f8), phase=phases[p],
                                    DRAGScaling=4e-10)[which]
    t_begin = time.perf_counter()
    rng = np.random.default_rng(20260003)
    gate = rng.integers(0, 8, (n_ch, depth))                       # (amp, phase) of every pulse
    per_ch = (np.arange(n_ch) % 8)[:, None] * 8 + gate            # + the channel's carrier
    stop = 100e-9 + 20e-9 * depth + 900e-9
    fns1 = [fn(f8, a, p, which) for which in (0, 1) for f8 in range(8) for a in range(2) for p in range(4)]
    if pair:
        fns = [fn(f8, a, p, None) for f8 in range(8) for a in range(2) for p in range(4)]
        idx = per_ch
        t0 = np.tile(100e-9 + 20e-9 * np.arange(depth) + 10e-9, (n_ch, 1))
    else:
        fns = fns1
        idx = np.stack([per_ch, 64 + per_ch], axis=1).reshape(2 * n_ch, depth)
        t0 = np.tile(100e-9 + 20e-9 * np.arange(depth) + 10e-9, (2 * n_ch, 1))
    templates = [PulseTemplate.trace(f) for f in fns]
    batch = pulse_train_batch(templates, idx, t0, 0, stop, 2e9)
    host_s = time.perf_counter() - t_begin

    def object_channel(row):
        """output row `row` (2 * channel + which) through the object API"""
        ch, which = divmod(row, 2)
        w = ns.WaveVStack([fns1[int(i) + 64 * which](float(t)) for i, t in zip(per_ch[ch], t0[0])])
        w.start, w.stop, w.sample_rate = 0, stop, 2e9
        return w
    return batch, host_s, object_channel


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--only', default='cfg3,cfg4,cfg5')
    ap.add_argument('--no-pair', dest='pair', action='store_false', help='cfg3: I and Q as separate channels')
    ap.add_argument('--cfg3v-channels', type=int, default=512,
                    help="channels of the builder-made cfg3 batch ('cfg3v' in --only); 512 = one GPU's share of 4096 on 8 GPUs")
    args = ap.parse_args()
    import torch
    import bench
    from waveforms_b200 import engine
    ns = bench.b200_namespace()
    from waveforms_b200 import multy_drag
    ns.drag_sin, ns.drag_sinx = multy_drag.drag_sin, multy_drag.drag_sinx
    peak = bench.measured_peak()[0]
    res = {}
    for cfg in args.only.split(','):
        if cfg == 'cfg3v':
            import time
            batch, host_s, object_channel = build_cfg3_vectorised(ns, args.cfg3v_channels, pair=args.pair)
            t0 = time.perf_counter()
            prog = engine.Program(batch, 0)
            out = prog.sample_device(dtype=engine.WFM_F64)
            torch.cuda.synchronize()
            create_s = time.perf_counter() - t0
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.reps + 1)]
            ev[0].record()
            for k in range(args.reps):
                prog.sample_device(dtype=engine.WFM_F64, out=out)
                ev[k + 1].record()
            torch.cuda.synchronize()
            ms = min(ev[k].elapsed_time(ev[k + 1]) for k in range(args.reps))
            n = int(batch.chan_n.sum())
            worst = 0.0
            for row in (1, batch.n_channels - 2):  # one Q and one I output against the object API (sampled alone)
                off, cnt = int(batch.chan_off[row]), int(batch.chan_n[row])
                want = object_channel(row).sample()
                worst = max(worst, float(np.max(np.abs(out[off:off + cnt].cpu().numpy() - want)) / np.max(np.abs(want))))
            ok = worst <= 4e-15
            res[cfg] = {'channels': batch.n_channels, 'pulses': int(batch.n_channels) * 1000, 'samples': n,
                        'iq_pairs': bool(args.pair), 'max_rel_diff_vs_object_api': worst,
                        'host_build_s': host_s, 'ir_GB': batch.nbytes() / 1e9, 'create_and_first_sample_s': create_s,
                        'ms': ms, 'GSa/s': n / ms / 1e6, 'roofline_frac': n * 8 / ms / 1e6 / peak,
                        'equals_object_api_to_4e-15': ok, 'layout': prog.info()}
            prog.close()
            del out
            torch.cuda.empty_cache()
            continue
        chans, base, batch, scale = build(cfg, ns, pair=args.pair)
        prog = engine.Program(batch, 0)
        out = torch.empty(batch.total_samples, dtype=torch.float64, device='cuda')
        prog.sample_device(dtype=engine.WFM_F64, out=out)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.reps + 1)]
        ev[0].record()
        for k in range(args.reps):
            prog.sample_device(dtype=engine.WFM_F64, out=out)
            ev[k + 1].record()
        torch.cuda.synchronize()
        ms = min(ev[k].elapsed_time(ev[k + 1]) for k in range(args.reps))
        n = int(batch.chan_n.sum())
        # size-independent check: every replica equals replica 0 times its (power-of-two) amplitude scale
        per = base.total_samples
        v = out.view(len(scale), per)
        ok = bool(torch.equal(v, v[0][None, :] * torch.from_numpy(scale).cuda()[:, None]))
        res[cfg] = {'channels': batch.n_channels, 'waves': len(batch.waves), 'samples': n, 'ms': ms, 'GSa/s': n / ms / 1e6, 'GB/s': n * 8 / ms / 1e6,
                    'roofline_frac': n * 8 / ms / 1e6 / peak, 'replicas_bit_exact': ok, 'layout': prog.info()}
        prog.close()
        del out
        torch.cuda.empty_cache()
    # configs 1 and 2 are tiny (0.16 / 64 MB of output): their figure of merit is latency
    import time
    lat = {}
    for name in ('cfg1', 'cfg2x1'):
        if name == 'cfg1':
            x_wav, y_wav = cases._readme(ns)
            chans = [x_wav, y_wav]
            for w in chans:
                w.start, w.stop, w.sample_rate = -1e-6, 9e-6, 1e9
        else:
            chans = bench.build_frame(ns)
        from waveforms_b200.batch import channel_grid
        from waveforms_b200.lowering import lower
        batch = lower([channel_grid(w) for w in chans]).pin()
        prog = engine.Program(batch, 0)
        out = torch.empty(batch.total_samples, dtype=torch.float64, device='cuda')
        for _ in range(3):
            prog.sample_device(dtype=engine.WFM_F64, out=out)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
        ev[0].record()
        for k in range(20):
            prog.sample_device(dtype=engine.WFM_F64, out=out)
            ev[k + 1].record()
        torch.cuda.synchronize()
        k1_us = min(ev[k].elapsed_time(ev[k + 1]) for k in range(20)) * 1e3
        t0 = time.perf_counter()
        for _ in range(20):
            prog.sample_device(dtype=engine.WFM_F64, out=out)
        torch.cuda.synchronize()
        launch_us = (time.perf_counter() - t0) / 20 * 1e6
        prog.close()
        host = torch.empty(batch.total_samples, dtype=torch.float64, pin_memory=True).numpy()
        ts = []
        for _ in range(6):
            t0 = time.perf_counter()
            p2 = engine.Program(batch, 0)
            p2.sample_host(out=host)
            p2.close()
            ts.append((time.perf_counter() - t0) * 1e6)
        n = int(batch.chan_n.sum())
        lat[name] = {'channels': len(chans), 'samples': n, 'k1_device_us': k1_us, 'k1_back_to_back_wall_us': launch_us,
                     'create_sample_host_destroy_us': min(ts[1:]), 'GSa/s_device': n / k1_us / 1e3}
    res['latency'] = lat
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
