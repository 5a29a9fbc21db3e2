"""Randomised parity at scale: >= 200 seeded waveform programs sampled by
the CUDA path and by the reference's own compiled evaluator (oracle/_ref; the oracle port
for the Python-level multi-DRAG ids) on identical tuples — north_star: "within 1e-12
relative (fp64) or 1e-6 (fp32) on identical random waveform programs".

The generator covers: all 17 basis functions; GHz carriers over 100 us (arguments of 1e5..1e6
rad); products, integer / negative / fractional exponents; `mixing` with DRAGScaling and with
block_freq; clip (`cut`); shifts; stacks with `offset` and `shift`; the three grid modes
(`sample()` = arange, `__call__(x)` = explicit abscissae incl. points exactly ON segment
bounds, `sample(chunk_size=...)` = linspace(endpoint=False) per chunk); fp64 and fp32."""
import warnings

import numpy as np
import pytest

from helpers import FP32_TOL, FP64_TOL, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def pair_small_batches(monkeypatch):
    """'auto' pairing starts at batch.PAIR_MIN_SAMPLES samples (small launches are latency-bound); the batches
    here are small and pairs are (part of) what is tested."""
    from waveforms_b200 import batch
    monkeypatch.setattr(batch, 'PAIR_MIN_SAMPLES', 0)

N_PROGRAMS = 240


@pytest.fixture(scope='module')
def X():
    from tools import bench_extras
    return bench_extras


def _envelope(ns, rng, T):
    """One pulse of duration scale T (seconds or arbitrary units), non-zero on a bounded support.
    Returns (pulse, differentiable): the reference has no derivative for SINC (its table entry raises
    IndexError) nor for the DRAG family, so those never go through the DRAG branch of ``mixing``."""
    from waveforms_b200 import multy_drag
    kind = int(rng.integers(15))
    w = rng.uniform(0.2, 1.0) * T
    if kind == 0:
        return (ns.cosPulse(w)), True
    if kind == 1:
        return (ns.cosPulse(w, plateau=rng.uniform(0.1, 0.5) * T)), True
    if kind == 2:
        return (ns.gaussian(w)), True
    if kind == 3:
        return (ns.gaussian(w, plateau=rng.uniform(0.1, 0.4) * T)), True
    if kind == 4:
        return ns.square(w, edge=rng.uniform(0.02, 0.2) * w), True  # erf edges
    if kind == 5:
        return (ns.square(w, edge=rng.uniform(0.05, 0.3) * w, type=('cos', 'linear')[int(rng.integers(2))])), True
    if kind == 6:
        return (ns.coshPulse(w, eps=rng.uniform(0.5, 3.0))), True
    if kind == 7:
        return (ns.mollifier(w) if rng.random() < 0.5 else ns.mollifier(w, plateau=0.2 * T, d=int(rng.integers(1, 3))) * (0.1 * T)**2), True
    if kind == 8:
        return (ns.sinc(rng.uniform(2, 12) / T) * ns.square(w)), False
    if kind == 9:
        return (ns.exp(rng.uniform(-3, 3) / T) * ns.square(w, edge=0.1 * w)), True
    if kind == 10:
        return (ns.gaussian(w, d=int(rng.integers(1, 4))) * (0.3 * T)**2 * rng.uniform(0.01, 0.1)), True
    if kind == 11:
        f0 = rng.uniform(1, 6) / T
        ch = ('linear', 'exponential', 'hyperbolic')[int(rng.integers(3))]
        return (ns.chirp(f0, rng.uniform(1.2, 2.0) * f0, w, rng.uniform(0, 3), ch)), True
    if kind == 12:
        pts = rng.standard_normal(int(rng.integers(2, 30)))
        return (ns.samplingPoints(-0.4 * w, 0.4 * w, pts) * ns.square(0.8 * w)), True
    if kind == 13:
        return (ns.poly([rng.uniform(-1, 1), rng.uniform(-1, 1) / T, rng.uniform(-1, 1) / T**2]) * ns.square(w)), True
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        f = rng.uniform(3, 12) / T
        r = rng.random()
        if r < 0.25:
            return multy_drag.drag_sinx(f, w, plateau=rng.uniform(0, 0.3) * T if rng.random() < 0.5 else 0, delta=rng.uniform(-0.1, 0.1) / T,
                                        block_freq=tuple(rng.uniform(8, 30, 2) * np.array([-1, 1]) / T), phase=rng.uniform(0, 6),
                                        t0=-0.5 * w), False
        if r < 0.6:
            return multy_drag.drag_sin(f, w, plateau=rng.uniform(0, 0.3) * T if rng.random() < 0.5 else 0, delta=rng.uniform(-0.1, 0.1) / T,
                                       block_freq=tuple(rng.uniform(-30, 30, int(rng.integers(1, 3))) / T), phase=rng.uniform(0, 6),
                                       t0=-0.5 * w), False
        return (ns.drag(f, w, plateau=0.1 * T, delta=0.05 / T, block_freq=(None, -8 / T)[int(rng.integers(2))], phase=rng.uniform(0, 6), t0=-0.5 * w)), False


def _program(ns, seed):
    """(waveform-or-stack, grid mode).  Half of the programs live on a nanosecond scale with GHz carriers
    sampled over 100 us (huge trigonometric arguments), half on a unit scale."""
    rng = np.random.default_rng(90000 + seed)
    ns_scale = seed % 2 == 0
    T = rng.uniform(20e-9, 400e-9) if ns_scale else rng.uniform(0.5, 2.0)
    span = 100e-6 if ns_scale else 40.0
    pulses = []
    for _ in range(int(rng.integers(2, 9))):
        p, differentiable = _envelope(ns, rng, T)
        t0 = rng.uniform(0.05, 0.95) * span
        p = rng.uniform(-1, 1) * p >> t0
        r = rng.random()
        if r < 0.45 and not differentiable:
            r = 0.5  # carrier by plain multiplication instead
        if r < 0.45:
            freq = rng.uniform(-4e9, 4e9) if ns_scale else rng.uniform(-30, 30)
            kw = dict(DRAGScaling=rng.uniform(0.001, 0.05) * T) if rng.random() < 0.6 else dict(block_freq=freq * rng.uniform(1.5, 3))
            p = ns.mixing(p, freq=freq, phase=rng.uniform(0, 6.28), **kw)[int(rng.integers(2))]
        elif r < 0.55:
            p = p * ns.cos(2 * np.pi * (rng.uniform(0.1e9, 5e9) if ns_scale else rng.uniform(1, 20)), rng.uniform(0, 6))
        elif r < 0.62:
            p = (0.3 * p + 0.0) * p                       # products: squares of envelopes
        elif r < 0.68:
            g = ns.gaussian(T) >> t0
            p = g**(0.5, 1.5, -0.5, 3)[int(rng.integers(4))] * rng.uniform(0.1, 1)   # single-term expression: any exponent
        pulses.append(p)
    mode = ('sample', 'call', 'chunks')[seed % 3]
    as_stack = rng.random() < 0.4
    if as_stack:
        w = ns.WaveVStack(pulses)
        if rng.random() < 0.6:
            w = w + rng.uniform(-0.5, 0.5)
        if rng.random() < 0.5:
            w = w >> rng.uniform(-0.02, 0.02) * span
        if mode == 'chunks':
            mode = 'sample'
    else:
        w = ns.zero()
        for p in pulses:
            w = w + p
        if rng.random() < 0.25:
            w = ns.cut(w, min=-rng.uniform(0.1, 0.6), max=rng.uniform(0.1, 0.6))
    n = int(rng.integers(1500, 9000))
    w.start, w.stop, w.sample_rate = 0.0, span, n / span
    return w, mode, rng


def _reference(X, w, x=None):
    from oracle import wfm_oracle as O
    if x is None:
        return X.cpu_sample(w)
    calc, _ = X.ref_calc()
    kw = {} if calc is None else {'calc': calc}
    with warnings.catch_warnings(), np.errstate(all='ignore'):
        warnings.simplefilter('ignore')
        if hasattr(w, 'wlist'):
            return O.stack_call(list(w.wlist), x, w.offset, w.shift, **kw)
        return O.waveform_call(w.bounds, w.seq, x, w.min, w.max, **kw)


@pytest.mark.parametrize('block', range(8))
def test_random_programs_against_the_reference_evaluator(ns, X, block):
    from waveforms_b200 import sample_batch
    per = N_PROGRAMS // 8
    checked = {'sample': 0, 'call': 0, 'chunks': 0, 'fp32': 0}
    sample_ws = []
    for seed in range(block * per, (block + 1) * per):
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            w, mode, rng = _program(ns, seed)
            if mode == 'sample':
                got, want = w.sample(), _reference(X, w)
                sample_ws.append(w)
            elif mode == 'call':
                # explicit abscissae: an irregular sorted grid that contains segment bounds EXACTLY
                bounds = np.array([b for b in (w.bounds if not hasattr(w, 'wlist') else w.wlist[0][0]) if np.isfinite(b)])
                x = np.sort(np.concatenate([rng.uniform(0, w.stop, 3000), bounds[(bounds > 0) & (bounds < w.stop)],
                                            np.linspace(0, w.stop, 501)]))
                got, want = w(x), _reference(X, w, x)
            else:
                chunk = int(rng.integers(500, 3000))
                got = np.concatenate(list(w.sample(chunk_size=chunk)))
                # the reference's chunked grid: linspace(start, stop_k, size_k, endpoint=False) per chunk (waveform.py:225-246)
                xs, start = [], w.start
                length = chunk / w.sample_rate
                while start < w.stop:
                    if start + length > w.stop:
                        stop, size = w.stop, round((w.stop - start) * w.sample_rate)
                    else:
                        stop, size = start + length, chunk
                    xs.append(np.linspace(start, stop, size, endpoint=False))
                    start = stop
                want = np.concatenate([_reference(X, w, x) for x in xs])
        assert got.shape == want.shape, (seed, mode)
        assert np.all(np.isfinite(want)), seed
        assert rel_err(got, want) <= FP64_TOL, (seed, mode, rel_err(got, want))
        checked[mode] += 1
    # fp32 output (and the batched path, I/Q pairing on) for the sample()-grid programs of the block
    if sample_ws:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            f32 = sample_batch(sample_ws, dtype=np.float32).numpy()
            f64 = sample_batch(sample_ws).numpy()
        for w, a, b in zip(sample_ws, f32, f64):
            want = _reference(X, w)
            assert rel_err(b, want) <= FP64_TOL
            assert rel_err(a.astype(np.float64), want) <= FP32_TOL
            checked['fp32'] += 1
    assert checked['sample'] + checked['call'] + checked['chunks'] == per
