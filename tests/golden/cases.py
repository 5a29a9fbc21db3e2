"""Case definitions shared by the golden-vector generator (run against the
unmodified reference in the build container) and by the parity tests (run
against waveforms_b200 + the oracle).

Every case is a function ``build(ns)`` taking a namespace that exposes the
reference's public names (``cosPulse``, ``mixing``, ``WaveVStack`` ...) and
returning ``(obj, grid)``:

    obj   Waveform or WaveVStack built with ``ns``
    grid  ('explicit', ndarray) -> obj(x)
          ('sample',)            -> obj.sample() (start/stop/sample_rate set)

Configs follow SURVEY.md §8(d) at reduced size (the oracle must finish in
seconds and fixtures stay small).
"""
from __future__ import annotations

import numpy as np

CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def _sampled(w, start, stop, rate):
    w.start, w.stop, w.sample_rate = start, stop, rate
    return w, ('sample', )


# ---- config 1: README example (reference README.md:28-52) ---------------------
def _readme(ns):
    pulse = ns.cosPulse(20e-9)
    x_wav, y_wav = ns.zero(), ns.zero()
    I, Q = ns.mixing(0.5 * pulse, freq=-20e6, DRAGScaling=0.2)
    x_wav += I
    y_wav += Q
    I, Q = ns.mixing(pulse >> 1e-6, freq=-20e6, phase=np.pi / 2,
                     DRAGScaling=0.2)
    x_wav += I
    y_wav += Q
    I, Q = ns.mixing((0.5 * pulse) >> 2e-6, freq=-20e6, DRAGScaling=0.2)
    x_wav += I
    y_wav += Q
    return x_wav, y_wav


@case
def readme_x_call(ns):
    return _readme(ns)[0], ('explicit', np.linspace(-1e-6, 9e-6, 10001))


@case
def readme_y_call(ns):
    return _readme(ns)[1], ('explicit', np.linspace(-1e-6, 9e-6, 10001))


@case
def readme_x_sample(ns):
    return _sampled(_readme(ns)[0], -1e-6, 9e-6, 1e9)


@case
def readme_y_sample(ns):
    return _sampled(_readme(ns)[1], -1e-6, 9e-6, 1e9)


# ---- config 2 units: XY channel (DRAG cosPulse/gaussian) and Z channel -----------
def xy_channel(ns, rng, n_pulse, pitch, t_end, rate, stack=True):
    pulses = []
    freq = rng.uniform(-200e6, 200e6)
    for k in range(n_pulse):
        env = ns.cosPulse(20e-9) if k % 2 == 0 else ns.gaussian(20e-9)
        t0 = round((50e-9 + k * pitch + rng.uniform(0, 100e-9)) * rate) / rate
        amp = rng.uniform(0.1, 1)
        I, Q = ns.mixing(amp * env >> t0, freq=freq,
                         phase=rng.uniform(0, 2 * np.pi),
                         DRAGScaling=rng.uniform(2e-10, 1e-9))
        pulses.append(I)
    if stack:
        w = ns.WaveVStack(pulses)
    else:
        w = ns.zero()
        for p in pulses:
            w = w + p
    return _sampled(w, 0, t_end, rate)


def z_channel(ns, rng, n_pulse, t_end, rate, edge=2e-9, wmin=20e-9,
              wmax=200e-9):
    w = ns.zero()
    slot = t_end / n_pulse
    for k in range(n_pulse):
        width = rng.uniform(wmin, wmax)
        centre = round((k + 0.5) * slot * rate) / rate
        w = w + rng.uniform(-0.5, 0.5) * (ns.square(width, edge=edge) >> centre)
    return _sampled(w, 0, t_end, rate)


@case
def cfg2_xy_stack(ns):
    return xy_channel(ns, np.random.default_rng(20260002), 25, 400e-9, 10e-6,
                      2e9)


@case
def cfg2_xy_merged(ns):
    return xy_channel(ns, np.random.default_rng(20260012), 12, 400e-9, 5e-6,
                      2e9, stack=False)


@case
def cfg2_z(ns):
    return z_channel(ns, np.random.default_rng(20260022), 10, 10e-6, 2e9)


# ---- config 3 unit: RB channel, back-to-back DRAG cosPulses, I and Q ------------
def rb_channel(ns, rng, depth, ch, rate=2e9, which=0):
    pulses = []
    freq = -20e6 * (1 + ch % 8)
    for k in range(depth):
        amp = (0.5, 1.0)[int(rng.integers(2))]
        phase = (0, np.pi / 2, np.pi, 3 * np.pi / 2)[int(rng.integers(4))]
        t0 = 100e-9 + 20e-9 * k + 10e-9
        IQ = ns.mixing(amp * ns.cosPulse(20e-9) >> t0, freq=freq, phase=phase,
                       DRAGScaling=4e-10)
        pulses.append(IQ[which])
    w = ns.WaveVStack(pulses)
    return _sampled(w, 0, 100e-9 + 20e-9 * depth + 900e-9, rate)


@case
def cfg3_rb_I(ns):
    return rb_channel(ns, np.random.default_rng(20260003), 60, 3, which=0)


@case
def cfg3_rb_Q(ns):
    return rb_channel(ns, np.random.default_rng(20260003), 60, 3, which=1)


# ---- config 4 unit: flux pulse train (erf edges, overlaps) ------------------------
def flux_channel(ns, rng, n_pulse, t_end, rate):
    w = ns.zero()
    for _ in range(n_pulse):
        width = rng.uniform(0.1e-6, 0.05 * t_end)
        centre = round(rng.uniform(0.05 * t_end, 0.95 * t_end) * rate) / rate
        w = w + rng.uniform(-0.5, 0.5) * (ns.square(width, edge=5e-9) >>
                                          centre)
    return _sampled(w, 0, t_end, rate)


@case
def cfg4_flux(ns):
    return flux_channel(ns, np.random.default_rng(20260004), 8, 10e-6, 2e9)


# ---- config 5 units: multi-notch DRAG ----------------------------------------------
@case
def cfg5_drag_sin(ns):
    w = 0.7 * ns.drag_sin(87e6, 30e-9, plateau=0, delta=1e6,
                          block_freq=(-250e6, ), phase=0.3, t0=100e-9)
    return _sampled(w, 0, 0.4e-6, 5e9)


@case
def cfg5_drag_sin_plateau(ns):
    w = ns.drag_sin(120e6, 24e-9, plateau=16e-9, delta=-2e6,
                    block_freq=(-250e6, 310e6, -95e6), phase=1.1, t0=50e-9)
    return _sampled(w, 0, 0.2e-6, 5e9)


@case
def cfg5_drag_sinx(ns):
    w = 0.9 * ns.drag_sinx(64e6, 30e-9, plateau=10e-9, delta=1e6,
                           block_freq=(-250e6, 180e6), phase=2.0, t0=100e-9)
    return _sampled(w, 0, 0.4e-6, 5e9)


@case
def cfg5_drag_none(ns):
    w = ns.drag_sin(50e6, 30e-9, t0=20e-9) + ns.drag(60e6, 25e-9, plateau=10e-9,
                                                     delta=2e6,
                                                     block_freq=-200e6,
                                                     phase=0.4, t0=80e-9)
    return _sampled(w, 0, 0.2e-6, 5e9)


# ---- every basis function, operators, clip, complex, exponents ---------------------
@case
def basis_zoo(ns):
    t = np.linspace(-3, 9, 4801)
    w = (ns.gaussian(2.0) >> 1) + 0.3 * (ns.sinc(40.0) >> 2.5)
    w = w + 0.2 * (ns.cosh(0.7) * ns.square(2.0) >> 4)
    w = w + 0.1 * (ns.sinh(0.9) * ns.square(1.0) >> 5.5)
    w = w + (ns.exp(-0.8) * (ns.square(2.0, edge=0.3) >> 7))
    w = w + 0.5 * (ns.mollifier(1.5) >> -1.5)
    w = w + 0.25 * (ns.mollifier(1.2, plateau=0.4, d=2) >> 0.2) * 1e-2
    return w, ('explicit', t)


@case
def basis_chirps(ns):
    t = np.linspace(-0.5, 10.5, 4401)
    w = ns.chirp(1, 2, 10, 4, 'linear') + 0.5 * ns.chirp(
        1.5, 0.5, 10, 0.3, 'exponential') - 0.25 * ns.chirp(
            1, 3, 10, 1.0, 'hyperbolic')
    return w, ('explicit', t)


@case
def basis_interp_poly(ns):
    t = np.linspace(-1, 6, 2801)
    pts = np.sin(np.linspace(0, 3, 17))**2
    w = ns.samplingPoints(0.5, 4.5, pts) + 0.1 * (ns.poly(
        [1, -1 / 2, 1 / 6, -1 / 12]) * ns.square(3.0) >> 2)
    w = w + ns.interp([0.0, 1.0, 2.5, 4.0, 5.0], [0.0, 1.0, -0.5, 0.25, 0.0])
    return w, ('explicit', t)


@case
def basis_dgauss_drag(ns):
    t = np.linspace(-4, 6, 5001)
    w = ns.gaussian(2.0, d=1) * 0.2 + (ns.gaussian(2.0, plateau=1.0, d=3) >> 2
                                       ) * 0.01
    w = w + ns.drag(1.3, 2.0, plateau=0.5, delta=0.1, block_freq=-0.7,
                    phase=0.4, t0=-3.5)
    w = w + ns.D(ns.gaussian(1.5) >> 4, 2) * 0.05
    return w, ('explicit', t)


@case
def ops_powers(ns):
    t = np.linspace(-2, 2, 1601)
    base = ns.cos(3.0, 0.2) + 1.5
    w = base**3 + (ns.exp(0.3)**-2) * (ns.square(3.0)) + (ns.gaussian(2.0)
                                                           **0.5)
    return w, ('explicit', t)


@case
def clip_minmax(ns):
    t = np.linspace(-2, 2, 1601)
    w = ns.cut(2.0 * ns.sin(4.0) * ns.square(3.0), min=-1.2, max=0.9)
    return w, ('explicit', t)


@case
def complex_amp(ns):
    t = np.linspace(-2, 2, 1001)
    w = 1j * (ns.cos(9) >> 1) + 1 * (ns.cos(9) >> 2) - 1j * (ns.cos(9) >> 3)
    return w, ('explicit', t)


@case
def complex_exp(ns):
    t = np.linspace(-2, 2, 1001)
    w = 2 * (ns.exp(1.01 + 22j)**2 << 1) * ns.exp(1.01 + 22j)
    return w, ('explicit', t)


@case
def stack_complex_amp(ns):
    # a stack accumulates in complex128 and returns the REAL part (waveform.py:681-693)
    t = np.linspace(-3, 3, 1201)
    wl = [(0.5 + 0.25j) * ns.cosPulse(1.5), 1j * (ns.gaussian(1.0) >> 0.7),
          (ns.cos(7.0) * ns.square(2.0)) >> -0.5]
    w = ns.WaveVStack(wl) + 0.125
    return w, ('explicit', t)


@case
def complex_zero_imag(ns):
    # (1j*1j) * w: amplitude (-1+0j) -> the reference returns complex128 with a zero imaginary part
    t = np.linspace(-2, 2, 801)
    w = (1j * 1j) * (ns.cos(5.0) * ns.square(2.0))
    return w, ('explicit', t)


@case
def complex_unsampled(ns):
    # the only complex segment lies outside the sampled range -> float64 output
    t = np.linspace(-2, 2, 801)
    w = ns.gaussian(1.0) + 1j * (ns.square(1.0) >> 10)
    return w, ('explicit', t)


@case
def filters_complex(ns):
    # sample-time IIR of a complex channel: scipy filters both planes
    from scipy.signal import butter, tf2sos
    b, a = butter(2, 40.0, 'lowpass', fs=1000)
    w = (1 + 0.5j) * (ns.square(0.8) >> 0.1) + 0.25
    w.filters = (tf2sos(b, a), 0.25)
    return _sampled(w, -1, 1, 1000)


@case
def boundary_hits(ns):
    # abscissae that coincide exactly with segment bounds: half-open [lo, hi)
    t = np.arange(-8, 9) * 0.25
    w = ns.square(2.0) + 0.5 * (ns.square(1.0) >> 0.5) + ns.step(0) * 0.125
    return w, ('explicit', t)


@case
def stack_ops(ns):
    t = np.linspace(-10, 10, 1001)
    wl = [ns.cos(1), ns.sin(2), ns.gaussian(3),
          ns.poly([1, -1 / 2, 1 / 6, -1 / 12])]
    w = (ns.WaveVStack(wl) * ns.sin(2) + 3) >> 0.6
    return w, ('explicit', t)


@case
def sample_cos(ns):
    # reference tests/test_waveform.py:13-16 grid: start -10, stop 10.02, rate 50
    return _sampled(ns.cos(1), -10, 10.02, 50)


@case
def filters_step(ns):
    # reference tests/test_waveform.py:169-186
    from scipy.signal import butter, tf2sos
    b, a = butter(3, 4.0, 'lowpass', fs=1000)
    w = ns.step(0)
    w.filters = (tf2sos(b, a), 0)
    return _sampled(w, -1, 1, 1000)


@case
def filters_flux_expdecay(ns):
    # config 4 pipeline, sample-time IIR with a non-zero initial value
    w, g = flux_channel(ns, np.random.default_rng(20260044), 6, 5e-6, 2e9)
    sos = np.array([[0.99015614, -1.97372497, 0.98357705, 1., -1.99346203,
                     0.99347026]])
    w = w + 0.1
    w.start, w.stop, w.sample_rate = 0, 5e-6, 2e9
    w.filters = (sos, 0.1)
    return w, g


def random_program(ns, seed):
    """Seeded random sums/products over the analytic basis functions with
    shifts, scalar factors and jittered bounds."""
    rng = np.random.default_rng(seed)
    w = ns.zero()
    for _ in range(int(rng.integers(3, 9))):
        kind = int(rng.integers(7))
        width = rng.uniform(0.2, 1.5)
        if kind == 0:
            p = ns.cosPulse(width)
        elif kind == 1:
            p = ns.gaussian(width)
        elif kind == 2:
            p = ns.square(width, edge=rng.uniform(0.02, 0.1) * width)
        elif kind == 3:
            p = ns.square(width, edge=0.1 * width, type='cos')
        elif kind == 4:
            p = ns.coshPulse(width, eps=rng.uniform(0.5, 3))
        elif kind == 5:
            p = ns.square(width, edge=0.2 * width, type='linear')
        else:
            p = ns.gaussian(width, plateau=rng.uniform(0.1, 0.5))
        p = rng.uniform(-1, 1) * p >> rng.uniform(-3, 3)
        if rng.random() < 0.6:
            p = ns.mixing(p, freq=rng.uniform(-8, 8),
                          phase=rng.uniform(0, 6.28),
                          DRAGScaling=rng.uniform(0.001, 0.02))[int(
                              rng.integers(2))]
        w = w + p
    return w


for _seed in range(4):

    def _mk(seed):

        def fn(ns):
            return random_program(ns, 7000 + seed), ('explicit',
                                                     np.linspace(
                                                         -5, 5, 4001))

        fn.__name__ = f'random_{seed}'
        return fn

    case(_mk(_seed))
