"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.

CPU restatement (NumPy/SciPy) of the reference's sampling hot path, used as the
checker in ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs.  Nothing under ``waveforms_b200/``
imports this module; the product path has no CPU fallback.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function
here against golden vectors produced by the unmodified reference
(feihoo87/waveforms 2.2.3, built from /root/reference in the build container
by ``tests/golden/make_golden.py``) and against the known-answer vectors the
reference's own tests hold (tests/test_waveform.py, tests/test_wavevstack.py).

The numeric primitives the reference calls live in un-vendored third-party
packages — numpy (>=1.13.3; 2.3.5 here) and scipy (>=1.0.0; 1.18.1 here),
/root/reference/pyproject.toml:32-38.  The oracle calls the SAME primitives
(np.cos, np.exp, np.sinc, np.interp, scipy.special.erf/hermite,
scipy.signal.sosfilt/lfilter/lfiltic/fftconvolve, np.fft) at the same call
sites and in the same order, so it is bit-identical to the reference wherever
NumPy itself is deterministic.

Each function cites the reference lines it follows (paths under
/root/reference/waveforms/).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.special as special
from scipy.signal import fftconvolve, lfilter, lfiltic, sosfilt

ZERO = ((), ())

LINEAR, GAUSSIAN, ERF, COS, SINC, EXP, INTERP, LINEARCHIRP, EXPONENTIALCHIRP, \
    HYPERBOLICCHIRP, COSH, SINH, DRAG, MOLLIFIER, D_GAUSSIAN, DRAG_SIN, \
    DRAG_SINX = range(1, 18)


# ---- basis functions: _waveform.pyx:290-371 ---------------------------------
def b_linear(t):
    return t


def b_gaussian(t, std_sq2):
    return np.exp(-(t / std_sq2)**2)


def b_d_gaussian(t, std_sq2, n):
    return (-1)**n / std_sq2**n * special.hermite(n)(t / std_sq2) * np.exp(
        -(t / std_sq2)**2)


def b_erf(t, std_sq2):
    return special.erf(t / std_sq2)


def b_cos(t, w):
    return np.cos(w * t)


def b_sinc(t, bw):
    return np.sinc(bw * t)


def b_exp(t, alpha):
    return np.exp(alpha * t)


def b_interp(t, start, stop, points):
    return np.interp(t, np.linspace(start, stop, len(points)), points)


def b_linearchirp(t, f0, f1, T, phi0):
    return np.sin(phi0 + 2 * np.pi * ((f1 - f0) / (2 * T) * t**2 + f0 * t))


def b_exponentialchirp(t, f0, alpha, phi0):
    return np.sin(phi0 + 2 * np.pi * f0 * (np.exp(alpha * t) - 1) / alpha)


def b_hyperbolicchirp(t, f0, k, phi0):
    return np.sin(phi0 + 2 * np.pi * f0 / k * np.log(1 + k * t))


def b_cosh(t, w):
    return np.cosh(w * t)


def b_sinh(t, w):
    return np.sinh(w * t)


def b_drag(t, t0, freq, width, delta, block_freq, phase):
    o = np.pi / width
    omega_x = np.sin(o * (t - t0))**2
    wt = 2 * np.pi * (freq + delta) * t - (2 * np.pi * delta * t0 + phase)
    if block_freq is None or block_freq - delta == 0:
        return omega_x * np.cos(wt)
    b = 1 / np.pi / 2 / (block_freq - delta)
    omega_y = -b * o * np.sin(2 * o * (t - t0))
    return omega_x * np.cos(wt) + omega_y * np.sin(wt)


def b_mollifier(t, r, d):
    x = t / r
    q = np.abs(x)**2 - 1
    with np.errstate(all='ignore'):
        if d == 0:
            return np.where(q >= 0, 0, np.exp(1 / q + 1))
        p = np.poly1d([-2, 0])
        for n in range(1, d):
            p = np.poly1d([1, 0, -2, 0, 1]) * p.deriv() + np.poly1d(
                [-4 * n, 0, 4 * n - 2, 0]) * p
        return np.where(q >= 0, 0,
                        np.exp(1 / q + 1) / (-q)**(2 * d)) * p(x) / r**d


# ---- multi-notch DRAG: multy_drag.py:9-174 -----------------------------------
def _notch_B(bs):
    out = np.zeros([len(bs) + 1, 2, 2])
    out[0] = np.array([np.identity(2)])
    for b in bs:
        out[1:] = out[1:] + out[:-1] @ np.array([[0, b], [-b, 0]])
    return out


def _sin_m_table(m, n, a=1):
    tab = np.zeros([n + 1, m + 1])
    tab[0, m] = 1
    for i in range(1, n + 1):
        if i % 2:
            tab[i][:-1] = tab[i - 1][1:] * np.arange(1, m + 1) * a
        else:
            tab[i][:-2] = tab[i - 2][2:] * np.arange(1, m) * np.arange(2, m + 1)
            tab[i] = tab[i] - tab[i - 2] * np.arange(m + 1)**2
            tab[i] = tab[i] * (a**2)
    return tab


def _sin_rows(t, t0, width, plateau, o, m, A_mat):
    conds = [
        t <= t0 + width / 2,
        (t > t0 + width / 2) * (t < t0 + plateau + width / 2),
        t >= t0 + plateau + width / 2
    ]
    ps = np.arange(m + 1)
    rows = (np.piecewise(t, conds, [
        lambda x: np.sin(o * (x - t0)), 0,
        lambda x: np.sin(o * (x - t0 - plateau))
    ]))**(ps.reshape([-1, 1]))
    rows[1::2] = rows[1::2] * np.piecewise(t, conds, [
        lambda x: np.cos(o * (x - t0)), 0,
        lambda x: np.cos(o * (x - t0 - plateau))
    ])
    return A_mat @ rows


def _notch_params(width, delta, block_freq):
    bs, m = [], 2
    if isinstance(block_freq, float):
        block_freq = (block_freq, )
    if block_freq is not None:
        bs = 1 / np.pi / 2 / (np.array(block_freq) - delta)
        m = max((len(bs) + 2) >> 1 << 1, m)
    o = np.pi / width
    return bs, m, o, _notch_B(bs), _sin_m_table(m, len(bs), o)


def _omega_sin(t, t0, width, delta, block_freq=None, plateau=0):
    bs, m, o, B_mat, A_mat = _notch_params(width, delta, block_freq)
    rows = _sin_rows(t, t0, width, plateau, o, m, A_mat)
    peak = np.ones([m + 1])
    peak[1::2] = 0
    peak = A_mat @ peak
    coe = np.einsum('ijk,ki->j', B_mat, np.array([peak, np.zeros_like(peak)]))
    coeff = np.sqrt(np.sum(np.abs(coe)**2))
    stack = np.array([rows, np.zeros_like(rows)])
    stack[0, 0][(t > t0 + width / 2) * (t < t0 + plateau + width / 2)] = 1
    return np.einsum('ijk,kim->jm', B_mat, stack) / coeff


def _tab_poly(f, x):
    from scipy.linalg import inv
    rhs = np.copy(f)
    rhs[0] -= 1
    m = f.shape[0]
    C_mat = np.zeros([m, m])
    for n in range(0, m):
        for l in range(0, m):
            C_mat[n, l] += (x**(m + l - n)) * math.factorial(
                m + l) / math.factorial(m + l - n)
    return np.poly1d([*np.flip(inv(C_mat) @ rhs), *np.zeros_like(f[:-1]), 1])


def _omega_sinx(t, t0, width, delta, block_freq=None, plateau=0, tab=0.618):
    bs, m, o, B_mat, A_mat = _notch_params(width, delta, block_freq)
    rows = _sin_rows(t, t0, width, plateau, o, m, A_mat)

    def edge(arg, x):
        v = np.sin(arg)**np.arange(m + 1)
        v[1::2] = v[1::2] * np.cos(arg)
        return _tab_poly(A_mat @ v, x)

    left = edge(o * (1 - tab) * width / 2, -tab * width / 2)
    right = edge(o * (1 + tab) * width / 2, tab * width / 2)
    stack = np.array([rows, np.zeros_like(rows)])
    stack[0, 0][(t > t0 + width / 2) * (t < t0 + plateau + width / 2)] = 1
    in_l = (t >= t0 + width / 2 - tab * width / 2) * (t <= t0 + width / 2)
    in_r = (t >= t0 + plateau + width / 2) * (
        t <= t0 + plateau + width / 2 + tab * width / 2)
    for n in range(0, len(bs) + 1):
        stack[0, n][in_l] = (np.polyder(left, m=n))(t[in_l] - t0 - width / 2)
        stack[0, n][in_r] = (np.polyder(right,
                                        m=n))(t[in_r] - t0 - plateau -
                                              width / 2)
    return np.einsum('ijk,kim->jm', B_mat, stack)


def b_drag_sin(t, t0, freq, width, delta, block_freq, phase, plateau=0):
    ox, oy = _omega_sin(t, t0, width, delta, block_freq, plateau)
    wt = 2 * np.pi * (freq + delta) * t - (2 * np.pi * delta * t0 + phase)
    return ox * np.cos(wt) + oy * np.sin(wt)


def b_drag_sinx(t, t0, freq, width, delta, block_freq, phase, plateau=0,
                tab=0.618):
    ox, oy = _omega_sinx(t, t0, width, delta, block_freq, plateau, tab)
    wt = 2 * np.pi * (freq + delta) * t - (2 * np.pi * delta * t0 + phase)
    return ox * np.cos(wt) + oy * np.sin(wt)


BASIS = {
    LINEAR: b_linear, GAUSSIAN: b_gaussian, ERF: b_erf, COS: b_cos,
    SINC: b_sinc, EXP: b_exp, INTERP: b_interp, LINEARCHIRP: b_linearchirp,
    EXPONENTIALCHIRP: b_exponentialchirp, HYPERBOLICCHIRP: b_hyperbolicchirp,
    COSH: b_cosh, SINH: b_sinh, DRAG: b_drag, MOLLIFIER: b_mollifier,
    D_GAUSSIAN: b_d_gaussian, DRAG_SIN: b_drag_sin, DRAG_SINX: b_drag_sinx,
}


# ---- evaluator: _waveform.pyx:130-169 -----------------------------------------
def eval_expr(expr, x, basis=BASIS):
    """_calc/_calc_m/_apply: sum of amp * prod(factor ** n) with the per-call
    factor memo; accumulators start from the ints 0 and 1."""
    memo = {}
    total = 0
    for (factors, exponents), amp in zip(*expr):
        prod = 1
        for f, n in zip(factors, exponents):
            if f not in memo:
                type_id, *args, shift = f
                memo[f] = basis[type_id](x - shift, *args)
            prod = prod * memo[f] if n == 1 else prod * memo[f]**n
        total = total + amp * prod
    return total


def calc_parts(bounds, seq, x, lo=-math.inf, hi=math.inf, basis=BASIS):
    """Segment k owns x[edges[k-1]:edges[k]], edges = searchsorted(x, bounds)
    (side='left'): half-open [bound[k-1], bound[k]).  Zero segments are skipped
    (and therefore never clipped)."""
    edges = np.searchsorted(x, bounds)
    parts, dtype = [], float
    start = 0
    for k, stop in enumerate(edges):
        if start < stop and seq[k] != ZERO:
            part = np.clip(eval_expr(seq[k], x[start:stop], basis), lo, hi)
            if isinstance(part, complex) or (isinstance(part, np.ndarray)
                                             and isinstance(part[0], complex)):
                dtype = complex
            parts.append((start, stop, part))
        start = stop
    return parts, dtype


def waveform_call(bounds, seq, x, lo=-math.inf, hi=math.inf, calc=calc_parts):
    """Waveform.__call__ + _fill_parts: waveform.py:524-552."""
    x = np.asarray(x, dtype=float)
    parts, dtype = calc(bounds, seq, x, lo, hi)
    out = np.zeros_like(x, dtype=dtype)
    for start, stop, part in parts:
        out[start:stop] += part
    return out


def stack_call(wlist, x, offset=0, shift=0, calc=calc_parts):
    """WaveVStack.__call__: waveform.py:679-693 (complex128 accumulator,
    members never clipped, real part returned)."""
    x = np.asarray(x, dtype=float)
    out = np.full_like(x, offset, dtype=np.complex128)
    if shift != 0:
        x = x - shift
    for bounds, seq in wlist:
        parts, _ = calc(bounds, seq, x)
        for start, stop, part in parts:
            out[start:stop] += part
    return out.real


def sample_grid(start, stop, sample_rate):
    """Waveform.sample's grid: waveform.py:190."""
    return np.arange(start, stop, 1 / sample_rate)


def apply_filters(sig, filters):
    """Sample-time IIR hook: waveform.py:193-203."""
    if filters is None:
        return sig
    sos, initial = filters
    sos = np.array(sos)
    if initial:
        return sosfilt(sos, sig - initial) + initial
    return sosfilt(sos, sig)


# ---- wire format: waveform.py:278-306 ---------------------------------------------
def parse_flat(l, pos=0):
    """Waveform._fromlist: [nseg, {bound, nsum, {amp, nmul, {n, nfun, *fun}}}]."""
    nseg = l[pos]
    pos += 1
    bounds, seq = [], []
    for _ in range(nseg):
        bound, nsum = l[pos:pos + 2]
        pos += 2
        terms, amps = [], []
        for _ in range(nsum):
            amp, nmul = l[pos:pos + 2]
            pos += 2
            fs, ns = [], []
            for _ in range(nmul):
                n, flen = l[pos:pos + 2]
                pos += 2
                ns.append(n)
                fs.append(tuple(l[pos:pos + flen]))
                pos += flen
            terms.append((tuple(fs), tuple(ns)))
            amps.append(amp)
        bounds.append(bound)
        seq.append((tuple(terms), tuple(amps)))
    return tuple(bounds), tuple(seq), pos


# ---- distortion.py apply functions ----------------------------------------------
def reflection_filter(f, A, tau):
    """distortion.py:188-205"""
    return (1 - A) / (1 - A * np.exp(-2j * np.pi * f * tau))


def reflection(sig, A, tau, sample_rate):
    """distortion.py:208-210"""
    freq = np.fft.fftfreq(len(sig), 1 / sample_rate)
    return np.fft.ifft(np.fft.fft(sig) * reflection_filter(freq, A, tau)).real


def correct_reflection(sig, A, tau, sample_rate):
    """distortion.py:218-221 (ndarray branch)"""
    freq = np.fft.fftfreq(len(sig), 1 / sample_rate)
    return np.fft.ifft(np.fft.fft(sig) / reflection_filter(freq, A, tau)).real


def combine_filters(filters):
    """distortion.py:226-244"""
    b, a = np.poly1d([1.0]), np.poly1d([1.0])
    for b_, a_ in filters:
        b = b * np.poly1d(b_)
        a = a * np.poly1d(a_)
    return b.coeffs, a.coeffs


def predistort(sig, filters=None, ker=None, initial=0.0):
    """distortion.py:289-337 with default initial_x/initial_y/zi."""
    if filters is not None:
        b, a = combine_filters(filters)
        zi = lfiltic(b, a, np.full((len(a) - 1, ), initial),
                     np.full((len(b) - 1, ), initial))
        sig, _ = lfilter(b, a, sig, zi=zi)
    if ker is None:
        return sig
    size = len(sig)
    sig = np.hstack((np.zeros_like(sig), sig, np.zeros_like(sig)))
    start = size + len(ker) // 2
    return fftconvolve(sig, ker, mode='full')[start:start + size]
