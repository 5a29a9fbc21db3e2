"""Multi-notch DRAG envelopes ``drag_sin`` / ``drag_sinx`` (basis ids 16/17).

Drop-in for /root/reference/waveforms/multy_drag.py (builders :180-193,
:216-232).  The reference evaluates the envelopes with NumPy per call
(:30-155); here everything that does not depend on the sample — the notch
matrices ``B``, the derivative table of sin^m, the normalisation, the tab
polynomials — is computed once on the host when the factor is lowered
(``pack_drag_sin`` / ``pack_drag_sinx``) and the device kernel
(csrc/wfm_multidrag.cuh) evaluates the remaining per-sample part.
"""
from __future__ import annotations

import math

import numpy as np

from ._algebra import (NDIGITS, DeviceBasis, _zero, basic_wave, inf, pi,
                       registerBaseFunc)
from .lowering import _f, register_packer
from .waveform import Waveform


# -- sample-independent pieces (same maths as multy_drag.py:9-27, :76-87) ------
def notch_series(bs):
    """M_k = sum over k-subsets of prod [[0,b],[-b,0]]: the 2x2 coefficients that
    combine the k-th envelope derivative into (Omega_x, Omega_y)."""
    acc = np.zeros([len(bs) + 1, 2, 2])
    acc[0] = np.identity(2)
    for b in bs:
        rot = np.array([[0, b], [-b, 0]])
        acc[1:] = acc[1:] + acc[:-1] @ rot
    return acc


def sin_power_derivatives(m: int, n: int, a: float = 1):
    """Row i: coefficients c[p] with d^i/dt^i sin^m(a t) =
    sum_p c[p] sin^p(a t) * (cos(a t) if p odd else 1)."""
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


def tab_polynomial(f: np.ndarray, x: float):
    """Polynomial 1 + sum_l c_l tau^(m+l) whose derivatives 0..m-1 at tau = x
    match ``f`` (the reference solves this with scipy.linalg.inv,
    multy_drag.py:76-87)."""
    from scipy.linalg import inv
    rhs = np.copy(f)
    rhs[0] -= 1
    m = f.shape[0]
    C_mat = np.zeros([m, m])
    for n in range(m):
        for l in range(m):
            C_mat[n, l] += (x**(m + l - n)) * math.factorial(
                m + l) / math.factorial(m + l - n)
    sol = inv(C_mat) @ rhs
    return np.poly1d([*np.flip(sol), *np.zeros_like(f[:-1]), 1])


def _notch_setup(width, delta, block_freq):
    bs, m = [], 2
    if isinstance(block_freq, float):
        block_freq = (block_freq, )
    if block_freq is not None:
        bs = 1 / np.pi / 2 / (np.array(block_freq) - delta)
        m = max((len(bs) + 2) >> 1 << 1, m)
    B = notch_series(bs)
    o = np.pi / width
    Amat = sin_power_derivatives(m, len(bs), o)
    return B, Amat, o, m, len(bs)


def _common_pool(t0, freq, width, delta, phase, plateau, B, Amat, m, norm):
    k1 = 2 * np.pi * (freq + delta)
    k2 = 2 * np.pi * delta * t0 + phase
    tm1 = t0 + width / 2
    tm2 = t0 + plateau + width / 2
    # G[j, p] = sum_i B[i, j, 0] * A[i, p]
    G = np.einsum('ij,ip->jp', B[:, :, 0], Amat) / norm
    # plateau region: the reference evaluates the sin-power rows at S = C = 0
    # (only p = 0 survives, 0**0 = 1) and then overwrites row 0 with 1
    rows = Amat[:, 0].copy()
    rows[0] = 1.0
    P = (B[:, :, 0] * rows[:, None]).sum(axis=0) / norm
    return [_f(k1), _f(k2), _f(tm1), _f(tm2), _f(plateau),
            float(m), float(P[0]), float(P[1]), *G[0].tolist(), *G[1].tolist()]


def pack_drag_sin(args):
    t0, freq, width, delta, block_freq, phase, *rest = args
    plateau = rest[0] if rest else 0
    B, Amat, o, m, nb = _notch_setup(width, delta, block_freq)
    peak = np.ones([m + 1])
    peak[1::2] = 0
    peak = Amat @ peak
    coe = np.einsum('ijk,ki->j', B, np.array([peak, np.zeros_like(peak)]))
    coeff = np.sqrt(np.sum(np.abs(coe)**2))
    pool = _common_pool(t0, freq, width, delta, phase, plateau, B, Amat, m,
                        coeff)
    pool += [0.0, 0.0, 0.0, 0.0]  # no tabs: tl, tr, half width, rows
    return _f(t0), float(o), tuple(pool)


def pack_drag_sinx(args):
    t0, freq, width, delta, block_freq, phase, *rest = args
    plateau = rest[0] if len(rest) > 0 else 0
    tab = rest[1] if len(rest) > 1 else 0.618
    B, Amat, o, m, nb = _notch_setup(width, delta, block_freq)

    def edge(sign):
        arg = o * (1 + sign * tab) * width / 2
        v = np.sin(arg)**np.arange(m + 1)
        v[1::2] = v[1::2] * np.cos(arg)
        return tab_polynomial(Amat @ v, sign * tab * width / 2)

    left, right = edge(-1), edge(+1)
    pool = _common_pool(t0, freq, width, delta, phase, plateau, B, Amat, m,
                        1.0)
    tl = t0 + width / 2 - tab * width / 2
    tr = t0 + plateau + width / 2 + tab * width / 2
    rows = nb + 1
    L = len(left.coeffs)
    pool += [_f(tl), _f(tr), float(width / 2), float(rows), float(L)]
    pool += B[:, 0, 0].tolist() + B[:, 1, 0].tolist()
    for poly in (left, right):
        for n in range(rows):
            c = np.atleast_1d(np.polyder(poly, m=n).coeffs)
            pool += [0.0] * (L - len(c)) + [float(v) for v in c]
    return _f(t0), float(o), tuple(pool)


DRAG_SIN = registerBaseFunc(DeviceBasis('DRAG_SIN'))
DRAG_SINX = registerBaseFunc(DeviceBasis('DRAG_SINX'))
register_packer(DRAG_SIN, pack_drag_sin)
register_packer(DRAG_SINX, pack_drag_sinx)


def _envelope(type_id, t0, width, plateau, *args):
    return Waveform(seq=(_zero, basic_wave(type_id, *args), _zero),
                    bounds=(round(t0, NDIGITS),
                            round(t0 + width + plateau, NDIGITS), +inf))


def drag_sin(freq, width, plateau=0, delta=0, block_freq=None, phase=0, t0=0):
    phase += pi * delta * (width + plateau)
    if isinstance(block_freq, float):
        block_freq = (block_freq, )
    return _envelope(DRAG_SIN, t0, width, plateau, t0, freq, width, delta,
                     block_freq, phase, plateau)


def drag_sinx(freq, width, plateau=0, delta=0, block_freq=None, phase=0, t0=0,
              tab=0.618):
    phase += pi * delta * (width + plateau)
    if isinstance(block_freq, float):
        block_freq = (block_freq, )
    return _envelope(DRAG_SINX, t0, width, plateau, t0, freq, width, delta,
                     block_freq, phase, plateau, tab)
