"""Host-side symbolic algebra for piecewise waveforms.

This is the Python half of the drop-in: it produces exactly the nested-tuple
representation the reference's Cython module produces
(/root/reference/waveforms/_waveform.pyx:15-48), because that representation is
(a) what the reference's golden ``tolist``/``totree`` vectors pin and (b) what
the device IR is lowered from (see ``lowering.py``).

Data model (all plain hashable tuples of Python scalars):

    expression  E = (terms, amps)          sum_i amps[i] * terms[i]
    term        T = (factors, exponents)   prod_k factors[k] ** exponents[k]
    factor      F = (type_id, *args, shift)   basis_fn(t - shift, *args)
    ZERO        = ((), ())                 the empty sum
    UNIT (as a term) = ((), ())            the empty product

Nothing in this module evaluates samples: sampling is the CUDA path
(``engine.py``).  Function numbering 1..15 follows
/root/reference/waveforms/_waveform.pyx:374-388; 16/17 are assigned by
``multy_drag.py`` exactly as the reference does on import.
"""
from __future__ import annotations

import pickle
from bisect import bisect_left
from itertools import product as _cartesian
from math import comb as _comb

import numpy as np
from numpy import e, inf, pi  # noqa: F401  (re-exported like the reference)

NDIGITS = 15  # _waveform.pyx:9 — bounds are rounded to 15 decimals on shift

ZERO = ((), ())
UNIT = ((), ())
_zero = ZERO  # reference spelling


def _const(c):
    """Constant expression; 0 collapses to ZERO (_waveform.pyx:29-32)."""
    return ZERO if c == 0 else ((UNIT, ), (c, ))


_one = _const(1.0)
_half = _const(1 / 2)
_two = _const(2.0)
_pi = _const(pi)
_two_pi = _const(2 * pi)
_half_pi = _const(pi / 2)


def is_const(x):
    """_waveform.pyx:43-44"""
    return x == ZERO or x[0] == (UNIT, )


def basic_wave(Type, *args, shift=0):
    """Single factor, exponent 1, amplitude 1.0 (_waveform.pyx:47-48)."""
    return ((((Type, *args, shift), ), (1, )), ), (1.0, )


# ---------------------------------------------------------------------------
# sorted (key, value) multiset merge — the one primitive behind add and mul.
# Semantics follow _waveform.pyx:51-65 including the [lo, hi) search window.
# ---------------------------------------------------------------------------
def _merge_into(keys, vals, key, val, lo, hi):
    pos = bisect_left(keys, key, lo, hi)
    if pos < hi and keys[pos] == key:
        val = val + vals[pos]
        if val == 0:
            del keys[pos]
            del vals[pos]
            return pos, hi - 1
        vals[pos] = val
        return pos, hi
    keys.insert(pos, key)
    vals.insert(pos, val)
    return pos, hi + 1


def add(x, y):
    """x + y on (keys, values) pairs; works both for expressions (terms, amps)
    and for terms (factors, exponents) — the latter is how a product merges
    equal factors by adding exponents (_waveform.pyx:82-88)."""
    keys, vals = list(x[0]), list(x[1])
    lo, hi = 0, len(keys)
    for key, val in zip(y[0], y[1]):
        lo, hi = _merge_into(keys, vals, key, val, lo, hi)
    return tuple(keys), tuple(vals)


def mul(x, y):
    """Distribute x*y term by term in itertools.product order
    (_waveform.pyx:68-79).  The search window follows the last insertion, as in
    the reference, so results are bit-identical structurally."""
    keys, vals = [], []
    lo = hi = 0
    for (ta, tb), (va, vb) in zip(_cartesian(x[0], y[0]),
                                  _cartesian(x[1], y[1])):
        amp = va * vb
        if amp == 0:
            continue
        lo, hi = _merge_into(keys, vals, add(ta, tb), amp, lo, hi)
    return tuple(keys), tuple(vals)


def shift(x, time):
    """Delay every factor by ``time`` (_waveform.pyx:91-102)."""
    if is_const(x):
        return x
    moved = []
    for factors, exponents in x[0]:
        moved.append((tuple((*f[:-1], f[-1] + time) for f in factors),
                      exponents))
    return tuple(moved), x[1]


def pow(x, n):
    """_waveform.pyx:105-127"""
    if x == ZERO:
        return ZERO
    if n == 0:
        return _one
    if is_const(x):
        return _const(x[1][0]**n)
    if len(x[0]) == 1:
        (factors, exponents), = x[0]
        amp, = x[1]
        return (((factors, tuple(n * m for m in exponents)), ), (amp**n, ))
    assert isinstance(n, int) and n > 0
    acc = _one
    for _ in range(n):
        acc = mul(acc, x)
    return acc


# ---------------------------------------------------------------------------
# piecewise merges
# ---------------------------------------------------------------------------
def merge_waveform(b1, s1, b2, s2, oper):
    """Two-pointer sweep over two piecewise functions whose bound lists both end
    in +inf; equal neighbouring results coalesce (_waveform.pyx:216-235)."""
    bounds, seq = [], []
    i = j = 0
    n1, n2 = len(b1), len(b2)
    while i < n1 or j < n2:
        piece = oper(s1[i], s2[j])
        edge = min(b1[i], b2[j])
        if seq and piece == seq[-1]:
            bounds[-1] = edge
        else:
            bounds.append(edge)
            seq.append(piece)
        step_i = edge == b1[i]
        step_j = edge == b2[j]
        i += step_i
        j += step_j
    return tuple(bounds), tuple(seq)


def wave_sum(waves):
    """n-ary sum of piecewise functions (_waveform.pyx:172-213)."""
    if not waves:
        return ((+inf, ), (ZERO, ))
    bounds, seq = waves[0]
    if len(waves) == 1:
        return bounds, seq
    bounds, seq = list(bounds), list(seq)

    for ob, os_ in waves[1:]:
        if len(ob) == 1:
            seq = [add(s, os_[0]) for s in seq]
        elif len(bounds) == 1:
            base = seq[0]
            bounds = list(ob)
            seq = [add(base, s) for s in os_]
        else:
            lo = 0
            for b, s in zip(ob, os_):
                pos = bisect_left(bounds, b, lo=lo)
                if bounds[pos] > b:
                    bounds.insert(pos, b)
                    seq.insert(pos, s if pos == 0 else add(s, seq[pos]))
                    last = pos - 1
                else:
                    last = pos
                for k in range(lo + 1, last + 1):
                    seq[k] = add(seq[k], s)
                lo = pos

    k = 0
    while k < len(bounds) - 1:
        if seq[k] == seq[k + 1]:
            del seq[k]
            del bounds[k]
        else:
            k += 1
    return tuple(bounds), tuple(seq)


# ---------------------------------------------------------------------------
# basis-function registry (ids are process-global, assigned in order from 1;
# _waveform.pyx:264-288).  The callables stored here are *descriptors only*:
# the product never evaluates them on the CPU — the CUDA kernel implements
# ids 1..17 natively (csrc/wfm_basis.cuh).  User-registered callables get an id
# and can be carried symbolically, but lowering them raises (lowering.py).
# ---------------------------------------------------------------------------
_next_type_id = 1
_baseFunc = {}
_derivativeBaseFunc = {}
_baseFunc_latex = {}


class DeviceBasis:
    """Marker stored in ``_baseFunc`` for the built-in ids: evaluation happens
    on the GPU (csrc/wfm_basis.cuh), there is no host callable."""
    __slots__ = ('name', )

    def __init__(self, name):
        self.name = name

    def __call__(self, *a, **k):
        raise RuntimeError(
            f'basis function {self.name} is evaluated by the CUDA sampling '
            'kernel; there is no CPU implementation in waveforms_b200')

    def __repr__(self):
        return f'<DeviceBasis {self.name}>'


def registerBaseFunc(func):
    global _next_type_id
    type_id = _next_type_id
    _next_type_id += 1
    _baseFunc[type_id] = func
    return type_id


def packBaseFunc():
    return pickle.dumps(_baseFunc)


def updateBaseFunc(buf):
    _baseFunc.update(pickle.loads(buf))


def registerDerivative(Type, dFunc):
    _derivativeBaseFunc[Type] = dFunc


def registerBaseFuncLatex(Type, dFunc):
    _baseFunc_latex[Type] = dFunc


(LINEAR, GAUSSIAN, ERF, COS, SINC, EXP, INTERP, LINEARCHIRP, EXPONENTIALCHIRP,
 HYPERBOLICCHIRP, COSH, SINH, DRAG, MOLLIFIER, D_GAUSSIAN) = (registerBaseFunc(
     DeviceBasis(name)) for name in (
         'LINEAR', 'GAUSSIAN', 'ERF', 'COS', 'SINC', 'EXP', 'INTERP',
         'LINEARCHIRP', 'EXPONENTIALCHIRP', 'HYPERBOLICCHIRP', 'COSH', 'SINH',
         'DRAG', 'MOLLIFIER', 'D_GAUSSIAN'))


# ---------------------------------------------------------------------------
# symbolic derivative  (_waveform.pyx:238-261, table :391-480)
# ---------------------------------------------------------------------------
def _single(factor, n=1, amp=1):
    return ((((factor, ), (n, )), ), (amp, ))


def _d_LINEAR(shift, *args):
    return _one


def _d_GAUSSIAN(shift, *args):
    s, = args
    return (((((LINEAR, shift), (GAUSSIAN, s, shift)), (1, 1)), ),
            (-2 / s**2, ))


def _d_ERF(shift, *args):
    s, = args
    return _single((GAUSSIAN, s, shift), 1, 2 / s / np.sqrt(pi))


def _d_COS(shift, *args):
    w = args[0]
    return _single((COS, w, shift - pi / w / 2), 1, w)


def _d_SINC(shift, *args):
    # The reference indexes args[1], which SINC does not have, so D(sinc)
    # raises IndexError there too (_waveform.pyx:410-413); kept for parity.
    return (((((LINEAR, shift), (COS, *args, shift)), (-1, 1)),
             (((LINEAR, shift), (COS, args[0], args[1] - pi / 2, shift)),
              (-2, 1))), (1, -1 / args[0]))


def _d_EXP(shift, *args):
    return _single((EXP, *args, shift), 1, args[0])


def _d_INTERP(shift, start, stop, points):
    grad = tuple(np.gradient(np.asarray(points)))
    return _single((INTERP, start, stop, grad, shift), 1,
                   (len(points) - 1) / (stop - start))


def _d_COSH(shift, *args):
    return _single((SINH, *args, shift), 1, args[0])


def _d_SINH(shift, *args):
    return _single((COSH, *args, shift), 1, args[0])


def _d_LINEARCHIRP(shift, f0, f1, T, phi0):
    quad = (LINEARCHIRP, f0, f1, T, phi0 + pi / 2, shift)
    terms = ((((quad, ), (1, ))), (((LINEAR, shift), quad), (1, 1)))
    amps = (2 * pi * f0, 2 * pi * (f1 - f0) / T)
    if f0 == 0:
        return terms[1:], amps[1:]
    return terms, amps


def _d_EXPONENTIALCHIRP(shift, f0, alpha, phi0):
    return (((((EXP, alpha, shift), (EXPONENTIALCHIRP, f0, alpha,
                                     phi0 + pi / 2, shift)), (1, 1)), ),
            (2 * pi * f0, ))


def _d_HYPERBOLICCHIRP(shift, f0, k, phi0):
    return (((((LINEAR, shift - 1 / k), (HYPERBOLICCHIRP, f0, k, phi0 + pi / 2,
                                         shift)), (-1, 1)), ), (2 * pi * f0, ))


def _d_MOLLIFIER(shift, r, d):
    return _single((MOLLIFIER, r, d + 1, shift), 1, 1)


def _d_D_GAUSSIAN(shift, std_sq2, n):
    return _single((D_GAUSSIAN, std_sq2, n + 1, shift), 1, 1)


for _tid, _fn in ((LINEAR, _d_LINEAR), (GAUSSIAN, _d_GAUSSIAN), (ERF, _d_ERF),
                  (COS, _d_COS), (SINC, _d_SINC), (EXP, _d_EXP),
                  (INTERP, _d_INTERP), (COSH, _d_COSH), (SINH, _d_SINH),
                  (LINEARCHIRP, _d_LINEARCHIRP),
                  (EXPONENTIALCHIRP, _d_EXPONENTIALCHIRP),
                  (HYPERBOLICCHIRP, _d_HYPERBOLICCHIRP),
                  (MOLLIFIER, _d_MOLLIFIER), (D_GAUSSIAN, _d_D_GAUSSIAN)):
    registerDerivative(_tid, _fn)


def _D_base(factor):
    type_id, *args, shift_ = factor
    return _derivativeBaseFunc[type_id](shift_, *args)


def _D(x):
    """d/dt of an expression: sum rule, Leibniz on the first factor, power
    rule, then the table (_waveform.pyx:243-261)."""
    if is_const(x):
        return ZERO
    terms, amps = x
    if len(amps) > 1:
        return add(_D((terms[:1], amps[:1])), _D((terms[1:], amps[1:])))
    (factors, exponents), amp = terms[0], amps[0]
    if len(factors) > 1:
        head = (((factors[:1], exponents[:1]), ), (amp, ))
        rest = (((factors[1:], exponents[1:]), ), (1, ))
        return add(mul(head, _D(rest)), mul(_D(head), rest))
    f, n = factors[0], exponents[0]
    if n == 1:
        return mul(_D_base(f), _const(amp))
    return mul(_single(f, n - 1, n * amp), _D(_single(f, 1, 1)))


# ---------------------------------------------------------------------------
# simplifier (_waveform.pyx:483-654): cos products -> sums, exp fusion,
# gaussian powers, equal-frequency phasor merge.  Host-only; never applied on
# the device side (simplify is not value-neutral at 1e-12).
# ---------------------------------------------------------------------------
# NOTE (provenance): ``_cos_power_n`` and ``_trigMul_t`` TRANSCRIBE the arithmetic of the reference's
# ``_waveform.pyx:483-515`` operation for operation (only the names of the locals differ): ``==`` between waveforms and
# the ``tolist()`` goldens of the reference's tests compare the simplified tuples EXACTLY, so the order of every
# floating-point operation here is part of the drop-in contract, not a design choice of this package.  They are
# host-only API glue and carry no claim of original design.
def _cos_power_n(factor, n):
    _, w, sh = factor
    out = ZERO
    for k in range(0, n // 2 + 1):
        if n == 2 * k:
            out = add(out, _const(_comb(n, k) / 2**n))
        else:
            out = add(
                out,
                _single((COS, (n - 2 * k) * w, sh), 1,
                        _comb(n, k) / 2**(n - 1)))
    return out


def _trigMul_t(x, y, v):
    """cos(a)cos(b) = cos(a+b)/2 + cos(a-b)/2 (_waveform.pyx:497-515)."""
    _, w1, t1 = x
    _, w2, t2 = y
    if w2 > w1:
        t1, t2 = t2, t1
        w1, w2 = w2, w1
    hi = (COS, w1 + w2, (w1 * t1 + w2 * t2) / (w1 + w2))
    if w1 == w2:
        c = v * np.cos(w1 * t1 - w2 * t2) / 2
        if c == 0:
            return (((hi, ), (1, )), ), (0.5 * v, )
        return (UNIT, ((hi, ), (1, ))), (c, 0.5 * v)
    lo = (COS, w1 - w2, (w1 * t1 - w2 * t2) / (w1 - w2))
    if lo[1] > hi[1]:
        lo, hi = hi, lo
    return (((lo, ), (1, )), ((hi, ), (1, ))), (0.5 * v, 0.5 * v)


def _trigMul(x, y):
    if is_const(x) or is_const(y):
        return mul(x, y)
    out = ZERO
    for (ta, tb), (va, vb) in zip(_cartesian(x[0], y[0]),
                                  _cartesian(x[1], y[1])):
        v = va * vb
        rest = _one
        trig = []
        for f, n in zip(ta[0] + tb[0], ta[1] + tb[1]):
            if f[0] == COS:
                trig.append(f)
            else:
                rest = mul(rest, _single(f, n, 1))
        if len(trig) == 1:
            piece = mul(rest, _single(trig[0], 1, v))
        elif len(trig) == 2:
            piece = mul(rest, _trigMul_t(trig[0], trig[1], v))
        else:
            piece = mul(rest, _const(v))
        out = add(out, piece)
    return out


def _exp_trig_Reduce(term, v):
    trig = _one
    alpha = 0
    sh = 0
    kept_f, kept_n = [], []
    for f, n in zip(*term):
        if f[0] == COS:
            trig = _trigMul(trig, _cos_power_n(f, n))
        elif f[0] == EXP:
            moment = alpha * sh + n * f[1] * f[-1]
            alpha += n * f[1]
            sh = 0 if alpha == 0 else moment / alpha
        elif f[0] == GAUSSIAN and n != 1:
            kept_f.append((f[0], f[1] / np.sqrt(n), f[2]))
            kept_n.append(1)
        else:
            kept_f.append(f)
            kept_n.append(n)
    out = (((tuple(kept_f), tuple(kept_n)), ), (v, ))
    if alpha != 0:
        out = mul(out, basic_wave(EXP, alpha, shift=sh))
    return mul(out, trig)


def _get_freq(term):
    freq, sh = 0, 0
    rest_f, rest_n = [], []
    for f, n in zip(*term):
        if f[0] == COS:
            if freq != 0:
                raise ValueError("run _exp_trig_Reduce first")
            freq, sh = f[1], f[-1]
        else:
            rest_f.append(f)
            rest_n.append(n)
    return freq, sh, (tuple(rest_f), tuple(rest_n))


def _phasor_sum(a0, s0, a1, s1, freq):
    re = a0 * np.cos(freq * s0) + a1 * np.cos(freq * s1)
    im = a0 * np.sin(freq * s0) + a1 * np.sin(freq * s1)
    return np.sqrt(re**2 + im**2), np.arctan2(im, re) / freq


def simplify(expr, eps):
    """_waveform.pyx:588-635.  Note the reference tests ``abs(v) >= eps`` with
    the loop variable left over from the reduction pass; reproduced verbatim
    (``v_last``) because ``==`` and ``wave_eval`` results depend on it."""
    table = {}
    v_last = None
    for term, amp in zip(*expr):
        for t, v in zip(*_exp_trig_Reduce(term, amp)):
            v_last = v
            freq, sh, rest = _get_freq(t)
            v_r, v_i, sh_r, sh_i = v.real, v.imag, sh, sh
            key = (rest, freq)
            if key in table:
                p_r, psh_r, p_i, psh_i = table[key]
                if freq == 0:
                    v_r, v_i = v.real + p_r, v.imag + p_i
                else:
                    v_r, sh_r = _phasor_sum(p_r, psh_r, v_r, sh_r, freq)
                    v_i, sh_i = _phasor_sum(p_i, psh_i, v_i, sh_i, freq)
            table[key] = v_r, sh_r, v_i, sh_i

    out = ZERO
    for (rest, freq), (v_r, sh_r, v_i, sh_i) in table.items():
        if freq == 0 and abs(v_last) >= eps:
            amp = v_r if v_i == 0 else v_r + 1j * v_i
            out = add(out, ((rest, ), (amp, )))
            continue
        big_r, big_i = abs(v_r) >= eps, abs(v_i) >= eps
        if not big_r and not big_i:
            continue
        if big_r and not big_i:
            osc = _single((COS, freq, sh_r), 1, v_r)
        elif big_i and not big_r:
            osc = _single((COS, freq, sh_i), 1, v_i * 1j)
        else:
            osc = ((((COS, freq, sh_r), ), (1, )), (((COS, freq, sh_i), ),
                                                     (1, ))), (v_r, v_i * 1j)
        out = add(out, mul(((rest, ), (1, )), osc))
    return out


def filter(expr, low, high, eps):
    """Keep spectral components in [low, high) (_waveform.pyx:638-654)."""
    expr = simplify(expr, eps)
    out = ZERO
    for term, amp in zip(*expr):
        for f, n in zip(*term):
            if f[0] == COS:
                if low <= f[1] < high:
                    out = add(out, ((term, ), (amp, )))
                break
        else:
            if low <= 0:
                out = add(out, ((term, ), (amp, )))
    return out
