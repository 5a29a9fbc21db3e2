"""Vectorised pulse-train builder: parameter arrays -> flat device IR, without one
Python object per pulse.

The reference builds a sequence by algebra on ``Waveform`` objects; one
``mixing(amp * cosPulse(w) >> t0, ...)`` costs ~150 us of Python, a 4096-channel
randomized-benchmarking batch of depth 1000 costs minutes (``merge_waveform`` /
``add`` / ``mul``, /root/reference/waveforms/_waveform.pyx:68-88,216-235) — far more
than sampling it.  Here a pulse SHAPE is traced ONCE through the ordinary object API
with a symbolic start time (``PulseTemplate.trace``); every floating-point operation
the algebra applies to that start time (``shift + time`` in ``shift()``,
_waveform.pyx:91-102; ``round(bound + time, 15)`` in ``Waveform.__rshift__``,
waveform.py:508-511; ``shift - pi/w/2`` in the derivative table, _waveform.pyx:391-480)
is recorded as an expression and replayed with NumPy over the array of start times, in
the same order and precision, so the IR of every pulse is what ``lower()`` would have
produced from the object the reference API builds — bit for bit in every table except
the (non-semantic) order of the shared argument pool.  The trace is VERIFIED at
creation against an independent plain-float build at a second start time; a pulse whose
dependence on its start time the tracer cannot follow raises ``UntraceablePulse``
instead of producing a wrong table.

Channels are stacks of NON-overlapping pulses (what ``WaveVStack`` of a gate sequence
is); overlapping members need the general merge in ``lowering._merge_members``.
"""
from __future__ import annotations

import math

import numpy as np

from . import _algebra as A
from . import engine
from .lowering import (PACKERS, FACTOR_DT, REF_DT, SEGPTR_DT, TERM_DT, WAVE_COMPLEX,
                       WAVE_DT, LoweredBatch, _Pools, _lower_segment, _plan_slots)


class UntraceablePulse(ValueError):
    """The pulse's tables depend on its start time in a way the tracer did not record."""


class SymTime(float):
    """A float that remembers how it was computed from the symbolic start time ``t0``.
    Its value is the probe start time's, so comparisons, sorting and hashing inside the
    algebra behave exactly as for a plain float.  Only the operations the algebra
    applies to times are traceable (+, -, unary -, round); anything else raises."""
    __slots__ = ('expr', )

    def __new__(cls, value, expr=('t0', )):
        self = super().__new__(cls, value)
        self.expr = expr
        return self

    @staticmethod
    def _expr(v):
        return v.expr if isinstance(v, SymTime) else ('const', float(v))

    def __add__(self, o):
        return SymTime(float(self) + float(o), ('add', self.expr, self._expr(o)))

    def __radd__(self, o):
        return SymTime(float(o) + float(self), ('add', self._expr(o), self.expr))

    def __sub__(self, o):
        return SymTime(float(self) - float(o), ('sub', self.expr, self._expr(o)))

    def __rsub__(self, o):
        return SymTime(float(o) - float(self), ('sub', self._expr(o), self.expr))

    def __neg__(self):
        return SymTime(-float(self), ('neg', self.expr))

    def __pos__(self):
        return self

    def __round__(self, nd=None):
        if nd is None:
            raise UntraceablePulse('round() to an integer of a symbolic time')
        return SymTime(round(float(self), nd), ('round', self.expr, nd))

    def _untraceable(self, *_):
        raise UntraceablePulse('the pulse multiplies / divides / exponentiates its start time; '
                               'only +, - and round() of the start time can be traced')

    __mul__ = __rmul__ = __truediv__ = __rtruediv__ = __pow__ = __rpow__ = _untraceable
    __floordiv__ = __rfloordiv__ = __mod__ = __rmod__ = _untraceable


_round_ufuncs = {}


def _eval(expr, t0):
    """Replay a recorded expression over the float64 array ``t0`` (element-wise IEEE
    operations: identical to the Python float arithmetic of the trace)."""
    op = expr[0]
    if op == 't0':
        return t0
    if op == 'const':
        return np.float64(expr[1])
    if op == 'neg':
        return -_eval(expr[1], t0)
    if op == 'round':
        nd = expr[2]
        uf = _round_ufuncs.get(nd)
        if uf is None:  # Python's correctly rounded decimal round(), not np.round's scaling
            uf = _round_ufuncs[nd] = np.frompyfunc(lambda v, nd=nd: round(v, nd), 1, 1)
        return uf(np.asarray(_eval(expr[1], t0), dtype=np.float64)).astype(np.float64)
    a, b = _eval(expr[1], t0), _eval(expr[2], t0)
    return a + b if op == 'add' else a - b


_cos_uf = np.frompyfunc(math.cos, 1, 1)  # the libm calls lowering._emit_rows makes
_sin_uf = np.frompyfunc(math.sin, 1, 1)


class PulseTemplate:
    """The lowered tables of ONE pulse as a function of its start time."""

    def __init__(self, fn, probe=1.0e-6, check=2.37e-6):
        self.fn = fn
        w = fn(SymTime(probe))
        self._extract(w.bounds, w.seq)
        self._verify(check)

    @classmethod
    def trace(cls, fn, **kw):
        """``fn(t0) -> Waveform`` built with the ordinary object API, e.g.
        ``lambda t0: mixing(0.5 * cosPulse(20e-9) >> t0, freq=-80e6, phase=pi/2,
        DRAGScaling=4e-10)[0]``."""
        return cls(fn, **kw)

    # -- tracing -------------------------------------------------------------------
    def _extract(self, bounds, seq):
        if not bounds or bounds[-1] != math.inf or seq[-1] != A.ZERO or seq[0] != A.ZERO:
            raise UntraceablePulse('a pulse template must be zero before its first and after its '
                                   'last bound')
        self.bound_expr = [SymTime._expr(b) for b in bounds[:-1]]
        pools = _Pools()
        self.seg_fac, self.seg_term = [], []       # local CSR starts of the finite segments
        self.sym_shift = []                        # (fac row, expr)
        self.rot_rows = []                         # (fac row, local arg_off, w, base-shift expr)
        has_args = []                              # fac rows that own a block of the argument pool
        cplx = False
        for s in seq[:-1]:
            self.seg_fac.append(pools.n_fac)
            self.seg_term.append(pools.n_term)
            if s == A.ZERO:
                continue
            order, seen = [], set()
            for factors, _ in s[0]:
                for f in factors:
                    if f not in seen:
                        seen.add(f)
                        order.append(f)
            rows, _ = _plan_slots(order)  # the plan _lower_segment is about to emit
            base = pools.n_fac
            plain = (tuple((tuple((*f[:-1], float(f[-1])) for f in factors), expo) for factors, expo in s[0]), s[1])
            cplx = _lower_segment(pools, [plain]) or cplx
            for r, row in enumerate(rows):
                has_args.append(row[0] == 'rot' or
                                (row[0] == 'plain' and bool(PACKERS[row[1][0]](row[1][1:-1])[2])))
                if row[0] == 'nop':
                    continue
                f = row[1]
                if isinstance(f[-1], SymTime):
                    self.sym_shift.append((base + r, f[-1].expr))
                if any(isinstance(a, SymTime) for a in f[1:-1]):
                    raise UntraceablePulse('a basis-function argument depends on the start time')
                if row[0] == 'rot':
                    self.rot_rows.append((base + r, pools.fac[base + r][1], f[1], SymTime._expr(row[3][-1])))
        self.n_seg = len(bounds) - 1
        self.facs = np.array(pools.fac, dtype=FACTOR_DT) if pools.fac else np.zeros(0, FACTOR_DT)
        self.terms = np.array(pools.term, dtype=TERM_DT) if pools.term else np.zeros(0, TERM_DT)
        self.refs = np.array(pools.ref, dtype=REF_DT) if pools.ref else np.zeros(0, REF_DT)
        self.args = np.asarray(pools.args, dtype=np.float64)
        self.seg_fac = np.asarray(self.seg_fac, dtype=np.int64)
        self.seg_term = np.asarray(self.seg_term, dtype=np.int64)
        self.has_args = np.asarray(has_args, dtype=np.int64)
        self.complex = cplx

    def instantiate(self, t0):
        """Tables of ``len(t0)`` instances: (bounds[P, n_seg], facs[P, nf], args[P, na]);
        ``arg_off`` stays local to the instance."""
        t0 = np.ascontiguousarray(t0, dtype=np.float64)
        # start times repeat across channels (a gate grid): evaluate the distinct ones only
        uniq, inverse = np.unique(t0, return_inverse=True)
        if len(uniq) <= len(t0) // 2:
            b, f, a = self.instantiate(uniq)
            return b[inverse], f[inverse], a[inverse]
        P = len(t0)
        bounds = np.empty((P, self.n_seg), dtype=np.float64)
        for j, e in enumerate(self.bound_expr):
            bounds[:, j] = _eval(e, t0)
        facs = np.tile(self.facs, P).reshape(P, len(self.facs))
        for r, e in self.sym_shift:
            facs['shift'][:, r] = _eval(e, t0)
        args = np.tile(self.args, P).reshape(P, len(self.args))
        for r, off, w, base_expr in self.rot_rows:
            # lowering._emit_rows: delta = w * (s_b - s_t); block (slot, s_b, delta, cos, sin)
            s_b = np.broadcast_to(_eval(base_expr, t0), (P, ))
            delta = w * (s_b - facs['shift'][:, r])
            args[:, off + 1] = s_b
            args[:, off + 2] = delta
            args[:, off + 3] = _cos_uf(delta).astype(np.float64)
            args[:, off + 4] = _sin_uf(delta).astype(np.float64)
        return bounds, facs, args

    def _verify(self, t_check):
        """Replay at a second start time against a plain-float build there."""
        w = self.fn(float(t_check))
        ref = PulseTemplate.__new__(PulseTemplate)
        ref._extract(w.bounds, w.seq)
        b, f, a = self.instantiate(np.array([t_check]))
        same = (ref.n_seg == self.n_seg and np.array_equal(np.array([float(x) for x in w.bounds[:-1]]), b[0])
                and np.array_equal(ref.facs, f[0]) and np.array_equal(ref.args, a[0])
                and np.array_equal(ref.terms, self.terms) and np.array_equal(ref.refs, self.refs))
        if not same:
            raise UntraceablePulse('the tables traced with a symbolic start time do not reproduce a '
                                   f'plain build at t0={t_check!r}: the pulse depends on its start '
                                   'time through an operation the tracer cannot follow')


def pulse_train_batch(templates, tmpl_idx, t0, start, stop, sample_rate) -> LoweredBatch:
    """One channel per row: channel ``c`` is the stack of pulses ``templates[tmpl_idx[c][k]]``
    started at ``t0[c][k]`` (time-ordered, non-overlapping), sampled on
    ``np.arange(start, stop, 1/sample_rate)`` — the ``LoweredBatch`` that
    ``lower([channel_grid(WaveVStack([fn(t) for ...]))])`` yields, built with NumPy.
    ``tmpl_idx`` / ``t0``: 2-D arrays or lists of 1-D arrays (ragged channels)."""
    n_ch = len(t0)
    counts = np.array([len(r) for r in t0], dtype=np.int64)
    T = np.concatenate([np.asarray(r, dtype=np.float64) for r in t0]) if n_ch else np.zeros(0)
    M = np.concatenate([np.asarray(r, dtype=np.int64) for r in tmpl_idx]) if n_ch else np.zeros(0, np.int64)
    if len(M) != len(T):
        raise ValueError('tmpl_idx and t0 differ in shape')
    P = len(T)
    ch_of = np.repeat(np.arange(n_ch), counts)
    ns_t = np.array([t.n_seg for t in templates], dtype=np.int64)
    nf_t = np.array([len(t.facs) for t in templates], dtype=np.int64)
    nt_t = np.array([len(t.terms) for t in templates], dtype=np.int64)
    nr_t = np.array([len(t.refs) for t in templates], dtype=np.int64)
    na_t = np.array([len(t.args) for t in templates], dtype=np.int64)

    def starts(per_pulse):
        out = np.zeros(P + 1, dtype=np.int64)
        np.cumsum(per_pulse, out=out[1:])
        return out

    fac_off, term_off = starts(nf_t[M]), starts(nt_t[M])
    ref_off, arg_off = starts(nr_t[M]), starts(na_t[M])
    if max(fac_off[-1], term_off[-1], ref_off[-1], arg_off[-1]) >= 2**31:
        raise ValueError('batch too large for 32-bit table indices; split it')
    seg_off = starts(ns_t[M])
    ch_first = np.zeros(n_ch + 1, dtype=np.int64)  # first pulse of every channel
    np.cumsum(counts, out=ch_first[1:])
    # every channel ends with one (+inf, zero) segment
    seg_pos = seg_off[:-1] + ch_of
    n_seg_full = int(seg_off[-1]) + n_ch
    seg_bound = np.empty(n_seg_full, dtype=np.float64)
    seg_fac = np.empty(n_seg_full + 1, dtype=np.int64)
    seg_term = np.empty(n_seg_full + 1, dtype=np.int64)
    first_edge = np.empty(P, dtype=np.float64)
    last_edge = np.empty(P, dtype=np.float64)
    facs = np.zeros(int(fac_off[-1]), dtype=FACTOR_DT)
    terms = np.zeros(int(term_off[-1]), dtype=TERM_DT)
    refs = np.zeros(int(ref_off[-1]), dtype=REF_DT)
    args = np.zeros(int(arg_off[-1]), dtype=np.float64)
    for m, tp in enumerate(templates):
        idx = np.nonzero(M == m)[0]
        if not len(idx):
            continue
        b, f, a = tp.instantiate(T[idx])
        first_edge[idx], last_edge[idx] = b[:, 0], b[:, -1]
        sp = seg_pos[idx][:, None] + np.arange(tp.n_seg)[None, :]
        seg_bound[sp] = b
        seg_fac[sp] = fac_off[idx][:, None] + tp.seg_fac[None, :]
        seg_term[sp] = term_off[idx][:, None] + tp.seg_term[None, :]
        if len(tp.facs):
            f['arg_off'] += (arg_off[idx][:, None] * tp.has_args[None, :]).astype(np.int32)
            facs[fac_off[idx][:, None] + np.arange(len(tp.facs))[None, :]] = f
        if len(tp.terms):
            t = np.tile(tp.terms, len(idx)).reshape(len(idx), len(tp.terms))
            t['ref_begin'] += ref_off[idx][:, None].astype(np.int32)
            terms[term_off[idx][:, None] + np.arange(len(tp.terms))[None, :]] = t
        if len(tp.refs):
            refs[ref_off[idx][:, None] + np.arange(len(tp.refs))[None, :]] = tp.refs[None, :]
        if len(tp.args):
            args[arg_off[idx][:, None] + np.arange(len(tp.args))[None, :]] = a
    # channel tails
    tail = seg_off[ch_first[1:]] + np.arange(n_ch)
    seg_bound[tail] = math.inf
    seg_fac[tail] = fac_off[ch_first[1:]]
    seg_term[tail] = term_off[ch_first[1:]]
    seg_fac[n_seg_full], seg_term[n_seg_full] = fac_off[-1], term_off[-1]
    # consecutive pulses of a channel: ordered, not overlapping; a shared edge appears once
    # (the union of the members' bounds is a set, lowering._merge_members)
    nxt = np.nonzero(ch_of[1:] == ch_of[:-1])[0] + 1  # pulses with a predecessor in their channel
    if len(nxt) and np.any(first_edge[nxt] < last_edge[nxt - 1]):
        bad = nxt[np.nonzero(first_edge[nxt] < last_edge[nxt - 1])[0][0]]
        raise ValueError(f'pulse {int(bad - ch_first[ch_of[bad]])} of channel {int(ch_of[bad])} starts before '
                         'its predecessor ends: overlapping members need WaveVStack + lower()')
    keep = np.ones(n_seg_full + 1, dtype=bool)
    dup = nxt[first_edge[nxt] == last_edge[nxt - 1]]
    keep[seg_pos[dup]] = False
    seg_bound = seg_bound[keep[:-1]]
    seg_ptr = np.zeros(int(keep.sum()), dtype=SEGPTR_DT)
    seg_ptr['fac'] = seg_fac[keep]
    seg_ptr['term'] = seg_term[keep]
    # per-channel rows
    kept_before = np.zeros(n_seg_full + 1, dtype=np.int64)
    np.cumsum(keep[:-1], out=kept_before[1:])
    ch_seg_lo = seg_off[ch_first[:-1]] + np.arange(n_ch)
    ch_seg_hi = tail + 1
    grid = engine.arange_grid(start, stop, 1 / sample_rate)
    waves = np.zeros(n_ch, dtype=WAVE_DT)
    waves['t0'], waves['delta'], waves['n'] = grid.t0, grid.delta, grid.n
    waves['out_off'] = np.arange(n_ch, dtype=np.int64) * ((grid.n + 3) & ~3)
    waves['seg_begin'] = kept_before[ch_seg_lo]
    waves['n_seg'] = kept_before[ch_seg_hi] - kept_before[ch_seg_lo]
    cplx_t = np.array([t.complex for t in templates], dtype=bool)
    ch_cplx = np.zeros(n_ch, dtype=bool)
    if P:
        np.logical_or.at(ch_cplx, ch_of, cplx_t[M])
    waves['flags'] = np.where(ch_cplx, WAVE_COMPLEX, 0)
    return LoweredBatch(waves=waves, seg_bound=seg_bound, seg_ptr=seg_ptr, facs=facs, terms=terms,
                        refs=refs, args=args, x=np.zeros(0, np.float64),
                        total_samples=int(n_ch * ((grid.n + 3) & ~3)), any_complex=bool(ch_cplx.any()))
