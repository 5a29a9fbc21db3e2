"""Drop-in ``Waveform`` object model (host side, Python).

Mirrors the public surface of /root/reference/waveforms/waveform.py —
``Waveform`` (:125-635), ``WaveVStack`` (:638-844), the builders (:886-895,
:1078-1484), ``D`` (:1055-1071) and ``mixing`` (:1487-1527) — but sampling
(``__call__`` / ``sample``) does not touch NumPy ufuncs: the piecewise program
is lowered to the flat device IR (``lowering.py``) and evaluated by the sm_100a
kernel behind the C-ABI (``engine.py`` → ``csrc/``).  There is no CPU fallback;
without the CUDA library these calls raise.

Out of scope here: LaTeX rendering and audio playback.
"""
from __future__ import annotations

from typing import Iterable

import numpy as np
from numpy import e, inf, pi  # noqa: F401

from ._algebra import (_D, COS, COSH, D_GAUSSIAN, DRAG, ERF, EXP,
                       EXPONENTIALCHIRP, GAUSSIAN, HYPERBOLICCHIRP, INTERP,
                       LINEAR, LINEARCHIRP, MOLLIFIER, NDIGITS, SINC, SINH,
                       _baseFunc, _baseFunc_latex, _const, _half, _one, _zero,
                       add, basic_wave, filter, is_const, merge_waveform, mul,
                       pow, registerBaseFunc, registerBaseFuncLatex,
                       registerDerivative, shift, simplify, wave_sum)

_filter_expr = filter
_simplify_expr = simplify


def _rnd(v):
    return round(v, NDIGITS)


def _as_sos(sos):
    """Same normalisation the reference applies before sosfilt
    (waveform.py:195-198)."""
    if not isinstance(sos, np.ndarray):
        sos = np.array(sos)
    elif not sos.flags.writeable:
        sos = sos.copy()
    return sos


class Waveform:
    __slots__ = ('bounds', 'seq', 'max', 'min', 'start', 'stop', 'sample_rate',
                 'filters', 'label')

    def __init__(self, bounds=(+inf, ), seq=(_zero, ), min=-inf, max=inf):
        self.bounds = bounds
        self.seq = seq
        self.max = max
        self.min = min
        self.start = None
        self.stop = None
        self.sample_rate = None
        self.filters = None
        self.label = None

    # -- support ----------------------------------------------------------
    @staticmethod
    def _begin(bounds, seq):
        for k, piece in enumerate(seq):
            if piece != _zero:
                return -inf if k == 0 else bounds[k - 1]
        return inf

    @staticmethod
    def _end(bounds, seq):
        for k in range(len(seq) - 1, -1, -1):
            if seq[k] != _zero:
                return inf if k == len(seq) - 1 else bounds[k]
        return -inf

    @property
    def begin(self):
        b = self._begin(self.bounds, self.seq)
        return b if self.start is None else max(self.start, b)

    @property
    def end(self):
        e_ = self._end(self.bounds, self.seq)
        return e_ if self.stop is None else min(self.stop, e_)

    # -- sampling: the hot path (reference waveform.py:173-257, :529-563) ---
    def _channel(self):
        from .lowering import Channel
        return Channel(members=[(self.bounds, self.seq)],
                       clip=(self.min, self.max))

    def sample(self,
               sample_rate=None,
               out: np.ndarray | None = None,
               chunk_size=None,
               function_lib=None,
               filters=None) -> np.ndarray | Iterable[np.ndarray]:
        if sample_rate is None:
            sample_rate = self.sample_rate
        if self.start is None or self.stop is None or sample_rate is None:
            raise ValueError(
                f'Waveform is not initialized. {self.start=}, {self.stop=}, {sample_rate=}'
            )
        if filters is None:
            filters = self.filters
        from . import engine
        engine.check_function_lib(function_lib)
        if chunk_size is not None:
            return self._sample_iter(sample_rate, chunk_size, out, filters)
        grid = engine.arange_grid(self.start, self.stop, 1 / sample_rate)
        sos = initial = None
        if filters is not None:
            sos, initial = filters
            sos = _as_sos(sos)
        return engine.sample_one(self._channel(), grid, out=out, sos=sos,
                                 initial=initial)

    def _sample_iter(self, sample_rate, chunk_size, out, filters):
        """Chunked streaming (reference waveform.py:209-257): per-chunk
        ``linspace(start, stop, size, endpoint=False)`` grid and the IIR state
        ``zi`` carried from chunk to chunk — on the device here."""
        from . import engine
        start = self.start
        start_n = 0
        sos = initial = zi = None
        if filters is not None:
            sos, initial = filters
            sos = _as_sos(sos)
            zi = np.zeros((sos.shape[0], 2))
        length = chunk_size / sample_rate
        chan = self._channel()
        while start < self.stop:
            if start + length > self.stop:
                length = self.stop - start
                stop = self.stop
                size = round((stop - start) * sample_rate)
            else:
                stop = start + length
                size = chunk_size
            grid = engine.linspace_grid(start, stop, size, endpoint=False)
            if sos is None:
                dst = None if out is None else out[start_n:]
                yield engine.sample_one(chan, grid, out=dst)
            else:
                sig, zi = engine.sample_one(chan, grid, sos=sos,
                                            initial=initial, zi=zi,
                                            return_zf=True)
                if out is not None:
                    out[start_n:start_n + size] = sig
                yield sig
            start = stop
            start_n += chunk_size

    def __call__(self, x, frag=False, out=None, accumulate=False,
                 function_lib=None):
        from . import engine
        engine.check_function_lib(function_lib)
        if isinstance(x, (int, float, complex)):
            return self.__call__(np.array([x]))[0]
        grid = engine.explicit_grid(x)
        if not frag:
            return engine.sample_one(self._channel(), grid, out=out,
                                     accumulate=accumulate, zero_out=True)
        parts = engine.sample_parts(self._channel(), grid)
        if out is None:
            return parts
        if accumulate:
            self._merge_parts(parts, out)
        else:
            out.clear()
            out.extend(parts)
        return out

    @staticmethod
    def _merge_parts(parts, out):
        raise NotImplementedError  # same as the reference (waveform.py:519)

    # -- flat / tree wire formats (reference waveform.py:259-382) ----------
    @staticmethod
    def _tolist(bounds, seq, ret=None):
        flat = [] if ret is None else ret
        flat.append(len(bounds))
        for (terms, amps), edge in zip(seq, bounds):
            flat += [edge, len(amps)]
            for (factors, exponents), amp in zip(terms, amps):
                flat += [amp, len(exponents)]
                for f, n in zip(factors, exponents):
                    flat += [n, len(f), *f]
        return flat

    @staticmethod
    def _fromlist(l, pos=0):

        def take(k):
            nonlocal pos
            if pos + k > len(l):
                raise ValueError('Invalid waveform format')
            chunk = tuple(l[pos:pos + k])
            pos += k
            return chunk

        nseg, = take(1)
        bounds, seq = [], []
        for _ in range(nseg):
            edge, nsum = take(2)
            bounds.append(edge)
            terms, amps = [], []
            for _ in range(nsum):
                amp, nmul = take(2)
                factors, exponents = [], []
                for _ in range(nmul):
                    n, flen = take(2)
                    exponents.append(n)
                    factors.append(take(flen))
                amps.append(amp)
                terms.append((tuple(factors), tuple(exponents)))
            seq.append((tuple(terms), tuple(amps)))
        return tuple(bounds), tuple(seq), pos

    @staticmethod
    def _filters_tolist(filters, flat):
        if filters is None:
            flat.append(None)
        else:
            sos, initial = filters
            coeffs = list(np.asarray(sos).reshape(-1))
            flat.append(len(coeffs))
            flat.extend(coeffs)
            flat.append(initial)

    @staticmethod
    def _filters_fromlist(l, pos, size):
        if size is None:
            return None, pos
        sos = np.array(l[pos:pos + size]).reshape(-1, 6)
        pos += size
        return (sos, l[pos]), pos + 1

    def tolist(self):
        flat = [self.max, self.min, self.start, self.stop, self.sample_rate]
        self._filters_tolist(self.filters, flat)
        return self._tolist(self.bounds, self.seq, flat)

    @classmethod
    def fromlist(cls, l):
        w = cls()
        w.max, w.min, w.start, w.stop, w.sample_rate, sos_size = l[:6]
        filt, pos = cls._filters_fromlist(l, 6, sos_size)
        if filt is not None:
            w.filters = filt
        w.bounds, w.seq, pos = cls._fromlist(l, pos)
        return w

    def totree(self):
        header = (self.max, self.min, self.start, self.stop, self.sample_rate,
                  self.filters)
        body = tuple(
            (edge,
             tuple((amp, tuple((n, f) for f, n in zip(factors, exponents)))
                   for (factors, exponents), amp in zip(terms, amps)))
            for (terms, amps), edge in zip(self.seq, self.bounds))
        return header, body

    @staticmethod
    def fromtree(tree):
        w = Waveform()
        header, body = tree
        w.max, w.min, w.start, w.stop, w.sample_rate, w.filters = header
        bounds, seq = [], []
        for edge, pieces in body:
            bounds.append(edge)
            terms = tuple((tuple(f for _, f in facs), tuple(n for n, _ in facs))
                          for _, facs in pieces)
            seq.append((terms, tuple(amp for amp, _ in pieces)))
        w.bounds, w.seq = tuple(bounds), tuple(seq)
        return w

    # -- symbolic transforms ----------------------------------------------
    def simplify(self, eps=1e-15):
        seq, bounds = [], []
        for piece, edge in zip(self.seq, self.bounds):
            piece = _simplify_expr(piece, eps)
            if seq and piece == seq[-1]:
                bounds[-1] = edge
            else:
                seq.append(piece)
                bounds.append(edge)
        return Waveform(tuple(bounds), tuple(seq))

    def filter(self, low=0, high=inf, eps=1e-15):
        return Waveform(
            self.bounds,
            tuple(_filter_expr(piece, low, high, eps) for piece in self.seq))

    def _comb(self, other, oper):
        return Waveform(*merge_waveform(self.bounds, self.seq, other.bounds,
                                        other.seq, oper))

    def __pow__(self, n):
        return Waveform(self.bounds, tuple(pow(piece, n) for piece in self.seq))

    def __add__(self, other):
        if not isinstance(other, Waveform):
            other = const(other)
        return self._comb(other, add)

    def __radd__(self, v):
        return const(v) + self

    def __or__(self, other):
        if isinstance(other, (int, float, complex)):
            other = const(other)
        self.marker + other.marker  # the reference evaluates (and discards) this
        return self._comb(other, lambda a, b: _one
                          if a != _zero or b != _zero else _zero)

    __ior__ = __or__

    def __and__(self, other):
        if isinstance(other, (int, float, complex)):
            other = const(other)
        self.marker + other.marker
        return self._comb(other, lambda a, b: _one
                          if a != _zero and b != _zero else _zero)

    __iand__ = __and__

    @property
    def marker(self):
        w = self.simplify()
        return Waveform(w.bounds,
                        tuple(_zero if s == _zero else _one for s in w.seq))

    def mask(self, edge: float = 0):
        """reference waveform.py:455-482"""
        w = self.marker
        inside = w.seq[0] == _zero
        bounds, seq = [], []
        if w.seq[0] == _zero:
            inside = False
            bounds.append(w.bounds[0] - edge)
            seq.append(_zero)
        for b, s in zip(w.bounds[1:], w.seq[1:]):
            if not inside and s != _zero:
                inside = True
                bounds.append(b + edge)
                seq.append(_one)
            elif inside and s == _zero:
                inside = False
                b = b - edge
                if b > bounds[-1]:
                    bounds.append(b)
                    seq.append(_zero)
                else:
                    bounds[-1] = b
        return Waveform(tuple(bounds), tuple(seq))

    def __mul__(self, other):
        if not isinstance(other, Waveform):
            other = const(other)
        return self._comb(other, mul)

    def __rmul__(self, v):
        return const(v) * self

    def __truediv__(self, other):
        if isinstance(other, Waveform):
            raise TypeError('division by waveform')
        return self * const(1 / other)

    def __neg__(self):
        return -1 * self

    def __sub__(self, other):
        return self + (-other)

    def __rsub__(self, v):
        return v + (-self)

    def __rshift__(self, time):
        return Waveform(tuple(_rnd(b + time) for b in self.bounds),
                        tuple(shift(piece, time) for piece in self.seq))

    def __lshift__(self, time):
        return self >> (-time)

    def __hash__(self):
        return hash((self.max, self.min, self.start, self.stop,
                     self.sample_rate, self.bounds, self.seq))

    def __eq__(self, o):
        if isinstance(o, (int, float, complex)):
            return self == const(o)
        if not isinstance(o, Waveform):
            return False
        a, b = self.simplify(), o.simplify()
        return (a.seq == b.seq and a.bounds == b.bounds
                and (a.max, a.min, a.start, a.stop) == (b.max, b.min, b.start,
                                                         b.stop))


class WaveVStack(Waveform):
    """Lazy sum of many piecewise members (reference waveform.py:638-844).

    On the device one stack is ONE output channel: ``lowering`` merges the
    members' bounds into a single segment table whose segments list the members'
    terms group by group, so a 1000-pulse channel is evaluated in one pass with
    the same accumulation order as the reference's complex128 accumulator."""

    def __init__(self, wlist: list[Waveform] = []):
        self.wlist = [(w.bounds, w.seq) for w in wlist]
        self.start = None
        self.stop = None
        self.sample_rate = None
        self.offset = 0
        self.shift = 0
        self.filters = None
        self.label = None
        self.function_lib = None

    def _extent(self, pick, agg, empty):
        if not self.wlist:
            return empty
        return agg(pick(b, s) for b, s in self.wlist)

    @property
    def begin(self):
        b = self._extent(self._begin, min, -inf)
        return b if self.start is None else max(self.start, b)

    @property
    def end(self):
        e_ = self._extent(self._end, max, inf)
        return e_ if self.stop is None else min(self.stop, e_)

    def _channel(self):
        from .lowering import Channel
        return Channel(members=list(self.wlist), clip=None,
                       offset=self.offset, pre_shift=self.shift,
                       real_only=True)

    def __call__(self, x, frag=False, out=None, function_lib=None):
        assert frag is False, 'WaveVStack does not support frag mode'
        from . import engine
        engine.check_function_lib(
            function_lib if function_lib is not None else self.function_lib)
        if isinstance(x, (int, float, complex)):
            return self.__call__(np.array([x]))[0]
        # like the reference, ``out`` is ignored and a fresh real array returned
        return engine.sample_one(self._channel(), engine.explicit_grid(x))

    def tolist(self):
        flat = [self.start, self.stop, self.offset, self.shift,
                self.sample_rate]
        self._filters_tolist(self.filters, flat)
        flat.append(len(self.wlist))
        for bounds, seq in self.wlist:
            self._tolist(bounds, seq, flat)
        return flat

    @classmethod
    def fromlist(cls, l):
        w = cls()
        w.start, w.stop, w.offset, w.shift, w.sample_rate, sos_size = l[:6]
        filt, pos = cls._filters_fromlist(l, 6, sos_size)
        if filt is not None:
            w.filters = filt
        count = l[pos]
        pos += 1
        for _ in range(count):
            bounds, seq, pos = cls._fromlist(l, pos)
            w.wlist.append((bounds, seq))
        return w

    def simplify(self, eps=1e-15):
        if not self.wlist:
            return zero()
        wav = Waveform(*wave_sum(self.wlist))
        if self.offset != 0:
            wav += self.offset
        if self.shift != 0:
            wav >>= self.shift
        wav = wav.simplify(eps)
        wav.start, wav.stop = self.start, self.stop
        wav.sample_rate = self.sample_rate
        wav.filters = self.filters
        wav.label = self.label
        return wav

    @staticmethod
    def _rshift(wlist, time):
        if time == 0:
            return wlist
        return [(tuple(_rnd(b + time) for b in bounds),
                 tuple(shift(piece, time) for piece in seq))
                for bounds, seq in wlist]

    def _like(self, wlist=None):
        ret = WaveVStack()
        if wlist is not None:
            ret.wlist = wlist
        ret.filters = self.filters
        ret.label = self.label
        return ret

    def __rshift__(self, time):
        ret = self._like(self.wlist)
        ret.sample_rate = self.sample_rate
        ret.start, ret.stop = self.start, self.stop
        ret.shift = self.shift + time
        ret.offset = self.offset
        return ret

    # NOTE (provenance): ``__add__`` / ``__mul__`` / ``__eq__`` / ``__getstate__`` / ``__setstate__`` below TRANSCRIBE the
    # behaviour of the reference's ``WaveVStack`` (waveform.py:771-844) branch for branch: which operand is shifted or
    # simplified first, where the offset goes, what the pickle tuple holds are all observable through the drop-in API
    # (tests/test_host_model.py compares ``tolist()`` of both packages on the reference's own stack tests).  Host glue;
    # no claim of original design.  What is new for stacks lives in ``_channel`` / ``lowering`` / the kernels.
    def __add__(self, other):
        ret = self._like(list(self.wlist))
        if isinstance(other, WaveVStack):
            if other.shift != self.shift:
                ret.wlist = self._rshift(ret.wlist, self.shift)
                ret.wlist.extend(self._rshift(other.wlist, other.shift))
            else:
                ret.wlist.extend(other.wlist)
            ret.offset = self.offset + other.offset
        elif isinstance(other, Waveform):
            other <<= self.shift
            ret.wlist.append((other.bounds, other.seq))
        else:
            ret.offset += other
        return ret

    def __radd__(self, v):
        return self + v

    def __mul__(self, other):
        if isinstance(other, Waveform):
            other = other.simplify() << self.shift
            ret = WaveVStack([Waveform(*w) * other for w in self.wlist])
            if self.offset != 0:
                w = other * self.offset
                ret.wlist.append((w.bounds, w.seq))
        else:
            ret = WaveVStack([Waveform(*w) * other for w in self.wlist])
            ret.offset = self.offset * other
        ret.filters = self.filters
        ret.label = self.label
        return ret

    def __rmul__(self, v):
        return self * v

    def __eq__(self, other):
        if self.wlist:
            return False
        return zero() == other

    __hash__ = None

    def __getstate__(self):
        function_lib = self.function_lib
        if function_lib:
            try:
                import dill
                function_lib = dill.dumps(function_lib)
            except Exception:
                function_lib = None
        return (self.wlist, self.start, self.stop, self.sample_rate,
                self.offset, self.shift, self.filters, self.label,
                function_lib)

    def __setstate__(self, state):
        (self.wlist, self.start, self.stop, self.sample_rate, self.offset,
         self.shift, self.filters, self.label, function_lib) = state
        if function_lib:
            try:
                import dill
                function_lib = dill.loads(function_lib)
            except Exception:
                function_lib = None
        self.function_lib = function_lib


# ---------------------------------------------------------------------------
# builders  (reference waveform.py:886-895, :1078-1484)
# ---------------------------------------------------------------------------
_zero_waveform = Waveform()
_one_waveform = Waveform(seq=(_one, ))


def zero():
    return _zero_waveform


def one():
    return _one_waveform


def const(c):
    return Waveform(seq=(_const(1.0 * c), ))


def D(wav: Waveform, d: int = 1) -> Waveform:
    """d-th symbolic derivative (reference waveform.py:1055-1071)."""
    assert d >= 0 and isinstance(d, int), "d must be a non-negative integer"
    for _ in range(d):
        wav = Waveform(bounds=wav.bounds, seq=tuple(_D(p) for p in wav.seq))
    return wav


def convolve(a, b):
    pass


def sign():
    return Waveform(bounds=(0, +inf), seq=(_const(-1), _one))


def _pulse3(lo, hi, body):
    """zero | body | zero on [lo, hi)."""
    return Waveform(bounds=(lo, hi, +inf), seq=(_zero, body, _zero))


def _half_plus_half(factor):
    """0.5 + 0.5*factor, written out as the reference does for erf/cos edges."""
    return ((((), ()), ((factor, ), (1, ))), (0.5, 0.5))


def step(edge, type='erf'):
    """type: "erf", "cos", "linear" (reference waveform.py:1082-1107)."""
    if edge == 0:
        return Waveform(bounds=(0, +inf), seq=(_zero, _one))
    if type == 'cos':
        rise = add(_half,
                   mul(_half, basic_wave(COS, pi / edge, shift=0.5 * edge)))
        knots = (_rnd(-edge / 2), _rnd(edge / 2), +inf)
    elif type == 'linear':
        rise = add(_half, mul(_const(1 / edge), basic_wave(LINEAR)))
        knots = (_rnd(-edge / 2), _rnd(edge / 2), +inf)
    else:
        rise = _half_plus_half((ERF, edge / 5, 0))
        knots = (-_rnd(edge), _rnd(edge), +inf)
    return Waveform(bounds=knots, seq=(_zero, rise, _one))


def square(width: float, edge: float = 0, type: str = 'erf') -> Waveform:
    if width <= 0:
        return zero()
    if edge == 0:
        return _pulse3(_rnd(-0.5 * width), _rnd(0.5 * width), _one)
    return ((step(edge, type=type) << width / 2) -
            (step(edge, type=type) >> width / 2))


def gaussian(width: float, plateau: float = 0.0, d: int | None = None):
    """width is two times FWHM: std*sqrt(2) = width / (4 sqrt(ln 2))
    (reference waveform.py:1123-1150)."""
    if width <= 0 and plateau <= 0.0:
        return zero()
    std_sq2 = width / 3.3302184446307908

    def body(sh):
        if d is None:
            return basic_wave(GAUSSIAN, std_sq2, shift=sh)
        return basic_wave(D_GAUSSIAN, std_sq2, d, shift=sh)

    if _rnd(0.5 * plateau) <= 0.0:
        return _pulse3(_rnd(-0.75 * width), _rnd(0.75 * width), body(0))
    return Waveform(bounds=(_rnd(-0.75 * width - 0.5 * plateau),
                            _rnd(-0.5 * plateau), _rnd(0.5 * plateau),
                            _rnd(0.75 * width + 0.5 * plateau), +inf),
                    seq=(_zero, body(-0.5 * plateau), _one,
                         body(0.5 * plateau), _zero))


def cos(w: float, phi: float = 0) -> Waveform:
    if w == 0:
        return const(np.cos(phi))
    if w < 0:
        phi, w = -phi, -w
    return Waveform(seq=(basic_wave(COS, w, shift=-phi / w), ))


def sin(w: float, phi: float = 0) -> Waveform:
    if w == 0:
        return const(np.sin(phi))
    if w < 0:
        phi, w = -phi + pi, -w
    return Waveform(seq=(basic_wave(COS, w, shift=(pi / 2 - phi) / w), ))


def exp(alpha: float | complex) -> Waveform:
    if isinstance(alpha, complex):
        osc = cos(alpha.imag) + 1j * sin(alpha.imag)
        return osc if alpha.real == 0 else exp(alpha.real) * osc
    return Waveform(seq=(basic_wave(EXP, alpha), ))


def sinc(bw: float) -> Waveform:
    if bw <= 0:
        return zero()
    width = 100 / bw
    return _pulse3(_rnd(-0.5 * width), _rnd(0.5 * width),
                   basic_wave(SINC, bw))


def cosPulse(width: float, plateau: float = 0.0) -> Waveform:
    if _rnd(0.5 * plateau) > 0:
        return square(plateau + 0.5 * width, edge=0.5 * width, type='cos')
    if width <= 0:
        return zero()
    return _pulse3(_rnd(-0.5 * width), _rnd(0.5 * width),
                   _half_plus_half((COS, 6.283185307179586 / width, 0)))


def hanning(width: float, plateau: float = 0.0) -> Waveform:
    return cosPulse(width, plateau=plateau)


def cosh(w: float) -> Waveform:
    return Waveform(seq=(basic_wave(COSH, w), ))


def sinh(w: float) -> Waveform:
    return Waveform(seq=(basic_wave(SINH, w), ))


def coshPulse(width: float, eps: float = 1.0, plateau: float = 0.0):
    """f(t) = (cosh(eps/2) - cosh(eps t/T)) / (cosh(eps/2) - 1) on [-T/2, T/2],
    optionally split around a plateau (reference waveform.py:1212-1265)."""
    if width <= 0 and plateau <= 0:
        return zero()
    w = eps / width
    A = np.cosh(eps / 2)
    amps = (A / (A - 1), -1 / (A - 1))

    def edge(sh):
        return ((((), ()), (((COSH, w, sh), ), (1, ))), amps)

    if plateau == 0.0 or _rnd(-0.5 * plateau) == _rnd(0.5 * plateau):
        return _pulse3(_rnd(-0.5 * width), _rnd(0.5 * width), edge(0))
    return Waveform(bounds=(_rnd(-0.5 * width - 0.5 * plateau),
                            _rnd(-0.5 * plateau), _rnd(0.5 * plateau),
                            _rnd(0.5 * width + 0.5 * plateau), +inf),
                    seq=(_zero, edge(-0.5 * plateau), _one,
                         edge(0.5 * plateau), _zero))


def general_cosine(duration: float, *arg: float) -> Waveform:
    wav = zero()
    coef = np.asarray(arg)
    coef /= coef[::2].sum()  # in-place like the reference (ints raise there too)
    for k, a in enumerate(coef, start=1):
        wav += a / 2 * (1 - (-1)**k * cos(k * 2 * pi / duration))
    return wav * square(duration)


def slepian(duration: float, *arg: float) -> Waveform:
    return general_cosine(duration, *arg)


def mollifier(width: float, plateau: float = 0.0, d: int = 0) -> Waveform:
    """exp(1/((x/r)^2-1)+1) inside |x|<r, r = width/2; ``d``-th derivative;
    optional plateau (reference waveform.py:1285-1317)."""
    assert d >= 0 and isinstance(d, int), "d must be a non-negative integer"
    assert width > 0, "width must be positive"
    r = width / 2
    if plateau <= 0:
        return _pulse3(-0.5 * width, 0.5 * width, basic_wave(MOLLIFIER, r, d))
    return Waveform(
        bounds=(-0.5 * width - 0.5 * plateau, -0.5 * plateau, 0.5 * plateau,
                0.5 * width + 0.5 * plateau, inf),
        seq=(_zero, basic_wave(MOLLIFIER, r, d, shift=-0.5 * plateau), _one,
             basic_wave(MOLLIFIER, r, d, shift=0.5 * plateau), _zero))


def _poly(*a):
    """a[0] + a[1] t + a[2] t^2 + ...   NB: like the reference
    (waveform.py:1320-1333) the amplitude tuple returned is ``a`` itself, so
    zero coefficients mis-align terms; kept for drop-in parity."""
    terms = []
    if a[0] != 0:
        terms.append(((), ()))
    for n, coef in enumerate(a[1:], start=1):
        if coef != 0:
            terms.append((((LINEAR, 0), ), (n, )))
    return tuple(terms), tuple(a)


def poly(a):
    return Waveform(seq=(_poly(*a), ))


def t():
    # malformed in the reference as well (waveform.py:1343-1344)
    return Waveform(seq=((((LINEAR, 0), ), (1, )), (1, )))


def drag(freq: float, width: float, plateau: float = 0, delta: float = 0,
         block_freq: float | None = None, phase: float = 0,
         t0: float = 0) -> Waveform:
    """reference waveform.py:1347-1379"""
    phase += pi * delta * (width + plateau)
    if plateau <= 0:
        return _pulse3(
            _rnd(t0), _rnd(t0 + width),
            basic_wave(DRAG, t0, freq, width, delta, block_freq, phase))
    w = 2 * pi * (freq + delta)
    carrier = basic_wave(COS, w, shift=(phase + 2 * pi * delta * t0) / w)
    if width <= 0:
        return _pulse3(_rnd(t0), _rnd(t0 + plateau), carrier)
    return Waveform(
        seq=(_zero,
             basic_wave(DRAG, t0, freq, width, delta, block_freq, phase),
             carrier,
             basic_wave(DRAG, t0 + plateau, freq, width, delta, block_freq,
                        phase - 2 * pi * delta * plateau), _zero),
        bounds=(_rnd(t0), _rnd(t0 + width / 2), _rnd(t0 + width / 2 + plateau),
                _rnd(t0 + width + plateau), +inf))


def chirp(f0: float, f1: float, T: float, phi0: float = 0,
          type: str = 'linear') -> Waveform:
    """type: "linear", "exponential", "hyperbolic"
    (reference waveform.py:1382-1421)."""
    if f0 == f1:
        return sin(f0, phi0)
    if T <= 0:
        raise ValueError('T must be positive')
    if type == 'linear':
        body = basic_wave(LINEARCHIRP, f0, f1, T, phi0)
    elif type in ['exp', 'exponential', 'geometric']:
        if f0 == 0:
            raise ValueError('f0 must be non-zero')
        body = basic_wave(EXPONENTIALCHIRP, f0, np.log(f1 / f0) / T, phi0)
    elif type in ['hyperbolic', 'hyp']:
        if f0 * f1 == 0:
            return const(np.sin(phi0))
        body = basic_wave(HYPERBOLICCHIRP, f0, (f0 - f1) / (f1 * T), phi0)
    else:
        raise ValueError(f'unknown type {type}')
    return _pulse3(0, _rnd(T), body)


def interp(x, y) -> Waveform:
    """Piecewise-linear through (x, y) (reference waveform.py:1424-1439)."""
    seq, bounds = [_zero], [x[0]]
    for x1, x2, y1, y2 in zip(x[:-1], x[1:], y[:-1], y[1:]):
        if x2 == x1:
            continue
        seq.append(
            add(mul(_const((y2 - y1) / (x2 - x1)), basic_wave(LINEAR,
                                                              shift=x1)),
                _const(y1)))
        bounds.append(x2)
    bounds.append(inf)
    seq.append(_zero)
    return Waveform(seq=tuple(seq),
                    bounds=tuple(_rnd(b) for b in bounds)).simplify()


def _gate(wav, start, stop):
    if start is not None:
        wav = wav * (step(0) >> start)
    if stop is not None:
        wav = wav * ((1 - step(0)) >> stop)
    return wav


def cut(wav: Waveform, start=None, stop=None, head=None, tail=None, min=None,
        max=None) -> Waveform:
    """reference waveform.py:1442-1467 (the one-sample evaluation used for
    ``head``/``tail`` goes through the CUDA path like every other sample)."""
    offset = 0
    if start is not None and head is not None:
        offset = head - wav(np.array([1.0 * start]))[0]
    elif stop is not None and tail is not None:
        offset = tail - wav(np.array([1.0 * stop]))[0]
    wav = _gate(wav + offset, start, stop)
    if min is not None:
        wav.min = min
    if max is not None:
        wav.max = max
    return wav


def function(fun, *args, start=None, stop=None):
    """Registers ``fun`` and builds the symbolic waveform like the reference
    (waveform.py:1470-1478).  Sampling it raises: arbitrary Python callables
    cannot be lowered to the device IR (see lowering.UnsupportedBasis)."""
    TYPEID = registerBaseFunc(fun)
    return _gate(Waveform(seq=(basic_wave(TYPEID, *args), )), start, stop)


def samplingPoints(start, stop, points):
    return _pulse3(_rnd(start), _rnd(stop),
                   basic_wave(INTERP, start, stop, tuple(points)))


def mixing(I: Waveform, Q: Waveform | None = None, *, phase: float = 0.0,
           freq: float = 0.0, ratioIQ: float = 1.0, phaseDiff: float = 0.0,
           block_freq: float | None = None,
           DRAGScaling: float | None = None) -> tuple[Waveform, Waveform]:
    """SSB or envelope mixing with optional DRAG
    (reference waveform.py:1487-1527).  The result is still symbolic: the
    up-conversion and the DRAG derivative terms are evaluated by the same
    kernel pass as the envelope, no intermediate array exists."""
    if Q is None:
        Q = zero()
    w = 2 * pi * freq
    if freq != 0.0:
        Iout = I * cos(w, -phase) + Q * sin(w, -phase)
        Qout = -I * sin(w, -phase + phaseDiff) + Q * cos(w, -phase + phaseDiff)
    else:
        Iout = I * np.cos(-phase) + Q * np.sin(-phase)
        Qout = -I * np.sin(-phase) + Q * np.cos(-phase)

    if block_freq is not None and block_freq != freq:
        a = block_freq / (block_freq - freq)
        b = 1 / (block_freq - freq)
        Iout, Qout = (a * Iout + b / (2 * pi) * D(Qout),
                      a * Qout - b / (2 * pi) * D(Iout))
    elif DRAGScaling is not None and DRAGScaling != 0:
        Iout, Qout = ((1 - w * DRAGScaling) * Iout - DRAGScaling * D(Qout),
                      (1 - w * DRAGScaling) * Qout + DRAGScaling * D(Iout))
    return Iout, ratioIQ * Qout


__all__ = [
    'D', 'Waveform', 'WaveVStack', 'chirp', 'const', 'cos', 'cosh',
    'coshPulse', 'cosPulse', 'cut', 'drag', 'exp', 'function', 'gaussian',
    'general_cosine', 'hanning', 'interp', 'mixing', 'mollifier', 'one',
    'poly', 'registerBaseFunc', 'registerDerivative', 'samplingPoints', 'sign',
    'sin', 'sinc', 'sinh', 'square', 'step', 't', 'zero'
]
