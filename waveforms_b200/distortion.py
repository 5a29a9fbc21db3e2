"""Flux-line predistortion — drop-in for /root/reference/waveforms/distortion.py.

Filter DESIGN stays on the host (tiny polynomial / root-finding work, SciPy as
in the reference): ``exp_decay_filter``, ``high_pass_filter``,
``combine_filters``, ``factor_filter``, ``stable_filter``, ``reflection_filter``,
``zDistortKernel``.  Filter APPLICATION to sampled signals runs on the GPU:

    predistort / distort  lfilter  -> wfm_lfilter (bit-faithful DF2T, K2)
                          kernel convolution -> wfm_fft_filter (K3)
    reflection / correct_reflection          -> wfm_fft_filter (K3)

Signals may be NumPy arrays (copied to the current CUDA device and back, like
any other call of this package) or CUDA torch tensors of shape (n,) or
(n_sig, n), which are processed as a batch and stay on the device.
"""
from __future__ import annotations

import warnings
from itertools import zip_longest
from typing import Sequence

import math

import numpy as np
from scipy.signal import lfiltic, tf2zpk, zpk2sos, zpk2tf

from . import dsp


def _to_device(sig, copy=True):
    """(tensor, was_numpy).  A CUDA tensor is copied unless ``copy=False`` (the caller
    names its own output buffer, so the input is only read)."""
    import torch
    if isinstance(sig, torch.Tensor):
        t = sig.to(dtype=torch.float64)
        if copy:
            return t.contiguous().clone(), False
        return (t if t.stride(-1) == 1 else t.contiguous()), False
    arr = np.ascontiguousarray(np.asarray(sig, dtype=np.float64))
    return torch.from_numpy(arr).cuda(), True


def _fresh_out(sig, dev):
    """Where a filter writes when the caller named no output: the uploaded copy of a NumPy input is ours to overwrite;
    a CUDA input is only read and the result goes into a new tensor (no defensive copy of the input first)."""
    import torch
    if not isinstance(sig, torch.Tensor) or dev.data_ptr() != sig.data_ptr():
        return dev
    return torch.empty(dev.shape, dtype=dev.dtype, device=dev.device)


def _from_device(t, was_numpy):
    return t.cpu().numpy() if was_numpy else t


# NOTE (provenance): the filter DESIGN helpers below (``shift``, ``extractKernel``, ``zDistortKernel``,
# ``high_pass_filter``, ``exp_decay_filter_old``, ``exp_decay_filter``, ``reflection_filter``, ``combine_filters``,
# ``factor_filter``, ``stable_filter``) restate the reference's closed-form formulas (distortion.py:12-286) with the
# same NumPy / SciPy calls in the same order — their outputs are compared with ``array_equal`` against golden vectors
# of the reference (tests/test_gpu_dsp.py::test_design_functions_golden), so nothing about them is free to differ.
# They are tiny host-side polynomial algebra; this package's own work is the APPLICATION of the filters on the GPU
# (``reflection``, ``correct_reflection``, ``predistort``, ``distort`` -> csrc/wfm_iir.cu, csrc/wfm_fft.cu).
def shift(signal: np.ndarray, delay: float, dt: float) -> np.ndarray:
    """delay a signal (reference distortion.py:12-39; host, three-tap
    interpolation + integer shift — not on the GPU path)."""
    points = int(delay // dt)
    delta = delay / dt - points
    if delta > 0:
        signal = np.convolve(signal, np.array([0, 1 - delta, delta]),
                             mode='same')
    if points == 0:
        return signal
    ret = np.zeros_like(signal)
    if points < 0:
        ret[:points] = signal[-points:]
    else:
        ret[points:] = signal[:-points]
    return ret


def extractKernel(sig_in, sig_out, sample_rate, bw=None, skip=0):
    """Calibration helper (reference distortion.py:42-48), host."""
    from scipy.fftpack import fft, ifft, ifftshift
    corr = fft(sig_in) / fft(sig_out)
    ker = np.real(ifftshift(ifft(corr)))
    if bw is not None and bw < 0.5 * sample_rate:
        k = np.exp(-0.5 * np.linspace(-3.0, 3.0, int(2 * sample_rate / bw))**2)
        ker = np.convolve(ker, k / k.sum(), mode='same')
    return ker[int(skip):len(ker) - int(skip)]


def zDistortKernel(dt: float, params: Sequence[tuple]) -> np.ndarray:
    """reference distortion.py:51-60 (design: a ~2k-point host transform)."""
    from scipy.fftpack import fftfreq, ifft, ifftshift
    t = 3 * np.asarray(params)[:, 0].max()
    omega = 2 * np.pi * fftfreq(int(t / dt) + 1, dt)
    H = 1
    for tau, A in params:
        H += (1j * A * omega * tau) / (1j * omega * tau + 1)
    return ifftshift(ifft(1 / H)).real


def high_pass_filter(tau, sample_rate):
    k = 2.0 * tau * sample_rate
    return [k / (1 + k), -k / (1 + k)], [1.0, (1 - k) / (1 + k)]


def exp_decay_filter_old(amp, tau, sample_rate):
    """reference distortion.py:73-99"""
    alpha = 1 - np.exp(-1 / (abs(sample_rate * tau) * (1 + amp)))
    if amp >= 0:
        k = amp / (1 + amp - alpha)
        a = [(1 - k + k * alpha), -(1 - k) * (1 - alpha)]
    else:
        k = -amp / (1 + amp) / (1 - alpha)
        a = [(1 + k - k * alpha), -(1 + k) * (1 - alpha)]
    b = [1 / a[0], -(1 - alpha) / a[0]]
    return b, [1, a[1] / a[0]]


def exp_decay_filter(amp, tau, sample_rate, inv: bool = False, output='ba'):
    """Multi-exponential step-response filter
    out(t) = u(t) (1 - sum_i A_i exp(-t/tau_i)); reference distortion.py:102-185.
    output: 'ba' | 'sos' | 'zpk'."""
    if isinstance(amp, (int, float, complex)):
        amp, tau = [amp], [tau]
    numerator, denominator = np.poly1d([0.0]), np.poly1d([1.0])
    for i, (A, t) in enumerate(zip(amp, tau)):
        denominator = denominator * np.poly1d([1, -1 / t])
        term = np.poly1d([-A, 0.0])
        for j, t_ in enumerate(tau):
            if j != i:
                term = term * np.poly1d([1, -1 / t_])
        numerator = numerator + term
    numerator = numerator + denominator

    z = np.exp(-numerator.roots / sample_rate)
    p = np.exp(-1 / (np.asarray(tau) * sample_rate))
    if inv:
        z, p = p, z
    p = p[np.abs(p) < 1]  # drop unstable poles
    k = (np.prod(1 - p) / np.prod(1 - z)).real
    if output == 'sos':
        return zpk2sos(z, p, k)
    if output == 'ba':
        return zpk2tf(z, p, k)
    if output == 'zpk':
        return z, p, k
    raise ValueError(f"Invalid output type: {output}")


def reflection_filter(f, A, tau):
    """H(f) = (1 - A) / (1 - A exp(-2 pi i f tau)); reference :188-205."""
    return (1 - A) / (1 - A * np.exp(-2j * np.pi * f * tau))


def reflection(sig, A, tau, sample_rate, out=None):
    """ifft(fft(sig) * H).real on the GPU (reference :208-210).  ``out``: CUDA tensor
    that receives the result when ``sig`` is a CUDA tensor (may be ``sig`` itself)."""
    dev, was_np = _to_device(sig, copy=False)
    res = dsp.reflection_device(dev, A, tau, sample_rate, inverse=False, out=_fresh_out(sig, dev) if out is None else out)
    return _from_device(res, was_np)


def correct_reflection(sig, A, tau, sample_rate=None, out=None):
    """Waveform -> symbolic inverse (sig/(1-A) - A/(1-A) (sig >> tau));
    sampled signal -> ifft(fft(sig) / H).real on the GPU (reference :213-223).
    ``out``: CUDA tensor that receives the result when ``sig`` is a CUDA tensor."""
    from .waveform import Waveform
    if isinstance(sig, Waveform):
        return 1 / (1 - A) * sig - A / (1 - A) * (sig >> tau)
    if sample_rate is None:
        raise ValueError('sample_rate is not given')
    dev, was_np = _to_device(sig, copy=False)
    res = dsp.reflection_device(dev, A, tau, sample_rate, inverse=True, out=_fresh_out(sig, dev) if out is None else out)
    return _from_device(res, was_np)


def combine_filters(filters):
    """Polynomial product of (b, a) pairs (reference :226-244)."""
    b, a = np.poly1d([1.0]), np.poly1d([1.0])
    for b_, a_ in filters:
        b = b * np.poly1d(b_)
        a = a * np.poly1d(a_)
    return b.coeffs, a.coeffs


_COMBINED = {}  # coefficient bytes of a filter list -> (b, a, all poles inside the unit circle)


def _combined(filters):
    """combine_filters + the stability test of predistort (reference :298-303), remembered per filter list: a calibration
    applies the same filters to every batch, and the root finding costs more host time than the device filter takes."""
    key = tuple((np.asarray(b_, dtype=np.float64).tobytes(), np.asarray(a_, dtype=np.float64).tobytes()) for b_, a_ in filters)
    hit = _COMBINED.get(key)
    if hit is None:
        b, a = combine_filters(filters)
        z, p, k = tf2zpk(b, a)
        hit = (b, a, bool(np.all(np.abs(p) < 1)), {})  # the dict: lfiltic states per `initial` level
        if len(_COMBINED) >= 64:
            _COMBINED.pop(next(iter(_COMBINED)))
        _COMBINED[key] = hit
    return hit


def factor_filter(b, a):
    """Split into first-order sections (reference :247-266)."""
    b, a = np.poly1d(b), np.poly1d(a)
    p, q = a.roots, b.roots
    b_amp = (b[0] / a[0])**(1 / max(len(q), len(p)))
    return [([b_amp, -b_amp * b_], [1, -a_])
            for a_, b_ in zip_longest(p, q, fillvalue=0)]


def stable_filter(exp_decay_filters: list, sample_rate: float):
    """reference :269-286 (including its (a, b) unpacking order)."""
    filters = []
    for amp, tau in exp_decay_filters:
        a, b = exp_decay_filter(amp, tau, sample_rate)
        filters.append((b, a))
    b, a = combine_filters(filters)
    z, p, k = tf2zpk(b, a)
    return bool(np.all(np.abs(p) < 1))


def _centered_kernel_response(ker, n):
    """Frequency response of the zero-padded linear convolution the reference
    performs with fftconvolve over a 3n-padded signal and then slices
    [n + K//2 : 2n + K//2] (reference :329-333): on a length-L circular grid
    (L >= n + K - 1) that is the kernel advanced by K//2 samples."""
    K = len(ker)
    L = dsp_next_fast_len(n + K - 1)
    h = np.zeros(L)
    h[:K] = ker
    h = np.roll(h, -(K // 2))
    return np.fft.fft(h), L


_MAX_POINTS = 6144  # csrc/wfm_fft.cu: kMaxPoints (one transform per CTA up to here, two levels above)


def _fft_radices(n):
    """The radix sequence csrc/wfm_fft.cu (factor_smooth) runs a length-n transform with, or None."""
    seq = []
    if n % 10 == 0:
        e, m = 0, n
        while m % 2 == 0:
            m //= 2
            e += 1
        if e % 3 == 1:
            seq.append(10)
            n //= 10
    for r in (7, 5, 3, 8, 4, 2):
        while n % r == 0:
            seq.append(r)
            n //= r
    return seq if n == 1 else None


_STAGE_COST = {2: 1.2, 3: 1.3, 4: 1.0, 5: 1.0, 7: 2.0, 8: 1.0, 10: 1.0}  # per point, relative (measured on cfg4-sized batches)


def _fft_cost(n):
    """Relative cost of the device's frequency-domain filter at transform length n (inf: it would take the generic
    Bluestein path): points x stages, both directions, on the split csrc/wfm_fft.cu (split_two_level) chooses."""
    def stages(length):
        seq = _fft_radices(length)
        if seq is None:
            return math.inf
        cost = sum(_STAGE_COST[r] for r in seq)
        return cost * (0.85 if length in (625, 640) else 1.0)  # the compile-time plans
    if n <= _MAX_POINTS:
        return n * 2 * stages(n)
    best, d = 0, 1
    while d * d <= n:
        if n % d == 0 and n // d <= _MAX_POINTS and _fft_radices(d) is not None and _fft_radices(n // d) is not None:
            best = d
        d += 1
    if not best:
        return math.inf
    return 1.15 * n * 2 * (stages(best) + stages(n // best))  # three launches and a complex scratch round trip


def dsp_next_fast_len(m, slack=0.05):
    """Transform length >= m for a zero-padded (linear) convolution on the device: among the 7-smooth lengths up to
    ``slack`` above the smallest one, the cheapest by ``_fft_cost`` — the smallest 7-smooth length is often a poor one
    (403 200 = 630 x 640 needs a radix-7 and two radix-3 stages; 409 600 = 640 x 640 runs 13 % faster)."""
    def smooth(v):
        for r in (2, 3, 5, 7):
            while v % r == 0:
                v //= r
        return v == 1
    while not smooth(m):
        m += 1
    best, best_cost = m, _fft_cost(m)
    for n in range(m + 1, int(m * (1 + slack)) + 1):
        if smooth(n):
            c = _fft_cost(n)
            if c < best_cost:
                best, best_cost = n, c
    return best


_KERNEL_RESPONSES = {}  # (kernel bytes, n, device) -> dsp.PreparedResponse, least recently used first


def _kernel_response(ker, n, device):
    """The centred-convolution response of ``ker`` for signals of ``n`` samples, prepared
    on ``device`` once: a calibration applies the same kernel to every batch."""
    key = (ker.tobytes(), int(n), int(device))
    resp = _KERNEL_RESPONSES.pop(key, None)
    if resp is None:
        Hk, _ = _centered_kernel_response(ker, n)
        resp = dsp.PreparedResponse(Hk, device)
        while len(_KERNEL_RESPONSES) >= 8:
            _KERNEL_RESPONSES.pop(next(iter(_KERNEL_RESPONSES))).close()
    _KERNEL_RESPONSES[key] = resp
    return resp


def predistort(sig, filters: list | None = None, ker=None, initial: float = 0.0,
               initial_x=None, initial_y=None, zi=None, return_zf: bool = False, iir_mode='exact'):
    """IIR predistortion (lfilter with lfiltic initial state) followed by an
    optional centred kernel convolution; reference :289-337.

    ``iir_mode`` (not in the reference): 'exact' = the sequential kernel, bit-identical to scipy.signal.lfilter;
    'scan' = block-parallel (combined filters up to order 4), 40 x faster on long batches, equal up to the filter's
    rounding-noise gain; 'auto' = exact up to ``dsp.IIR_AUTO_EXACT_MAX`` samples per signal."""
    # the caller's tensor is only read: each stage writes a buffer of this call (the first one a fresh one), so a CUDA
    # input costs no defensive copy (819 MB read + written for cfg4's 256 x 400 000 samples: as long as the IIR itself)
    import torch
    dev, was_np = _to_device(sig, copy=False)
    owned = was_np or not isinstance(sig, torch.Tensor) or dev.data_ptr() != sig.data_ptr()
    zf = None

    def target(cur):
        if owned:
            return cur
        out = torch.empty_like(cur)
        return out if out.stride() == cur.stride() else None

    if filters is not None:
        b, a, stable, zi_of = _combined(filters)
        if not stable:
            warnings.warn('Warning: filter is unstable')
        if zi is None and initial_x is None and initial_y is None:
            # the steady state for a constant level: remembered per filter list (scipy's lfiltic costs more host time
            # than the block-parallel filter takes on the device)
            zi = zi_of.get(float(initial))
            if zi is None:
                zi = lfiltic(b, a, np.full((len(a) - 1, ), initial), np.full((len(b) - 1, ), initial))
                if len(zi_of) < 64:
                    zi_of[float(initial)] = zi
        if zi is None:
            if initial_x is None:
                initial_x = np.full((len(b) - 1, ), initial)
            else:
                initial_x = np.asarray(initial_x)[:len(b) - 1]
            if initial_y is None:
                initial_y = np.full((len(a) - 1, ), initial)
            else:
                initial_y = np.asarray(initial_y)[:len(a) - 1]
            zi = lfiltic(b, a, initial_y, initial_x)
        out = target(dev)
        if out is None:  # a view whose layout empty_like does not reproduce: filter a private copy in place
            dev = dev.contiguous().clone()
            out = dev
        # (the final state costs a device->host copy and a synchronisation: only when it is asked for)
        dev, zf = dsp.lfilter_device(b, a, dev, zi=zi, want_zf=return_zf, mode=iir_mode, out=out)
        owned = True
        if zf is not None and dev.dim() == 1:
            zf = zf[0]
    if ker is not None:
        ker = np.ascontiguousarray(np.asarray(ker, dtype=np.float64))
        out = target(dev)
        if out is None:
            dev = dev.contiguous().clone()
            out = dev
        # the zero padding of the linear convolution exists only inside the transform
        dev = _kernel_response(ker, dev.shape[-1], dev.device.index).apply(dev, out=out)
        owned = True
    if not owned:
        dev = dev.clone()  # nothing ran: the reference still returns a new array
    out = _from_device(dev, was_np)
    return (out, zf) if return_zf else out


def distort(points, params, sample_rate, initial=0.0, iir_mode='exact'):
    """reference :340-346 (``iir_mode``: see ``predistort``)"""
    filters = []
    for amp, tau in np.asarray(params).reshape(-1, 2):
        b, a = exp_decay_filter(amp, abs(tau), sample_rate)
        filters.append((b, a))
    return predistort(points, filters, initial=initial, iir_mode=iir_mode)


def phase_curve(t, params, df_dphi, pulse_width, start, wav, sample_rate):
    """Calibration fitting helper (reference :349-366): samples ``wav`` and
    distorts it on the GPU, the boxcar integration and interpolation stay host."""
    lim = max(np.max(np.abs(t)), 20e-6)
    num = round(2 * lim * sample_rate)
    tlist = np.arange(num) / sample_rate - lim
    points = wav(tlist)
    pulse_points = round(pulse_width * sample_rate)
    start_points = round((start + pulse_width) * sample_rate) - 1
    ker = np.hstack([np.ones(pulse_points) / sample_rate,
                     np.zeros(start_points)])
    points = np.convolve(2 * np.pi * df_dphi *
                         distort(points, params, sample_rate), ker, mode='same')
    return np.interp(t, tlist, points)
