"""Host object model: the reference's own unit tests for everything that does not
sample (/root/reference/tests/test_waveform.py:38-65, 141-166;
tests/test_wavevstack.py:28-43, 91-110, 140-143), plus structural identity with
the unmodified reference on every golden case (same ``tolist()``)."""
import numpy as np
import pytest

import cases
from helpers import load_golden

from waveforms_b200 import *  # noqa: F401,F403
from waveforms_b200 import Waveform, WaveVStack, wave_eval
from waveforms_b200._waveform import wave_sum

CASE_NAMES = sorted(load_golden())


@pytest.mark.parametrize('name', CASE_NAMES)
def test_same_structure_as_reference(name, golden, ns):
    obj, grid = cases.CASES[name](ns)
    assert obj.tolist() == golden[name]['flat']


def _pulse():
    pulse = gaussian(10) >> 5
    pulse += gaussian(10) >> 50
    return pulse * cos(200)


def test_tolist_golden():
    pulse = _pulse()
    l = pulse.tolist()
    assert l == [
        np.inf, -np.inf, None, None, None, None, 5, -2.5, 0, 12.5, 1, 1.0, 2,
        1, 3, 2, 3.0028060219661246, 5, 1, 3, 4, 200, 0.0, 42.5, 0, 57.5, 1,
        1.0, 2, 1, 3, 2, 3.0028060219661246, 50, 1, 3, 4, 200, 0.0, np.inf, 0
    ]
    assert Waveform.fromlist(l) == pulse


def test_totree_golden():
    pulse = _pulse()
    t = pulse.totree()
    assert t == ((np.inf, -np.inf, None, None, None, None),
                 ((-2.5, ()), (12.5, ((1.0, ((1, (2, 3.0028060219661246, 5)),
                                             (1, (4, 200, 0.0)))), )),
                  (42.5, ()), (57.5, ((1.0, ((1, (2, 3.0028060219661246, 50)),
                                             (1, (4, 200, 0.0)))), )), (np.inf,
                                                                        ())))
    assert Waveform.fromtree(t) == pulse


def test_parser():
    assert wave_eval("one()") == one()
    assert wave_eval("zero()") == zero()
    assert wave_eval("pi") == pi
    assert wave_eval("e") == e

    w1 = (gaussian(10) << 100) + square(20, edge=5, type='linear') * cos(2 * pi * 23.1)
    w2 = wave_eval("(gaussian(10) << 100) + square(20, edge=5, type='linear') * cos(2*pi*23.1)")
    w3 = wave_eval("((gaussian(10) << 50) + ((square(20, 5, type='linear') * cos(2*pi*23.1)) >> 50)) << 50")
    w4 = wave_eval("(gaussian(10) << 100) + square(20, 5, 'linear') * cos(2*pi*23.1)")
    assert w1 == w2 and w1 == w3 and w1 == w4

    w1 = poly([1, -1 / 2, 1 / 6, -1 / 12])
    assert w1 == wave_eval("poly([1, -1/2, 1/6, -1/12])")
    assert w1 == wave_eval("poly((1, -1/2, 1/6, -1/12))")


def test_parser_antlr_precedence():
    """Waveform.g4:8-22 as ANTLR4 parses it (see waveform_parser.py docstring)."""
    assert wave_eval("2**3**2") == const(64)          # left associative
    assert wave_eval("2^3") == const(8)
    assert wave_eval("-1 + 3") == const(-4)           # unary minus binds loosest
    assert wave_eval("2 * -1 + 3") == const(-8)
    assert wave_eval("1 + 2 * 3") == const(7)
    assert wave_eval("gaussian(10) >> 2 + 3") == (gaussian(10) >> 5)
    assert wave_eval("cos(2*pi*1e6) * 1j") == cos(2 * pi * 1e6) * 1j
    for bad in ("foo", "a = 3", "gaussian(10", "1 +", "nosuch(3)", "gaussian(10) 3", "$"):
        with pytest.raises(SyntaxError):
            wave_eval(bad)


def test_wavevstack_tolist_golden():
    wlist = [cos(1), sin(2), gaussian(3), poly([1, -1 / 2, 1 / 6, -1 / 12])]
    w = WaveVStack(wlist)
    l = w.tolist()
    assert l == [
        None, None, 0, 0, None, None, 4, 1, np.inf, 1, 1.0, 1, 1, 3, 4, 1, 0.0,
        1, np.inf, 1, 1.0, 1, 1, 3, 4, 2, 0.7853981633974483, 3, -2.25, 0,
        2.25, 1, 1.0, 1, 1, 3, 2, 0.9008418065898374, 0, np.inf, 0, 1, np.inf,
        4, 1, 0, -0.5, 1, 1, 2, 1, 0, 0.16666666666666666, 1, 2, 2, 1, 0,
        -0.08333333333333333, 1, 3, 2, 1, 0
    ]
    w2 = WaveVStack.fromlist(l)
    assert isinstance(w2, WaveVStack) and w2.wlist == w.wlist


def test_wavevstack_simplify_equals_sum():
    wlist = [cos(1), sin(2), gaussian(3), poly([1, -1 / 2, 1 / 6, -1 / 12])]
    w1 = zero()
    for w in wlist:
        w1 += w
    assert WaveVStack(wlist).simplify() == w1

    w1, w2 = zero(), []
    assert w1 == WaveVStack(w2).simplify()
    for freq in np.linspace(6.1, 6.5, 11) * 1e9:
        pulse = square(1e-6) >> 95e-6
        w1 += pulse * cos(2 * pi * freq)
        w2.append(pulse * cos(2 * pi * freq))
        assert w1 == WaveVStack(w2).simplify()
    rng = np.random.default_rng(0)
    for freq in np.linspace(6.1, 6.5, 3) * 1e9:
        pulse = square(1e-6) >> (95e-6 + rng.standard_normal() * 1e-9)
        w1 += pulse * cos(2 * pi * freq)
        w2.append(pulse * cos(2 * pi * freq))
        assert w1 == WaveVStack(w2).simplify()
    w1 += cos(2 * pi * freq * 0.9)
    w2.append(cos(2 * pi * freq * 0.9))
    assert w1 == WaveVStack(w2).simplify()


def test_wave_sum_cancels():
    assert wave_sum([((-1.0, np.inf), (((), ()), ((((), ()), ), (0.02, )))),
                     ((-1.0, np.inf), (((), ()), ((((), ()), ), (-0.02, ))))
                     ]) == ((np.inf, ), (((), ()), ))


def test_operator_semantics():
    with pytest.raises(TypeError):
        cos(1) / cos(2)
    w = square(1e-9) >> 1 / 3
    assert w.bounds == (0.333333332833333, 0.333333333833333, np.inf)  # round(., 15) decimals
    assert (cos(1) * cos(1)).seq[0][0][0][1] == (2, )                 # equal factors add exponents
    assert D(cosPulse(1.0)).seq[1] == (((((4, 6.283185307179586, -0.25), ), (1, )), ), (3.141592653589793, ))
    with pytest.raises(ValueError):
        gaussian(1).sample()
    m = (gaussian(2) >> 3).marker
    assert m.seq == (((), ()), ((((), ()), ), (1.0, )), ((), ()))


def test_roundtrips_keep_filters():
    from scipy.signal import butter, tf2sos
    w = step(0)
    w.start, w.stop, w.sample_rate = -1, 1, 1000
    w.filters = (tf2sos(*butter(3, 4.0, 'lowpass', fs=1000)), 0)
    w2 = Waveform.fromlist(w.tolist())
    assert np.array_equal(w2.filters[0], w.filters[0]) and w2.filters[1] == 0
    w3 = Waveform.fromtree(w.totree())
    assert w3.bounds == w.bounds and w3.seq == w.seq and w3.sample_rate == 1000


def test_padded_fft_length_is_chosen_by_cost():
    """distortion.dsp_next_fast_len: a 7-smooth length >= m within 5 % of the smallest one, the cheapest for the device's
    plans (radix sequence and two-level split mirrored from csrc/wfm_fft.cu)."""
    from waveforms_b200 import distortion as D

    def smooth(v):
        for r in (2, 3, 5, 7):
            while v % r == 0:
                v //= r
        return v == 1
    for m in (1, 2, 97, 1000, 4097, 6145, 20000, 123457, 401799, 1000003):
        n = D.dsp_next_fast_len(m)
        least = m
        while not smooth(least):
            least += 1
        assert n >= m and smooth(n) and n <= least * 1.05 + 1
        assert D._fft_cost(n) <= D._fft_cost(least)
    # cfg4's kernel convolution: 400 000 samples + 1800 taps
    assert D.dsp_next_fast_len(400000 + 1800 - 1) == 409600          # 640 x 640, not 403 200 = 630 x 640
    assert D._fft_radices(640) == [10, 8, 8] and D._fft_radices(625) == [5, 5, 5, 5] and D._fft_radices(630) == [10, 7, 3, 3]
    assert D._fft_radices(11) is None and D._fft_cost(2 ** 26) == float('inf')  # no split with both factors <= 6144
