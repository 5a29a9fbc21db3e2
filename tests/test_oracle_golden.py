"""Pin the oracle: every golden vector produced by the unmodified reference
(tests/golden/make_golden.py) and the known-answer vectors of the reference's own
tests must be reproduced by the CPU restatement in oracle/wfm_oracle.py.

The goldens were generated on the build container; NumPy's SIMD dispatch may pick
other transcendental kernels on another host, so the comparison allows 1e-13
relative (bit-exact on the generating host)."""
import numpy as np
import pytest

from helpers import load_golden, oracle_eval, rel_err

CASE_NAMES = sorted(load_golden())


@pytest.mark.parametrize('name', CASE_NAMES)
def test_oracle_reproduces_reference(name, golden):
    rec = golden[name]
    got = oracle_eval(rec)
    assert got.dtype == rec['expect'].dtype
    assert rel_err(got, rec['expect']) <= 1e-13


@pytest.mark.parametrize('name', [n for n in CASE_NAMES if not n.startswith('cfg5')])
def test_compiled_reference_evaluator_agrees(name, golden):
    """When oracle/_ref holds the reference's own compiled calc_parts
    (oracle/build_ref.py), it must agree with the restatement bit for bit."""
    from oracle.build_ref import load
    ref = load()
    if ref is None:
        pytest.skip('oracle/_ref not built')

    def calc(bounds, seq, x, lo=-np.inf, hi=np.inf):
        return ref.calc_parts(bounds, seq, x, ref._baseFunc, lo, hi)

    rec = golden[name]
    assert np.array_equal(oracle_eval(rec, calc=calc), oracle_eval(rec))


def test_reference_known_answers():
    """Closed forms asserted by /root/reference/tests/test_waveform.py:8-35, 116-138."""
    from oracle import wfm_oracle as O
    t = np.linspace(-10, 10, 1001)
    cos1 = ((((( O.COS, 1, 0.0), ), (1, )), ), (1.0, ))
    y = O.waveform_call((np.inf, ), (cos1, ), t)
    assert np.allclose(y, np.cos(t), atol=1e-4)
    x = O.sample_grid(-10, 10.02, 50)
    assert len(x) == 1001 and np.allclose(x, t)
    tt = np.linspace(0, 10, 1000, endpoint=False)
    lin = (((((O.LINEARCHIRP, 1, 2, 10, 4, 0), ), (1, )), ), (1.0, ))
    y = O.waveform_call((0, 10, np.inf), (O.ZERO, lin, O.ZERO), tt)
    assert np.allclose(y, np.sin(4 + 2 * np.pi * ((2 - 1) / (2 * 10) * tt**2 + tt)))


def test_half_open_segments():
    """square(2): x=-1 -> 1, x=+1 -> 0; zero segments are not clipped."""
    from oracle import wfm_oracle as O
    one = ((((), ()), ), (1.0, ))
    bounds, seq = (-1.0, 1.0, np.inf), (O.ZERO, one, O.ZERO)
    y = O.waveform_call(bounds, seq, np.array([-1.0, 0.0, 1.0]))
    assert y.tolist() == [1.0, 1.0, 0.0]
    y = O.waveform_call(bounds, seq, np.array([-2.0, 0.0]), lo=0.25, hi=0.5)
    assert y.tolist() == [0.0, 0.5]


def test_dsp_oracle_reproduces_reference(dsp_golden):
    from oracle import wfm_oracle as O
    from scipy.signal import sosfilt
    g = dsp_golden
    sig, fs = g['sig'], g['fs']
    tol = 1e-13
    assert rel_err(sosfilt(g['exp_decay_sos'], sig), g['sosfilt']) <= tol
    assert rel_err(O.reflection(sig, 0.05, 13.3e-9, fs), g['reflection']) <= tol
    assert rel_err(O.correct_reflection(sig, 0.05, 13.3e-9, fs), g['correct_reflection']) <= tol
    assert rel_err(O.predistort(sig, ker=g['zDistortKernel']), g['predistort_ker']) <= tol
    for m in (997, 1000, 1215, 2048, 3125):
        s, want = g[f'correct_reflection_{m}']
        assert rel_err(O.correct_reflection(s, 0.07, 11.1e-9, fs), want) <= tol
