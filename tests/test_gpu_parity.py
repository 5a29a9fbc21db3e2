"""GPU parity: every golden case sampled by the CUDA path (through the C-ABI)
against (a) the reference's committed output and (b) the oracle on the same
wire-format program.  fp64 tolerance 1e-12 relative, fp32 1e-6."""
import numpy as np
import pytest

from helpers import FP32_TOL, FP64_TOL, b200_eval, b200_object, load_golden, oracle_eval, rel_err

pytestmark = pytest.mark.gpu

CASE_NAMES = sorted(load_golden())


@pytest.mark.parametrize('name', CASE_NAMES)
def test_golden_fp64(name, golden):
    rec = golden[name]
    got = b200_eval(rec)
    want = rec['expect']
    assert got.dtype == want.dtype
    assert got.shape == want.shape  # sample count equals the reference's
    assert rel_err(got, want) <= FP64_TOL, name
    assert rel_err(got, oracle_eval(rec)) <= FP64_TOL, name


@pytest.mark.parametrize('name', [n for n in CASE_NAMES if not n.startswith(('complex', 'filters'))])
def test_golden_fp32(name, golden):
    from waveforms_b200 import sample_batch
    rec = golden[name]
    if rec['grid'][0] != 'sample':
        pytest.skip('fp32 output is offered on the batched sample() path')
    obj = b200_object(rec)
    got = sample_batch([obj], dtype=np.float32).numpy()[0]
    assert got.dtype == np.float32
    assert rel_err(got.astype(np.float64), rec['expect']) <= FP32_TOL


def test_segment_indexing_exact(golden):
    """Half-open [lo, hi) ownership on abscissae that hit bounds exactly
    (SURVEY §8a-4): values must be bit-identical here (constants only)."""
    rec = golden['boundary_hits']
    assert np.array_equal(b200_eval(rec), rec['expect'])


def test_scalar_call(ns):
    w = ns.square(2.0)
    assert w(-1.0) == 1.0 and w(1.0) == 0.0 and isinstance(w(0.5), np.float64)


def test_frag_and_out(ns, golden):
    x = np.linspace(-3, 3, 601)
    w = ns.gaussian(2.0) + (ns.cosPulse(1.0) >> 1.5)
    full = w(x)
    parts = w(x, frag=True)
    rebuilt = np.zeros_like(x)
    for a, b, part in parts:
        rebuilt[a:b] += part
    assert np.array_equal(rebuilt, full)
    out = np.ones_like(x)
    assert w(x, out=out) is out and np.array_equal(out, full)
    acc = np.ones_like(x)
    w(x, out=acc, accumulate=True)
    assert np.array_equal(acc, full + 1)


def test_batch_matches_single(ns):
    from waveforms_b200 import sample_batch
    rng = np.random.default_rng(5)
    ws = []
    for k in range(7):
        w = rng.uniform(0.2, 1) * ns.cosPulse(30e-9) >> (40e-9 + 37e-9 * k)
        I, Q = ns.mixing(w, freq=rng.uniform(-100e6, 100e6), phase=rng.uniform(0, 6), DRAGScaling=3e-10)
        for ch in (I, Q):
            ch.start, ch.stop, ch.sample_rate = 0, (0.4e-6 + 13e-9 * k), 2e9  # ragged lengths
            ws.append(ch)
    res = sample_batch(ws).numpy()
    for w, y in zip(ws, res):
        assert np.array_equal(y, w.sample())


def test_unsupported_basis_raises(ns):
    from waveforms_b200.lowering import UnsupportedBasis
    w = ns.function(lambda t: t**2)
    with pytest.raises(UnsupportedBasis):
        w(np.linspace(0, 1, 5))
