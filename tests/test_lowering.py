"""Lowering invariants (host logic, no GPU): segment-table structure, factor
de-duplication, stack merging, grids."""
import numpy as np
import pytest

from waveforms_b200 import engine, lowering as L
from waveforms_b200 import WaveVStack, cosPulse, gaussian, mixing, square, zero, samplingPoints


def lower_one(w, grid):
    return L.lower([(w._channel(), grid)])


def test_readme_channel_structure():
    I, _ = mixing(0.5 * cosPulse(20e-9), freq=-20e6, DRAGScaling=0.2)
    b = lower_one(I, engine.arange_grid(-1e-6, 9e-6, 1e-9))
    assert b.waves['n'][0] == 10000 and b.waves['n_seg'][0] == 3
    assert b.seg_bound.tolist() == [-1e-08, 1e-08, np.inf]
    # active segment: 5 terms whose 8 factor references share 5 DISTINCT cos factors (the
    # reference memoises per factor tuple).  Two frequencies -> 2 COS_SINCOS rows (+ their
    # NOP sine rows) and 3 COS_ROT rows.
    assert b.seg_ptr['term'].tolist() == [0, 0, 5, 5]
    assert b.seg_ptr["fac"].tolist() == [0, 0, 7, 7] and len(b.refs) == 8
    assert b.facs['func'].tolist() == [L.COS_SINCOS, L.NOP, L.COS_SINCOS, L.NOP, L.COS_ROT, L.COS_ROT, L.COS_ROT]
    for row in b.facs[4:]:
        base = int(b.args[row['arg_off']])
        assert b.facs[base]['func'] == L.COS_SINCOS and b.facs[base]['a0'] == row['a0']
        D = b.args[row['arg_off'] + 2]
        assert D == row['a0'] * (b.args[row['arg_off'] + 1] - row['shift'])
        assert b.args[row['arg_off'] + 3] == np.cos(D) and b.args[row['arg_off'] + 4] == np.sin(D)
    assert not set(b.refs['slot'].tolist()) & {1, 3}          # nothing refers to a sine slot
    assert b.terms['flags'].tolist() == [0, 0, 0, 0, L.TERM_GROUP_END]
    assert not b.any_complex


def test_arange_grid_matches_numpy():
    for a, z, s in [(-1e-6, 9e-6, 1e-9), (0, 100e-6, 0.5e-9), (-10, 10.02, 1 / 50), (0, 4e-6, 1 / 5e9), (0, 1.0, 0.3)]:
        g = engine.arange_grid(a, z, s)
        assert np.array_equal(g.materialize(), np.arange(a, z, s))
    assert engine.arange_grid(1.0, 1.0, 0.1).n == 0


def test_linspace_grid_matches_numpy():
    for a, z, n, ep in [(-1e-6, 9e-6, 10001, True), (0., 1e-5, 20000, False), (-10, 10, 1001, True), (0, 1, 1, True)]:
        g = engine.linspace_grid(a, z, n, endpoint=ep)
        assert np.array_equal(g.materialize(), np.linspace(a, z, n, endpoint=ep))


def test_stack_merge_keeps_member_order_and_groups():
    a = square(2.0)            # [-1, 1)
    b = 0.5 * (square(2.0) >> 1)  # [0, 2)
    s = WaveVStack([a, b])
    batch = lower_one(s, engine.explicit_grid(np.linspace(-2, 3, 11)))
    assert batch.seg_bound.tolist() == [-1.0, 0.0, 1.0, 2.0, np.inf]
    assert np.diff(batch.seg_ptr['term']).tolist() == [0, 1, 2, 1, 0]
    # overlap segment [0,1): member a's term then member b's, each closing its own group
    t = batch.terms[1:3]
    assert t['amp_re'].tolist() == [1.0, 0.5] and (t['flags'] == L.TERM_GROUP_END).all()


def test_out_offsets_are_aligned_and_ragged():
    ws = []
    for n in (5, 6, 1, 9):
        w = gaussian(1.0)
        ws.append((w._channel(), engine.linspace_grid(-1, 1, n)))
    b = L.lower(ws)
    assert b.waves['out_off'].tolist() == [0, 8, 16, 20] and b.total_samples == 32


def test_interp_table_goes_to_arg_pool():
    pts = tuple(np.linspace(0, 1, 7))
    b = lower_one(samplingPoints(0.0, 3.0, pts), engine.linspace_grid(0, 3, 10))
    f = b.facs[0]
    assert f['func'] == 7 and f['a0'] == 0.0 and f['a1'] == 3.0
    assert b.args[f['arg_off']] == 7 and b.args[f['arg_off'] + 1] == 0.5
    assert b.args[f['arg_off'] + 2:f['arg_off'] + 9].tolist() == list(pts)


def test_user_function_cannot_be_lowered():
    from waveforms_b200 import function
    w = function(lambda t: t)
    with pytest.raises(L.UnsupportedBasis):
        lower_one(w, engine.linspace_grid(0, 1, 4))


def test_shard_ranges_balance():
    from waveforms_b200.batch import shard_ranges
    r = shard_ranges([10] * 16, 4)
    assert r == [(0, 4), (4, 8), (8, 12), (12, 16)]
    r = shard_ranges([100, 1, 1, 1, 1, 100], 2)
    assert r[0][0] == 0 and r[-1][1] == 6 and all(a <= b for a, b in r)
    assert shard_ranges([5, 5], 4)[-1][1] == 2
