"""Multi-GPU host logic on CPU: channels are partitioned over ranks
with no data-path collective.  The world-size-2 gloo job below checks that every
rank derives the same partition on its own, that the shards tile the channel list
exactly once, and that lowering a shard yields the same IR rows as the matching
slice of the whole batch.  gloo is used by the TEST to compare ranks; the product
path issues no collective."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from waveforms_b200 import batch as B
from waveforms_b200 import cosPulse, mixing, square, zero

ROOT = Path(__file__).resolve().parent.parent


def make_channels(n=11, seed=7):
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        if k % 3 == 2:
            w = zero()
            for j in range(1 + k % 4):
                w = w + rng.uniform(-0.5, 0.5) * (square(50e-9, edge=2e-9) >> (100e-9 * (j + 1)))
        else:
            w, _ = mixing(rng.uniform(0.2, 1) * cosPulse(20e-9) >> (40e-9 + 10e-9 * k), freq=-20e6 * (1 + k % 8),
                          phase=rng.uniform(0, 6), DRAGScaling=4e-10)
        w.start, w.stop, w.sample_rate = 0.0, 0.3e-6 + 0.1e-6 * (k % 5), 2e9  # ragged lengths
        out.append(w)
    return out


def test_shard_ranges_tile_and_balance():
    rng = np.random.default_rng(0)
    for n_ch, n_sh in [(1, 1), (5, 8), (40, 2), (4096, 8), (100000, 8), (7, 3)]:
        w = rng.integers(1, 50000, n_ch)
        r = B.shard_ranges(w, n_sh)
        assert len(r) == n_sh and r[0][0] == 0 and r[-1][1] == n_ch
        assert all(a[1] == b[0] for a, b in zip(r, r[1:])) and all(lo <= hi for lo, hi in r)
        if n_ch >= 8 * n_sh:
            loads = [w[lo:hi].sum() for lo, hi in r]
            assert max(loads) - min(loads) <= 2 * w.max()
    assert B.shard_ranges([], 4) == [(0, 0)] * 4
    eq = B.shard_ranges([10] * 16, 4)
    assert eq == [(0, 4), (4, 8), (8, 12), (12, 16)]


def test_rank_shard_env_and_errors(monkeypatch):
    monkeypatch.setenv('RANK', '1')
    monkeypatch.setenv('WORLD_SIZE', '2')
    assert B.rank_shard([1, 1, 1, 1]) == (2, 4)
    with pytest.raises(ValueError):
        B.rank_shard([1, 1], rank=2, world=2)


def test_shard_lowering_equals_slice_of_whole():
    chans = make_channels()
    whole = B.lower([B.channel_grid(w) for w in chans])
    seen = 0
    for rank in range(3):
        lo, hi, part = B.lower_rank_shard(chans, rank, 3)
        assert lo == seen
        seen = hi
        assert np.array_equal(part.waves['n'], whole.waves['n'][lo:hi])
        s0 = int(whole.waves['seg_begin'][lo]) if hi > lo else 0
        ns = len(part.seg_bound)
        assert np.array_equal(part.seg_bound, whole.seg_bound[s0:s0 + ns])
        f0, t0 = int(whole.seg_ptr['fac'][s0]), int(whole.seg_ptr['term'][s0])
        assert np.array_equal(part.seg_ptr['fac'], whole.seg_ptr['fac'][s0:s0 + ns + 1] - f0)
        assert np.array_equal(part.terms['amp_re'], whole.terms['amp_re'][t0:t0 + len(part.terms)])
        assert np.array_equal(part.facs['shift'], whole.facs['shift'][f0:f0 + len(part.facs)])
        assert part.waves['out_off'][0] == 0 if hi > lo else True
    assert seen == len(chans)


WORKER = r'''
import os, sys, hashlib
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + '/tests')
import numpy as np
import torch.distributed as dist
from test_sharding import make_channels
from waveforms_b200 import batch as B
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
chans = make_channels()
lo, hi, part = B.lower_rank_shard(chans)           # RANK / WORLD_SIZE from torchrun's env
sig = [hashlib.sha1(part.terms['amp_re'].tobytes() + part.seg_bound.tobytes()).hexdigest(),
       int(part.waves['n'].sum()), int(part.total_samples)]
got = [None] * world
dist.all_gather_object(got, (rank, lo, hi, sig))    # the TEST compares ranks; the product issues no collective
if rank == 0:
    got.sort()
    assert [g[0] for g in got] == list(range(world))
    assert got[0][1] == 0 and got[-1][2] == len(chans)
    assert all(a[2] == b[1] for a, b in zip(got, got[1:]))
    total = sum(g[3][1] for g in got)
    assert total == sum(B.channel_grid(w)[1].n for w in chans), total
    for r, l, h, s in got:                           # every rank derived the partition rank 0 derives for it
        l2, h2, p2 = B.lower_rank_shard(chans, r, world)
        assert (l, h) == (l2, h2)
        assert s[0] == hashlib.sha1(p2.terms['amp_re'].tobytes() + p2.seg_bound.tobytes()).hexdigest()
    print('SHARDING_OK', [(g[1], g[2]) for g in got])
dist.barrier()
dist.destroy_process_group()
'''


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER.format(root=str(ROOT)))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='', OMP_NUM_THREADS='1')
    res = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                          '--master-addr', '127.0.0.1', '--master-port', str(_free_port()), str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert 'SHARDING_OK' in res.stdout


def test_reference_arm_other_ranks_do_no_work():
    """bench.py --impl reference under torchrun: rank 0 alone runs; the others exit 0 silently."""
    env = dict(os.environ, RANK='1', LOCAL_RANK='1', WORLD_SIZE='2', CUDA_VISIBLE_DEVICES='')
    res = subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1'],
                         capture_output=True, text=True, timeout=120, env=env)
    assert res.returncode == 0 and res.stdout.strip() == '', res.stdout + res.stderr
