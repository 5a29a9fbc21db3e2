"""Batched sampling: many channels, one kernel launch per GPU.

``sample_batch`` is what an upstream scheduler calls instead of looping
``Waveform.sample()`` over channels (the reference has no batched entry point;
its closest container is ``WaveVStack``, waveform.py:638).  Channels are
independent, so multi-GPU execution is a partition of the channel list — one
program and one output buffer per device, no collective.
"""
from __future__ import annotations

import numpy as np

from . import engine
from .lowering import can_pair, find_pairs, lower


PAIR_MIN_SAMPLES = 4_000_000  # 'auto' pairing of I/Q channels starts at this batch size (per device)


def channel_grid(w, sample_rate=None):
    """(Channel, Grid) of a Waveform/WaveVStack exactly as ``sample()`` would
    evaluate it (waveform.py:173-190)."""
    rate = w.sample_rate if sample_rate is None else sample_rate
    if w.start is None or w.stop is None or rate is None:
        raise ValueError(
            f'Waveform is not initialized. {w.start=}, {w.stop=}, sample_rate={rate}'
        )
    return w._channel(), engine.arange_grid(w.start, w.stop, 1 / rate)


def plan_pairs(items, pair_iq='auto'):
    """The items ``lower`` gets for one device's channels: (Channel, Grid), or
    ((Channel, Channel), Grid) for two channels evaluated as one I/Q pair."""
    items = list(items)
    if pair_iq == 'auto':
        if sum(g.n for _, g in items) >= PAIR_MIN_SAMPLES:
            return find_pairs(items)
        return items
    if pair_iq is True:
        if len(items) % 2:
            raise ValueError('pair_iq=True needs an even number of channels')
        for a, b in zip(items[0::2], items[1::2]):
            if not can_pair(a, b, min_shared=0.0):
                raise ValueError('pair_iq=True: channels differ in grid, clip or are complex')
        return [((a[0], b[0]), a[1]) for a, b in zip(items[0::2], items[1::2])]
    return items


def shard_ranges(weights, n_shards):
    """Contiguous ranges of channels with balanced total weight (samples)."""
    weights = np.asarray(weights, dtype=np.float64)
    csum = np.concatenate([[0.0], np.cumsum(weights)])
    total = csum[-1]
    cuts = [0]
    for s in range(1, n_shards):
        cuts.append(int(np.searchsorted(csum, total * s / n_shards)))
    cuts.append(len(weights))
    cuts = np.maximum.accumulate(np.clip(cuts, 0, len(weights)))
    return [(int(cuts[i]), int(cuts[i + 1])) for i in range(n_shards)]


class BatchResult:
    """Device-resident result of ``sample_batch``: one flat tensor per device
    plus the per-channel (device, offset, length) table."""

    def __init__(self, tensors, table, dtype):
        self.tensors = tensors
        self.table = table
        self.dtype = dtype

    def __len__(self):
        return len(self.table)

    def channel(self, i):
        dev, off, n = self.table[i]
        return self.tensors[dev][off:off + n]

    def numpy(self):
        host = [t.cpu().numpy() for t in self.tensors]
        return [host[dev][off:off + n] for dev, off, n in self.table]


def sample_batch(waveforms, sample_rate=None, dtype=np.float64, devices=None,
                 filters='own', iir_mode='auto', pair_iq='auto', fast_fp32=False):
    """Sample every waveform in ``waveforms`` on its own start/stop/sample_rate
    grid.  ``dtype``: np.float64 (reference parity, 1e-12) or np.float32
    (fp32 output, 1e-6).  ``devices``: list of CUDA device indices to shard the
    channels over (default: the current device).  Returns a ``BatchResult``
    whose tensors stay on the GPUs.

    ``filters='own'`` applies each waveform's ``.filters`` (sample-time IIR,
    waveform.py:193-203) on the device; ``None`` skips them.  ``iir_mode``:
    'exact' (bit-identical to scipy, sequential in time) | 'scan' (block-parallel,
    equal up to the filter's rounding-noise gain) | 'auto' (exact up to 32 768
    samples per channel, scan above; the default here) | None (``dsp.IIR_MODE``).

    ``fast_fp32`` (float32 output only): evaluate with the fp32 evaluator
    (``WFM_F32_FAST``: ~10 % faster; its error grows with the cancellation between a
    segment's terms, see include/wfm_b200.h) instead of fp64 arithmetic rounded at
    the store, which meets 1e-6 on every program.

    (Batches below ``PAIR_MIN_SAMPLES`` samples stay unpaired under 'auto': a pair is one
    work item per tile instead of two, and a launch that cannot fill the GPU is bound by
    the latency of its longest item — the README example is 1.8 x slower paired.)

    ``pair_iq``: 'auto' evaluates ADJACENT channels that share their grid and most
    of their basis functions — the I and Q of one ``mixing()`` call
    (waveform.py:1487-1527) — as one I/Q pair: every cos / envelope factor is
    computed once and feeds both outputs (the second row's cosines are rotated from
    the shared sincos: equal to unpaired sampling to a few ulp); ``False`` keeps every channel
    on its own; ``True`` requires channels (0, 1), (2, 3) ... to pair."""
    import torch
    engine.require_gpu()
    items = [channel_grid(w, sample_rate) for w in waveforms]
    if devices is None:
        devices = [torch.cuda.current_device()]
    np_dtype = np.dtype(dtype)
    code = {np.dtype(np.float64): engine.WFM_F64,
            np.dtype(np.float32): engine.WFM_F32,
            np.dtype(np.complex128): engine.WFM_C128}[np_dtype]
    filtered = filters == 'own' and any(w.filters is not None for w in waveforms)
    if filtered and code == engine.WFM_C128:
        raise TypeError('sample_batch(filters=\'own\') filters real channels; sample complex '
                        'channels one by one (Waveform.sample) or pass filters=None')
    # the IIR runs on float64 samples: an fp32 batch with filters is sampled and filtered in
    # float64 and cast at the end
    run_code = engine.WFM_F64 if (filtered and code == engine.WFM_F32) else code
    if run_code == engine.WFM_F32 and fast_fp32:
        run_code = engine.WFM_F32_FAST
    ranges = shard_ranges([g.n for _, g in items], len(devices))
    if np_dtype == np.dtype(np.complex128):
        pair_iq = False
    if pair_iq is True:
        # an odd cut would split a pair: move it to the next even channel
        ranges = [(lo + (lo & 1) if lo < len(items) else lo, hi + (hi & 1) if hi < len(items) else hi)
                  for lo, hi in ranges]

    shards = [(dev, lo, hi, lower(plan_pairs(items[lo:hi], pair_iq))) for dev, (lo, hi) in zip(devices, ranges)]

    def run(shard):
        # one host thread per device: upload, pre-pass, K1 and the filters of every device run
        # concurrently; nothing here waits for another device
        dev, lo, hi, batch = shard
        with torch.cuda.device(dev):
            prog = engine.Program(batch, dev)
            out = prog.sample_device(dtype=run_code)
            if filtered:
                from .dsp import apply_channel_filters
                apply_channel_filters(out, batch, waveforms[lo:hi], mode=iir_mode)
            if run_code == engine.WFM_F64 and code == engine.WFM_F32:
                out = out.to(torch.float32)
        return prog, out

    results = _run_shards(run, shards)
    tensors, table = [], []
    for slot, ((dev, lo, hi, batch), (prog, out)) in enumerate(zip(shards, results)):
        tensors.append(out)
        for k in range(batch.n_channels):
            table.append((slot, int(batch.chan_off[k]), int(batch.chan_n[k])))
    for prog, _ in results:  # destroy waits for the program's last launch: only after ALL devices were launched
        prog.close()
    return BatchResult(tensors, table, dtype)


def _run_shards(fn, shards):
    """fn(shard) for every shard: inline for one device, one host thread per device
    otherwise (ctypes releases the GIL inside the library, so uploads, pre-passes and
    launches of different devices overlap)."""
    if len(shards) <= 1:
        return [fn(s) for s in shards]
    import concurrent.futures as cf
    with cf.ThreadPoolExecutor(max_workers=len(shards)) as pool:
        return list(pool.map(fn, shards))


def from_wire(record, kind=None):
    """A waveform shipped WITHOUT Python objects: the reference's
    flat list (``Waveform.tolist``, waveform.py:259-276; ``WaveVStack.tolist``,
    :823-835 — the two layouts cannot be told apart, so ``kind='stack'`` selects
    the latter) or its ``(header, body)`` tree (``totree``, :342-353).  Objects pass
    through unchanged."""
    from .waveform import Waveform, WaveVStack
    if isinstance(record, Waveform):
        return record
    if kind == 'stack':
        return WaveVStack.fromlist(list(record))
    if isinstance(record, tuple) and len(record) == 2 and isinstance(record[1], tuple):
        return Waveform.fromtree(record)
    if isinstance(record, (list, tuple)):
        return Waveform.fromlist(list(record))
    raise TypeError(f'not a waveform, a tolist() list or a totree() tree: {type(record).__name__}')


def sample_wire(records, kinds=None, **kw):
    """``sample_batch`` for channels given in the reference's wire formats (see
    ``from_wire``); ``kinds[i] = 'stack'`` marks ``WaveVStack.tolist`` records."""
    kinds = kinds or [None] * len(records)
    return sample_batch([from_wire(r, k) for r, k in zip(records, kinds)], **kw)


def sample_pulse_trains(templates, tmpl_idx, t0, start, stop, sample_rate,
                        dtype=np.float64, devices=None, params=None, compact='auto'):
    """Channels given as PARAMETER ARRAYS instead of objects: channel ``c`` is the
    stack of pulses ``templates[tmpl_idx[c][k]]`` started at ``t0[c][k]``
    (``builder.PulseTemplate.trace`` / ``builder.pulse_train_batch``), sampled on
    ``np.arange(start, stop, 1/sample_rate)``; ``params = {name: array like t0}``
    carries the templates' further per-pulse parameters (amplitude, phase ...).
    Same sharding and result type as ``sample_batch``; the tables are those the
    object API would have produced.  ``compact``: upload templates + per-pulse payloads and write the per-pulse rows
    on the device (``builder.CompactBatch``, ``wfm_expand_templates``) — 'auto' = whenever the channels allow it
    (time-ordered, non-overlapping pulses), True = insist, False = build the full tables on the host."""
    import torch
    from .builder import pulse_train_batch
    engine.require_gpu()
    if devices is None:
        devices = [torch.cuda.current_device()]
    code = {np.dtype(np.float64): engine.WFM_F64,
            np.dtype(np.float32): engine.WFM_F32,
            np.dtype(np.complex128): engine.WFM_C128}[np.dtype(dtype)]
    ranges = shard_ranges([max(len(r), 1) for r in t0], len(devices))
    shards = [(dev, lo, hi) for dev, (lo, hi) in zip(devices, ranges)]

    def run(shard):
        dev, lo, hi = shard
        args = (templates, tmpl_idx[lo:hi], t0[lo:hi], start, stop, sample_rate)
        kw = dict(params={k: v[lo:hi] for k, v in (params or {}).items()})
        if compact == 'auto':
            try:
                batch = pulse_train_batch(*args, compact=True, **kw)
            except ValueError as e:
                if 'compact' not in str(e):
                    raise
                batch = pulse_train_batch(*args, **kw)
        else:
            batch = pulse_train_batch(*args, compact=bool(compact), **kw)
        with torch.cuda.device(dev):
            prog = engine.Program(batch, dev)
            out = prog.sample_device(dtype=code)
        return prog, out, batch

    results = _run_shards(run, shards)
    tensors, table = [], []
    for slot, ((dev, lo, hi), (prog, out, batch)) in enumerate(zip(shards, results)):
        tensors.append(out)
        for k in range(batch.n_channels):
            table.append((slot, int(batch.chan_off[k]), int(batch.chan_n[k])))
    for prog, _, _ in results:
        prog.close()
    return BatchResult(tensors, table, dtype)


def rank_shard(weights, rank=None, world=None):
    """The contiguous channel range ``(lo, hi)`` this process owns when the job
    runs as one process per GPU (torchrun): RANK / WORLD_SIZE from the
    environment unless given.  Every rank computes the same partition from the
    same weights, so no exchange is needed to agree on it."""
    import os
    rank = int(os.environ.get('RANK', '0')) if rank is None else int(rank)
    world = int(os.environ.get('WORLD_SIZE', '1')) if world is None else int(world)
    if not 0 <= rank < world:
        raise ValueError(f'rank {rank} outside world of {world}')
    return shard_ranges(weights, world)[rank]


def lower_rank_shard(waveforms, rank=None, world=None, sample_rate=None):
    """Host half of the one-process-per-GPU path: lower only this rank's shard.
    Returns ``(lo, hi, LoweredBatch)``; ``sample_batch_rank`` runs it."""
    items = [channel_grid(w, sample_rate) for w in waveforms]
    lo, hi = rank_shard([g.n for _, g in items], rank, world)
    return lo, hi, lower(items[lo:hi])


def sample_batch_rank(waveforms, rank=None, world=None, sample_rate=None,
                      dtype=np.float64, device=None, filters='own'):
    """One process per GPU: sample this rank's shard of ``waveforms`` on
    ``device`` (default: LOCAL_RANK).  Returns ``(lo, hi, BatchResult)`` with the
    result's channel ``i`` = ``waveforms[lo + i]``.  No collective is issued."""
    import os
    if device is None:
        device = int(os.environ.get('LOCAL_RANK', '0'))
    lo, hi = rank_shard([channel_grid(w, sample_rate)[1].n for w in waveforms],
                        rank, world)
    res = sample_batch(waveforms[lo:hi], sample_rate=sample_rate, dtype=dtype,
                       devices=[device], filters=filters)
    return lo, hi, res
