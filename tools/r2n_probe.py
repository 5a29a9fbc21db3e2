"""Which random programs of tests/test_gpu_random.py miss the fp32 tolerance, and where."""
import sys, warnings
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/tests/golden')
import numpy as np
from helpers import b200_namespace, rel_err
import test_gpu_random as T
from tools import bench_extras as X
from waveforms_b200 import sample_batch
from waveforms_b200.lowering import lower
from waveforms_b200.batch import channel_grid
ns = b200_namespace()
for seed in range(T.N_PROGRAMS):
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        w, mode, rng = T._program(ns, seed)
        if mode != 'sample': continue
        got = sample_batch([w], dtype=np.float32).numpy()[0].astype(np.float64)
        want = T._reference(X, w)
    e = rel_err(got, want)
    if not (e <= 1e-6):
        k = int(np.argmax(np.abs(got - want)))
        b = lower([channel_grid(w)])
        # the segment holding sample k
        xs = np.arange(len(want)) / w.sample_rate
        seg = int(np.searchsorted(b.seg_bound, xs[k] - (w.shift if hasattr(w, 'wlist') else 0), side='right'))
        p0, p1 = b.seg_ptr[seg], b.seg_ptr[seg + 1]
        amps = b.terms['amp_re'][p0['term']:p1['term']]
        print(seed, 'err', '%.2e' % e, 'max', '%.3g' % np.abs(want).max(), 'at', k, '%.8g %.8g' % (got[k], want[k]),
              'seg funcs', b.facs['func'][p0['fac']:p1['fac']].tolist(), 'amps', ['%.3g' % a for a in amps][:12],
              'a0', ['%.3g' % a for a in b.facs['a0'][p0['fac']:p1['fac']]][:8], 'stack' if hasattr(w, 'wlist') else 'wave')
