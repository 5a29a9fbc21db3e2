import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests/golden')
import numpy as np, torch, bench
from waveforms_b200 import engine
from waveforms_b200.batch import channel_grid
from waveforms_b200.lowering import lower, replicate
ns = bench.b200_namespace()
rng = np.random.default_rng(1)
chans = []
for q in range(4):
    pulses = None
    for k in range(50):
        I, Q = ns.mixing(rng.uniform(0.1,1)*ns.cosPulse(20e-9) >> (100e-9 + 400e-9*k), freq=50e6, phase=rng.uniform(0,6), DRAGScaling=4e-10)
        w = I + 1j*Q
        pulses = w if pulses is None else pulses + w
    pulses.start, pulses.stop, pulses.sample_rate = 0, 20.5e-6, 2e9
    chans.append(pulses)
base = lower([channel_grid(w) for w in chans])
batch = replicate(base, 256)
prog = engine.Program(batch, 0)
out = torch.empty(batch.total_samples, dtype=torch.complex128, device='cuda')
prog.sample_device(dtype=engine.WFM_C128, out=out); torch.cuda.synchronize()
ev=[torch.cuda.Event(enable_timing=True) for _ in range(4)]
ev[0].record()
for k in range(3):
    prog.sample_device(dtype=engine.WFM_C128, out=out); ev[k+1].record()
torch.cuda.synchronize()
ms=min(ev[k].elapsed_time(ev[k+1]) for k in range(3))
n=int(batch.waves['n'].sum())
print('complex128:', n, 'samples', ms, 'ms', n/ms/1e6, 'GSa/s', n*16/ms/1e6, 'GB/s')
