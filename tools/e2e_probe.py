import sys, time, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests/golden')
import numpy as np, torch
import bench
from waveforms_b200 import engine
from waveforms_b200.batch import channel_grid
from waveforms_b200.lowering import lower, replicate
ns = bench.b200_namespace()
chans = bench.build_frame(ns)
frame = lower([channel_grid(w) for w in chans])
batch = replicate(frame, 64, amp_scale=np.random.default_rng(1).uniform(0.5,1,64)).pin()
host = torch.empty(batch.total_samples, dtype=torch.float64, pin_memory=True).numpy()
for i in range(4):
    t0=time.perf_counter(); p=engine.Program(batch,0); t1=time.perf_counter(); p.sample_host(out=host); t2=time.perf_counter(); p.close(); t3=time.perf_counter()
    print(f'create {1e3*(t1-t0):.1f} sample_host {1e3*(t2-t1):.1f} close {1e3*(t3-t2):.1f} ms', file=sys.stderr)
