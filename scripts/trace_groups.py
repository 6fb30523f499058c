"""Timeline of the layer groups of one ADMM.update(): per group and wave, when the eigensolver phase starts
and ends on the device (CUDA events, ms after the fork), plus the host time spent enqueueing.

    python scripts/trace_groups.py [config]      config: resnet50_tt (default), deit_small_tt, ...
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200')):
    sys.path.insert(0, p)
import torch

import workloads
from admm import ADMM

DEV = 'cuda:0'
key = sys.argv[1] if len(sys.argv) > 1 else 'resnet50_tt'
wb, hb, fmt = workloads.CONFIGS[key]
model = workloads.ParamBag(wb(seed=0), device=DEV)
hp = hb()
admm = ADMM(model, 1e-3, hp.fresh() if hasattr(hp, 'fresh') else hp, fmt, DEV)
admm.concurrent_groups = (len(sys.argv) <= 2 or sys.argv[2] != 'single')
admm.update(update_u=False)
for _ in range(3):
    admm.update()
torch.cuda.synchronize()
for gi, (plan, names) in enumerate(admm._plans):
    ks = [[st['k'] for st in w['steps'] if not st['identity']] for w in plan.ws]
    print('group', gi, len(names), 'layers; eig sizes per layer:', sorted(set(map(tuple, ks))))
ts = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    admm.update()
    b.record()
    b.synchronize()
    ts.append(a.elapsed_time(b))
print('untraced update(): median {:.2f} ms  min {:.2f} ms'.format(sorted(ts)[len(ts) // 2], min(ts)))
for gi, (plan, names) in enumerate(admm._plans):
    plan.trace = []
t0 = torch.cuda.Event(enable_timing=True)
t0.record()
h0 = time.perf_counter()
admm.update()
h1 = time.perf_counter()
t1 = torch.cuda.Event(enable_timing=True)
t1.record()
torch.cuda.synchronize()
print('host time of update() {:.2f} ms; device time {:.2f} ms'.format((h1 - h0) * 1e3, t0.elapsed_time(t1)))
for gi, (plan, names) in enumerate(admm._plans):
    print('group', gi, ' '.join('{}@{:.2f}'.format(lbl, t0.elapsed_time(ev)) for lbl, ev in plan.trace))
