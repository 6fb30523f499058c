"""The critical chain alone: ADMM update of the three layer4 3x3 convolutions of ResNet-50 TT-general
(eigenproblems 32 -> 480 -> 512 -> 32 per layer).  For ncu launch lists / timing of the between-phase kernels.

    python scripts/chain_only.py [n_updates]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200')):
    sys.path.insert(0, p)
import torch

import workloads
from admm import ADMM

DEV = 'cuda:0'
wb, hb, fmt = workloads.CONFIGS['resnet50_tt']
names = ['layer4.{}.conv2.weight'.format(i) for i in range(3)]
weights = {n: w for n, w in wb(seed=0).items() if n in names}
admm = ADMM(workloads.ParamBag(weights, device=DEV), 1e-3, hb(), fmt, DEV)
admm.update(update_u=False)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ts = []
for _ in range(n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    admm.update()
    b.record()
    b.synchronize()
    ts.append(a.elapsed_time(b))
print('chain-only update ms:', ' '.join('{:.2f}'.format(t) for t in ts))
print('jacobi sweeps per TT step:', {n: v for n, v in admm.sweeps.items()})
# timeline of one more update (CUDA events recorded by the plan, ms after the start)
plan = admm._plans[0][0]
plan.trace = []
t0 = torch.cuda.Event(enable_timing=True)
t0.record()
admm.update()
torch.cuda.synchronize()
print('timeline:', ' '.join('{}@{:.2f}'.format(lbl, t0.elapsed_time(ev)) for lbl, ev in plan.trace))
