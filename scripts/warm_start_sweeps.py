"""Jacobi sweeps and update time over successive ADMM updates, with and without the warm start.

    python scripts/warm_start_sweeps.py [config]
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
key = sys.argv[1] if len(sys.argv) > 1 else 'resnet50_tt'
wb, hb, fmt = workloads.CONFIGS[key]
for warm in (False, True):
    model = workloads.ParamBag(wb(seed=0), device=DEV)
    hp = hb()
    admm = ADMM(model, 1e-3, hp.fresh() if hasattr(hp, 'fresh') else hp, fmt, DEV)
    admm.update(update_u=False)
    for plan, _ in admm._plans:
        if hasattr(plan, 'warm_start'):
            plan.warm_start = warm
    line = []
    for it in range(12):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        admm.update()
        b.record()
        b.synchronize()
        sw = [s for v in admm.sweeps.values() for s in v]
        line.append('{:.1f}ms/{}'.format(a.elapsed_time(b), max(sw)))
    print('warm' if warm else 'cold', key, ' '.join(line), flush=True)
    for n, v in admm.sweeps.items():
        if 'layer4' in n or 'layer3.0' in n or 'blocks.0.' in n:
            print('   ', n, v)
    del admm, model
