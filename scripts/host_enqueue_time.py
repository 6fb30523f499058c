"""Host time of ADMM.update() (enqueue only) against its device time: is the update launch-bound on the host?"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200')]
import hp_tables  # noqa: E402
import workloads  # noqa: E402
from admm import ADMM  # noqa: E402

dev = 'cuda:0'
model = workloads.ParamBag(workloads.resnet50_weights(seed=0), device=dev)
admm = ADMM(model, 1e-3, hp_tables.tt_resnet50_general_3x().fresh(), 'tt', dev)
for _ in range(3):
    admm.update()
torch.cuda.synchronize()
host, total = [], []
for _ in range(10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    admm.update()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    host.append((t1 - t0) * 1e3)
    total.append((t2 - t0) * 1e3)
print('host time after enqueueing each layer group (ms):', [round(v, 2) for v in admm.enqueue_ms_per_group])
print('groups:', [(len(nm), nm[0]) for pl, nm in admm._plans])
print('update(): host return after %.2f ms (min %.2f), device done after %.2f ms (min %.2f)' % (
    sum(host) / len(host), min(host), sum(total) / len(total), min(total)))
