"""Stage timeline of every fp64 eigensolver call of one ADMM.update(): reduce / eigenvalues / vectors / back-transformation
(ms after the start of the update), for the full ResNet-50 update and for the critical chain alone."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200')]
import tta_runtime as rt  # noqa: E402
import workloads  # noqa: E402
from admm import ADMM  # noqa: E402

dev = 'cuda:0'
wb, hb, fmt = workloads.CONFIGS['resnet50_tt']


def trace(weights, label):
    admm = ADMM(workloads.ParamBag(weights, device=dev), 1e-3, hb(), fmt, dev)
    for _ in range(3):
        admm.update()
    torch.cuda.synchronize()
    lib = rt.lib()
    lib.tta_symeig_profile_enable(2)
    origin = torch.cuda.Event(enable_timing=True)
    end = torch.cuda.Event(enable_timing=True)
    origin.record()
    admm.update()
    end.record()
    torch.cuda.synchronize()
    out = np.zeros(7 * 64, dtype=np.float64)
    n = lib.tta_symeig_stage_profile_read(ctypes.c_void_p(origin.cuda_event), ctypes.c_void_p(out.ctypes.data), 64)
    lib.tta_symeig_profile_enable(0)
    print('== %s: update %.2f ms' % (label, origin.elapsed_time(end)))
    for r in out[:7 * n].reshape(n, 7):
        print('  k<=%4d x%2d  start %.2f | reduce %.2f | eigval %.2f | eigvec %.2f | backtr %.2f   (end %.2f)' % (
            r[0], r[1], r[2], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[6] - r[5], r[6]))


full = wb(seed=0)
trace(full, 'full network')
names = ['layer4.{}.conv2.weight'.format(i) for i in range(3)]
trace({n: w for n, w in full.items() if n in names}, 'critical chain alone')
