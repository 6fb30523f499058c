"""Worst relative Z error vs the CPU oracle for each TT config under several Jacobi stopping thresholds.
    python scripts/parity_sweep.py [stop_rel ...]     (GPU; default 3e-4 1e-4 0)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import numpy as np
import torch

import tta_runtime as rt
import workloads
from helpers import rel_fro
from oracle import port

DEV = 'cuda:0'


def run(key, stop_rel, oracle_cache):
    from admm import ADMM
    wb, hb, fmt = workloads.CONFIGS[key]
    weights = wb()
    rt.jacobi_set_stop_rel(stop_rel)
    model = workloads.ParamBag(weights, device=DEV)
    a = ADMM(model, 1e-3, hb(), fmt, DEV)
    if key not in oracle_cache:
        o = port.OracleADMM({n: w.numpy() for n, w in weights.items()}, 1e-3, hb(), fmt)
        zs = []
        o.update(update_u=False)
        zs.append({n: o.z[n].copy() for n in weights})
        for _ in range(2):
            o.update()
            zs.append({n: o.z[n].copy() for n in weights})
        oracle_cache[key] = zs
    zs = oracle_cache[key]
    errs = []
    a.update(update_u=False)
    errs.append({n: rel_fro(a.z[n].cpu().numpy(), zs[0][n]) for n in weights})
    for i in range(2):
        a.update()
        errs.append({n: rel_fro(a.z[n].cpu().numpy(), zs[i + 1][n]) for n in weights})
    worst = [max(e.values()) for e in errs]
    arg = max(errs[2], key=errs[2].get)
    sw = max(max(v) for v in a.sweeps.values())
    print('%-22s stop_rel %-8g worst Z err per update: %.2e %.2e %.2e  (%s)  sweeps max %d'
          % (key, stop_rel, worst[0], worst[1], worst[2], arg, sw), flush=True)


if __name__ == '__main__':
    stops = [float(v) for v in sys.argv[1:]] or [3e-4, 1e-4, 0.0]
    cache = {}
    for key in ('deit_small_tt', 'resnet50_tt', 'resnet32_tt'):
        for s in stops:
            run(key, s, cache)
    rt.jacobi_set_stop_rel(3e-4)
