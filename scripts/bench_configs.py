"""ADMM Z+U update time of every BASELINE.json config on one B200 (CUDA events, median of `--steps`).

    python scripts/bench_configs.py [--steps 5] [--sweep 64,128,256,512]
Prints one JSON line per config: milliseconds per `ADMM.update()` and layers/s; for the Tucker-2 sweep
(config 5) one line per channel count / rank ratio.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200')):
    sys.path.insert(0, p)
import numpy as np
import torch

import hp_tables
import workloads
from admm import ADMM

DEV = 'cuda:0'


def time_updates(admm, steps):
    admm.update(update_u=False)
    admm.update()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        admm.update()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--sweep', default='64,128,256,512')
    args = ap.parse_args()
    for key, (wb, hb, fmt) in workloads.CONFIGS.items():
        model = workloads.ParamBag(wb(seed=0), device=DEV)
        hp = hb()
        admm = ADMM(model, 1e-3, hp.fresh() if hasattr(hp, 'fresh') else hp, fmt, DEV)
        ms = time_updates(admm, args.steps)
        n = len(admm._names)
        sw = [s for v in admm.sweeps.values() for s in v]
        print(json.dumps({'config': key, 'format': fmt, 'layers': n, 'elements': admm._total_numel, 'ms_per_update': ms,
                          'layers_per_s': n / (ms / 1e3), 'jacobi_sweeps_max': int(max(sw)) if sw else None}), flush=True)
        del admm, model
    for c in [int(v) for v in args.sweep.split(',') if v]:
        for ratio in (0.25, 0.5):
            r = int(c * ratio)
            model = workloads.ParamBag(workloads.tucker_sweep_weight(c), device=DEV)
            hp = hp_tables.tucker_sweep(c, ratio)
            admm = ADMM(model, 1e-3, hp, 'tk', DEV)
            ms = time_updates(admm, max(2, args.steps // 2))
            plan = admm._plans[0][0]
            print(json.dumps({'config': 'tucker2_sweep', 'channels': c, 'rank': r, 'ms_per_update': ms,
                              'hooi_sweeps': plan.hooi_sweeps.get('weight')}), flush=True)
            del admm, model


if __name__ == '__main__':
    main()
