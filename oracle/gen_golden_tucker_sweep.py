"""tests/golden/tucker_sweep.json: the restated-tensorly oracle (oracle/port.py, PARITY UNPINNED) on the large
cases of BASELINE config 5 (3x3 conv weights with 512 / 1024 channels), as summary records: HOOI sweep count plus
Frobenius norm / sum / 32-point probe of Z.  The oracle needs 6 .. 50 s per case on the build container's cores, so
the GPU tests compare against these records instead of re-running it on the GPU box.

    python -m oracle.gen_golden_tucker_sweep
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'dnn-compression-tensor-admm_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import hp_tables  # noqa: E402
import workloads  # noqa: E402
from oracle import port  # noqa: E402
from oracle.gen_golden import summarize  # noqa: E402

CASES = [(512, 0.25), (512, 0.5), (1024, 0.25)]


def main():
    out = []
    for C, frac in CASES:
        w = workloads.tucker_sweep_weight(C)['weight'].numpy()
        ranks = hp_tables.tucker_sweep(C, frac).ranks['weight']
        t0 = time.time()
        z, sweeps = port.project_tk(w, ranks, return_sweeps=True)
        out.append({'C': C, 'frac': frac, 'ranks': [int(r) for r in ranks], 'hooi_sweeps': int(sweeps),
                    'pinned': False, 'z': summarize(z), 'oracle_seconds': round(time.time() - t0, 1)})
        print(out[-1]['C'], out[-1]['frac'], 'sweeps', sweeps, 'in', out[-1]['oracle_seconds'], 's', flush=True)
    with open(os.path.join(ROOT, 'tests', 'golden', 'tucker_sweep.json'), 'w') as f:
        json.dump(out, f, sort_keys=True)


if __name__ == '__main__':
    main()
