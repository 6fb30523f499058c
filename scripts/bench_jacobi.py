"""GPU micro-benchmark of the batched Jacobi eigensolver: time, sweeps and accuracy per problem size.

    python scripts/bench_jacobi.py [k:count ...]      (default: 512:6 256:12 130:3 64:4)
Compares the gram-rotate-apply cluster kernel with the column-rotation cluster kernel on the same
matrices (CUDA events around the batched call, best of 3 after one warm-up).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200')):
    sys.path.insert(0, p)
import numpy as np
import torch

import projector
import tta_runtime as rt

DEV = 'cuda:0'


def problems(spec, seed=0):
    rng = np.random.RandomState(seed)
    out = []
    for k, cnt in spec:
        for _ in range(cnt):
            A = rng.randn(k, 4 * k).astype(np.float32)
            out.append((k, (A.astype(np.float64) @ A.astype(np.float64).T)))
    return out


def run(probs, gra, repeats=3):
    etab = np.zeros(len(probs), dtype=rt.EIG_TASK)
    xs, x0 = [], []
    for i, (k, G) in enumerate(probs):
        ld, kpad, bw = projector.eig_geometry(k) if gra else projector.eig_geometry(k, pmax=16)
        X = np.zeros((kpad, ld), dtype=np.float32)
        X[:k, :k] = G.T.astype(np.float32)
        x = torch.from_numpy(X.reshape(-1)).to(DEV)
        xs.append((x, ld, kpad))
        x0.append(x.clone())
        etab[i] = (x.data_ptr(), k, ld, kpad, bw)
    tab = rt.TaskTable(etab, DEV)
    scratch = torch.empty(rt.jacobi_scratch_bytes(tab) // 4 + 16, dtype=torch.int32, device=DEV)
    rt.jacobi_enable_gra(gra)
    best = 1e30
    try:
        for it in range(repeats + 1):
            for (x, _, _), src in zip(xs, x0):
                x.copy_(src)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sweeps = rt.jacobi_eigh(tab, scratch, tol=5e-7, max_sweeps=40)
            e1.record()
            torch.cuda.synchronize()
            if it:
                best = min(best, e0.elapsed_time(e1))
    finally:
        rt.jacobi_enable_gra(True)
    worst_orth, worst_lam = 0.0, 0.0
    for (k, G), (x, ld, kpad) in zip(probs, xs):
        X = x.cpu().numpy().reshape(kpad, ld).astype(np.float64)[:, :k]
        nrm = np.linalg.norm(X, axis=1)
        order = np.argsort(-nrm)[:k]
        Q = X[order] / nrm[order, None]
        worst_orth = max(worst_orth, np.abs(Q @ Q.T - np.eye(k)).max())
        ref = np.linalg.eigvalsh(G)[::-1]
        worst_lam = max(worst_lam, np.abs(nrm[order] - ref).max() / ref[0])
    return best, sweeps, worst_orth, worst_lam


if __name__ == '__main__':
    spec = [tuple(int(v) for v in a.split(':')) for a in sys.argv[1:]] or [(512, 6), (256, 12), (130, 3), (64, 4)]
    for group in [[s] for s in spec] + [spec]:
        probs = problems(group)
        for gra in (True, False):
            ms, sweeps, orth, lam = run(probs, gra)
            print('%-28s %-16s %8.3f ms  sweeps %s  max|QtQ-I| %.2e  max dlam/lam0 %.2e'
                  % (group, 'gram-rotate-apply' if gra else 'column-rotation', ms,
                     sorted(set(int(s) for s in sweeps)), orth, lam), flush=True)
