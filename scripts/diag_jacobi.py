"""Diagnostic (GPU): run the Jacobi solver for a fixed number of sweeps and report invariants."""
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


def run(k, max_sweeps, multilaunch=False, seed=0):
    rng = np.random.RandomState(seed)
    A = rng.randn(k, 3 * k)
    G = (A @ A.T).astype(np.float32)
    ld, kpad, bw = projector.eig_geometry(k)
    X = np.zeros((kpad, ld), dtype=np.float32)
    X[:k, :k] = G.T
    x = torch.from_numpy(X.reshape(-1)).to(DEV)
    etab = np.zeros(1, dtype=rt.EIG_TASK)
    etab[0] = (x.data_ptr(), k, ld, kpad, bw)
    tab = rt.TaskTable(etab, DEV)
    scratch = torch.empty(rt.jacobi_scratch_bytes(tab) // 4 + 16, dtype=torch.int32, device=DEV)
    rt.jacobi_force_multilaunch(multilaunch)
    try:
        sw = rt.jacobi_eigh(tab, scratch, tol=5e-7, max_sweeps=max_sweeps)
        status = 'converged'
    except rt.TtaError as e:
        sw = [max_sweeps]
        status = 'NOT converged'
    finally:
        rt.jacobi_force_multilaunch(False)
    Xo = x.cpu().numpy().reshape(kpad, ld).astype(np.float64)[:, :k]     # rows = columns x_j
    G64 = G.astype(np.float64)
    inv = np.linalg.norm(Xo.T @ Xo - G64 @ G64) / np.linalg.norm(G64 @ G64)   # X X^T must stay G^2
    gram = Xo @ Xo.T
    nrm = np.sqrt(np.maximum(np.diag(gram), 1e-300))
    cos = gram / np.outer(nrm, nrm)
    np.fill_diagonal(cos, 0)
    live = nrm > 1e-6 * nrm.max()
    off = np.abs(cos[np.ix_(live, live)]).max()
    print('k=%d geom=%s %s max_sweeps=%d -> %s sweeps=%d | invariant err %.2e | max |cos| %.2e | live cols %d'
          % (k, (ld, kpad, bw), 'multilaunch' if multilaunch else 'cluster', max_sweeps, status, sw[0], inv, off, live.sum()))


if __name__ == '__main__':
    for k in (32, 64, 130):
        for ms in (1, 2, 4, 8, 40):
            run(k, ms)
        run(k, 40, multilaunch=True)
