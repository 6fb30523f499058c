"""Error of the tcgen05 3xTF32 GEMM (fp32 TMEM accumulation) against fp64, as a function of the reduction length:
decides how often a tensor-core Gram kernel has to flush its accumulator (Gram on tcgen05)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200')]
import tta_runtime as rt  # noqa: E402

DEV = 'cuda:0'


def run(A, B, tc):
    M, K = A.shape
    N = B.shape[1]
    a, b = torch.from_numpy(A).to(DEV), torch.from_numpy(B).to(DEV)
    c = torch.zeros(M, N, device=DEV)
    tab = np.zeros(1, dtype=rt.GEMM_TASK)
    tab[0] = (a.data_ptr(), b.data_ptr(), c.data_ptr(), 0, K, 1, N, 1, N, M, N, K, 0)
    rt.gemm_enable_tc(2 if tc else 0)
    try:
        rt.gemm(rt.TaskTable(tab, DEV))
        torch.cuda.synchronize()
    finally:
        rt.gemm_enable_tc(1)
    return c.cpu().numpy().astype(np.float64)


rng = np.random.RandomState(0)
for K in (8, 16, 32, 64, 128, 256, 512, 2048, 8192):
    A = rng.randn(256, K).astype(np.float32)
    B = rng.randn(K, 256).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    gref = A.astype(np.float64) @ A.astype(np.float64).T
    line = 'K=%5d' % K
    for tc in (True, False):
        c = run(A, B, tc)
        g = run(A, np.ascontiguousarray(A.T), tc)
        dg = np.diag(g) / np.diag(gref) - 1.0
        scale = np.sqrt(np.outer(np.diag(gref), np.diag(gref)))
        off = (g - gref) / scale
        np.fill_diagonal(off, 0.0)
        line += ' | %s gemm %.2e  gram diag bias %.2e spread %.2e  offdiag rms %.2e' % (
            'tc' if tc else 'cc', np.linalg.norm(c - ref) / np.linalg.norm(ref), dg.mean(), dg.std(),
            np.sqrt((off ** 2).mean()))
    print(line, flush=True)
