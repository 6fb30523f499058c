"""ncu target: one tensor-core Gram (k = 480, reduction 4608, W + U in place) and one tensor-core GEMM (105 x 4608 x 480,
the step-2 projection of a layer4 3x3 convolution), each launched twice (the second launch is the one to read)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200')]
import projector  # noqa: E402
import tta_runtime as rt  # noqa: E402

DEV = 'cuda:0'


def main():
    k, red = 480, 4608
    w = torch.randn(k, red, device=DEV)
    u = torch.randn(k, red, device=DEV) * 0.1
    nsplit = projector.gram_splits(k, red)
    part = torch.empty(nsplit * k * k, dtype=torch.float64, device=DEV)
    g64 = torch.empty(k * k, dtype=torch.float64, device=DEV)
    tab = np.zeros(1, dtype=rt.GRAM_TASK)
    tab[0] = (w.data_ptr(), part.data_ptr(), 0, g64.data_ptr(), red, 0, 1, k, 1, red, nsplit, k, k, u.data_ptr())
    gt = rt.TaskTable(tab, DEV)
    M, N, K = 105, 4608, 480
    a = torch.randn(M, K, device=DEV)
    b = torch.randn(K, N, device=DEV)
    c = torch.empty(M, N, device=DEV)
    mm = np.zeros(1, dtype=rt.GEMM_TASK)
    mm[0] = (a.data_ptr(), b.data_ptr(), c.data_ptr(), 0, K, 1, N, 1, N, M, N, K, 0)
    mt = rt.TaskTable(mm, DEV)
    for _ in range(2):
        rt.gram(gt)
        rt.gemm(mt)
        torch.cuda.synchronize()
    print('nsplit', nsplit, 'ok')


if __name__ == '__main__':
    main()
