"""Times the fp32 GEMM operator (tta_gemm_batched) on the task shapes of the ResNet-50 TT-general projection, tensor-core
route (csrc/gemm_tf32.cu) against the CUDA-core kernel (csrc/gemm.cu), one task per launch, L2-cold between launches."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200')]
import tta_runtime as rt  # noqa: E402

DEV = 'cuda:0'
# (label, M, N, K, a K-contiguous?, b N-contiguous?)
SHAPES = [('conv3x3 step1 carry', 30, 73728, 32, True, True),
          ('conv3x3 step2 carry', 105, 4608, 480, True, True),
          ('conv3x3 step3 core', 945, 105, 512, True, False),
          ('conv3x3 step4 core', 1680, 30, 32, True, False),
          ('1x1 512x1024 carry', 130, 1024, 512, True, True),
          ('1x1 2048x512 core', 2048, 130, 512, True, False),
          ('recon 32x30 . 30x1680', 32, 1680, 30, True, True),
          ('recon 512x105 . 105x945', 512, 945, 105, True, True),
          ('recon 4608x105 . 105x480', 4608, 480, 105, True, True),
          ('recon 73728x30 . 30x32', 73728, 32, 30, True, True),
          ('tucker C=256 P0', 2304, 128, 256, True, False),
          ('tucker C=512 P0', 4608, 256, 512, True, False),
          ('tucker C=1024 P0', 9216, 512, 1024, True, False),
          ('tucker C=1024 P1', 512, 9216, 1024, True, True),
          ('tucker C=2048 P0', 18432, 1024, 2048, True, False)]


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for label, M, N, K, ak, bn in SHAPES:
        a = torch.randn(M, K, device=DEV) if ak else torch.randn(K, M, device=DEV)
        b = torch.randn(K, N, device=DEV) if bn else torch.randn(N, K, device=DEV)
        c = torch.empty(M, N, device=DEV)
        tab = np.zeros(1, dtype=rt.GEMM_TASK)
        sai, sak = (K, 1) if ak else (1, M)
        sbk, sbj = (N, 1) if bn else (1, K)
        tab[0] = (a.data_ptr(), b.data_ptr(), c.data_ptr(), 0, sai, sak, sbk, sbj, N, M, N, K, 0)
        table = rt.TaskTable(tab, DEV)
        res = {}
        for tc in (True, False):
            rt.gemm_enable_tc(2 if tc else 0)
            ts = []
            for it in range(7):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rt.gemm(table)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            res[tc] = float(np.median(ts[2:]))
        rt.gemm_enable_tc(1)
        fl = 2.0 * M * N * K
        by = 4.0 * (M * K + K * N + M * N)
        print('%-26s M=%6d N=%6d K=%4d  tc %7.1f us (%6.1f TF/s, %5.0f GB/s)   cc %7.1f us (%5.1f TF/s)' % (
            label, M, N, K, res[True], fl / res[True] / 1e6, by / res[True] / 1e3, res[False], fl / res[False] / 1e6), flush=True)


if __name__ == '__main__':
    main()
