"""Small invocations of the tensor-core kernels and of one projection (a quick smoke of every kernel family;
also the input for `compute-sanitizer --tool memcheck python scripts/sanitize_small.py` where that tool is available)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200')):
    sys.path.insert(0, p)
import torch

import tta_runtime as rt
import workloads
from admm import ADMM

DEV = 'cuda:0'
torch.manual_seed(0)
for (M, K1, N1, N2) in ((300, 72, 40, 50), (513, 384, 320, 1152)):
    x = torch.randn(M, K1, device=DEV).to(torch.bfloat16)
    w1 = torch.randn(N1, K1, device=DEV).to(torch.bfloat16)
    w2 = torch.randn(N2, (N1 + 7) // 8 * 8, device=DEV).to(torch.bfloat16)
    for dt in (torch.float32, torch.bfloat16):
        ldy = (N2 + 7) // 8 * 8
        y = torch.zeros(M, ldy, device=DEV, dtype=dt)
        rt.lowrank2_fwd(x, w1, w2, torch.randn(N2, device=DEV), y, M, K1, N1, N2, ld2=w2.shape[1], ldy=ldy)
for (M, N, K) in ((300, 264, 40), (1000, 1120, 320)):
    a = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    b = torch.randn(N, K, device=DEV).to(torch.bfloat16)
    c = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    rt.gemm_bf16_tc(a, b, c, M, N, K, bias=torch.randn(N, device=DEV))
wb, hb, fmt = workloads.CONFIGS['resnet32_tt']
names = ['layer1.0.conv1.weight', 'layer3.1.conv2.weight']
weights = {n: w for n, w in wb().items() if n in names}
admm = ADMM(workloads.ParamBag(weights, device=DEV), 1e-3, hb(), fmt, DEV)
admm.update(update_u=False)
admm.update()
admm.update()
torch.cuda.synchronize()
# ---- round-2 tensor-core kernels: Gram (TMA + tcgen05 3xTF32), GEMM (cp.async + tcgen05 3xTF32), fused convolution,
# weight-gradient GEMM ----
import numpy as np
for (m, n, nsplit) in ((40, 1000, 3), (130, 520, 2), (300, 100, 1)):
    A = torch.randn(m, n, device=DEV)
    U = torch.randn(m, n, device=DEV)
    for mode in ('row', 'col'):
        k = m if mode == 'row' else n
        ld, kpad = (k + 3) // 4 * 4, (k + 15) // 16 * 16
        x = torch.zeros(kpad * ld, device=DEV)
        g64 = torch.zeros(k * k, dtype=torch.float64, device=DEV)
        part = torch.zeros(nsplit * k * k, dtype=torch.float64, device=DEV)
        tab = np.zeros(1, dtype=rt.GRAM_TASK)
        if mode == 'row':
            tab[0] = (A.data_ptr(), part.data_ptr(), x.data_ptr(), g64.data_ptr(), n, 0, 1, k, 1, n, nsplit, ld, kpad, U.data_ptr())
        else:
            tab[0] = (A.data_ptr(), part.data_ptr(), x.data_ptr(), g64.data_ptr(), 1, 0, n, k, 1, m, nsplit, ld, kpad, U.data_ptr())
        rt.gram(rt.TaskTable(tab, DEV))
rt.gemm_enable_tc(2)
for (M, N, K, ta, tb) in ((130, 1024, 130, False, False), (257, 129, 33, True, True), (49, 25, 8, False, True)):
    a = torch.randn(K, M, device=DEV) if ta else torch.randn(M, K, device=DEV)
    b = torch.randn(N, K, device=DEV) if tb else torch.randn(K, N, device=DEV)
    c = torch.zeros(M, N, device=DEV)
    tab = np.zeros(1, dtype=rt.GEMM_TASK)
    sai, sak = (1, M) if ta else (K, 1)
    sbk, sbj = (1, K) if tb else (N, 1)
    tab[0] = (a.data_ptr(), b.data_ptr(), c.data_ptr(), 0, sai, sak, sbk, sbj, N, M, N, K, 0)
    rt.gemm(rt.TaskTable(tab, DEV))
rt.gemm_enable_tc(1)
for (B, Cin, H, W, Ra, Rb, Cout, stride) in ((3, 16, 32, 32, 16, 16, 16, 1), (2, 7, 13, 9, 5, 6, 11, 1), (2, 16, 16, 16, 16, 32, 32, 2),
                                            (4, 64, 8, 8, 27, 29, 64, 1)):
    x = torch.randn(B, Cin, H, W, device=DEV)
    blob = rt.ttconv_tc_pack(torch.randn(Ra, Cin, device=DEV), torch.randn(Rb, Ra, 3, 3, device=DEV), torch.randn(Cout, Rb, device=DEV),
                             torch.randn(Cout, device=DEV))
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    y = torch.zeros(B, Cout, Ho, Wo, device=DEV)
    rt.ttconv_tc_fwd(x, blob, y, B, Cin, H, W, Ra, Rb, Cout, 3, stride, 1)
for (M, N, K, lda, ldb) in ((40, 72, 777, 48, 72), (256, 384, 3000, 256, 384)):
    a = torch.randn(K, lda, device=DEV).to(torch.bfloat16)
    b = torch.randn(K, ldb, device=DEV).to(torch.bfloat16)
    rt.gemm_bf16_tn(a, b, torch.zeros(M, N, device=DEV), M, N, K, lda=lda, ldb=ldb, ldc=N)
torch.cuda.synchronize()
print('sanitize_small done')
