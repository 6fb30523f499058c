"""One shape of the persistent TMA GEMM behind tta_gemm_bf16_tc, a few launches (for ncu):  python scripts/prof_gemm_tma.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200')):
    sys.path.insert(0, p)
import torch

import tta_runtime as rt

DEV = 'cuda:0'
M, N, K = 256 * 197, 1120, 320
a = torch.randn(M, K, device=DEV).to(torch.bfloat16)
b = torch.randn(N, K, device=DEV).to(torch.bfloat16)
c = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
for _ in range(4):
    rt.gemm_bf16_tc(a, b, c, M, N, K)
torch.cuda.synchronize()
