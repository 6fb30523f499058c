"""tta_gemm_bf16_tn (weight-gradient GEMM, csrc/gemm_tn.cu) against torch.mm (cuBLAS) on the DeiT-small gradient shapes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200')]
import tta_runtime as rt  # noqa: E402

dev = 'cuda:0'


def t(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / iters


K = 256 * 197
for (M, N) in ((1152, 256), (256, 384), (384, 256), (1536, 256), (256, 1536), (320, 384)):
    a = torch.randn(K, M, device=dev).to(torch.bfloat16)
    b = torch.randn(K, N, device=dev).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev)
    ms_own = t(lambda: rt.gemm_bf16_tn(a, b, out, M, N, K))
    ms_lib = t(lambda: torch.mm(a.t(), b, out_dtype=torch.float32))
    fl = 2.0 * M * N * K
    print('dW %4d x %4d, %d tokens: own %.3f ms (%.0f TFLOP/s)   torch.mm %.3f ms (%.0f TFLOP/s)' % (
        M, N, K, ms_own, fl / ms_own / 1e9, ms_lib, fl / ms_lib / 1e9), flush=True)
