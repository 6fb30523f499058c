"""One shape of the fused two-factor forward, a few launches (for ncu):  python scripts/prof_lowrank2.py [f32|bf16]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200')):
    sys.path.insert(0, p)
import torch

import tta_runtime as rt

DEV = 'cuda:0'
M, K1, N1, N2 = 256 * 197, 384, 320, 1152
out_f32 = (sys.argv[1] if len(sys.argv) > 1 else 'f32') == 'f32'
x = torch.randn(M, K1, device=DEV).to(torch.bfloat16)
w1 = torch.randn(N1, K1, device=DEV).to(torch.bfloat16)
w2 = torch.randn(N2, N1, device=DEV).to(torch.bfloat16)
bias = torch.randn(N2, device=DEV)
y = torch.empty(M, N2, device=DEV, dtype=torch.float32 if out_f32 else torch.bfloat16)
for _ in range(4):
    rt.lowrank2_fwd(x, w1, w2, bias, y, M, K1, N1, N2)
torch.cuda.synchronize()
