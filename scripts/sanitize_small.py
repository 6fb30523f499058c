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
print('sanitize_small done')
