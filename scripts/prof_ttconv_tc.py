"""ncu target: the tensor-core fused convolution on the ttm_resnet32 layer1 shape (batch 128, 16 channels, 32 x 32) and on a
64-channel 56 x 56 shape, three launches each (read the last)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200')]
import tta_runtime as rt  # noqa: E402

dev = 'cuda:0'
for (cin, hw, ra, rb, cout) in ((16, 32, 16, 16, 16), (64, 56, 40, 40, 64)):
    x = torch.randn(128, cin, hw, hw, device=dev)
    y = torch.empty(128, cout, hw, hw, device=dev)
    blob = rt.ttconv_tc_pack(torch.randn(ra, cin, device=dev), torch.randn(rb, ra, 3, 3, device=dev), torch.randn(cout, rb, device=dev), None)
    for _ in range(3):
        rt.ttconv_tc_fwd(x, blob, y, 128, cin, hw, hw, ra, rb, cout, 3, 1, 1)
    torch.cuda.synchronize()
print('ok')
