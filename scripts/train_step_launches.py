"""ncu target: one fused training step (forward + backward) of the 48 DeiT-small TTLinearM layers, batch 32 -- the launch
list shows which kernels the fused training path runs (profiles/: no cuBLAS / cuDNN kernel among them)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200')]
import hp_tables  # noqa: E402
import TTLinear  # noqa: E402

dev = 'cuda:0'
hp = hp_tables.tt_deit_small_2x()
d = 384
dims = {'attn.qkv.weight': (d, 3 * d), 'attn.proj.weight': (d, d), 'mlp.fc1.weight': (d, 4 * d), 'mlp.fc2.weight': (4 * d, d)}
tokens = 32 * 197
xs = {d: torch.randn(tokens, d, device=dev), 4 * d: torch.randn(tokens, 4 * d, device=dev)}
layers = []
for name in list(hp.ranks)[:8]:           # two transformer blocks are enough for the launch list
    fin, fout = dims[name.split('.', 2)[2]]
    layer = TTLinear.TTLinearM(fin, fout, bias=True, hp_dict=hp, name=name).to(dev)
    layer.fused_training = True
    layers.append((layer, xs[fin]))
for it in range(2):
    for layer, x in layers:
        xg = x.detach().requires_grad_(True)
        y = layer(xg)
        y.backward(torch.ones_like(y))
    torch.cuda.synchronize()
print('ok')
