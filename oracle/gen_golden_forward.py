"""Generate tests/golden/forward_modules.{json,npz}: outputs of the UNMODIFIED reference layer modules
(/root/reference TTConv.py, TTLinear.py, TKConv.py, TKLinear.py, imported through oracle/ref_shims.py) on small
seeded inputs, built through their `dense_w` constructors.

    python -m oracle.gen_golden_forward            # from the repo root (build container only)

Every case stores the constructor arguments, the dense weight / bias, the input x and the reference output y
(fp32).  tests/test_gpu_forward.py feeds the same arguments to this repo's drop-in modules and compares y.
This pins, in particular, `TTConv2dR`'s untransposed decomposition (TTConv.py:284-288,321: for k x k kernels it is
a DIFFERENT projection from admm.py:96) against the reference itself.  The TK* modules run on the restated
tensorly `partial_tucker` (oracle/port.py): their cases are labelled `pinned: false`.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'dnn-compression-tensor-admm_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import ref_shims  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


class Hp:
    def __init__(self, ranks, tt_shapes=None):
        self.ranks = ranks
        if tt_shapes is not None:
            self.tt_shapes = tt_shapes


def conv_cases():
    # (module, cin, cout, k, stride, pad, H, tt_shapes, ranks, bias): tt_resnet32_hp 3x / tk_resnet32_hp 3x entries
    # of layer2.0.conv1 (16 -> 32, stride 2), layer3.0.conv2 (64 -> 64) and a 1x1 case
    # (TTConv2dM adds its bias as `out += self.bias` on an NCHW tensor, TTConv.py:150-151: it only broadcasts when
    # W' == out_channels, so the reference itself cannot run that module with a bias -- cases use bias=False)
    tt = [('TTConv2dM', 16, 32, 3, 2, 1, 16, [8, 4, 9, 4, 4], [1, 8, 32, 16, 4, 1], False),
          ('TTConv2dM', 64, 64, 3, 1, 1, 8, [8, 8, 9, 8, 8], [1, 8, 28, 28, 8, 1], False),
          ('TTConv2dR', 16, 32, 3, 2, 1, 16, [8, 4, 9, 4, 4], [1, 8, 32, 16, 4, 1], True),
          ('TTConv2dR', 64, 64, 3, 1, 1, 8, [8, 8, 9, 8, 8], [1, 8, 28, 28, 8, 1], False),
          ('TTConv2dR', 24, 40, 1, 1, 0, 8, [40, 1, 24], [1, 10, 10, 1], True)]
    tk = [(m, 16, 32, 3, 2, 1, 16, None, [18, 9], True) for m in ('TKConv2dC', 'TKConv2dM', 'TKConv2dR')] + \
         [(m, 64, 64, 3, 1, 1, 8, None, [24, 20], False) for m in ('TKConv2dC', 'TKConv2dM', 'TKConv2dR')]
    # SVDConv.py: 1 x 1 convolutions, rank list of length 1 (as resnet_inet_tt.py:48-50 passes it) or a bare int
    svd = [('SVDConv2dR', 24, 40, 1, 1, 0, 8, None, [10], True), ('SVDConv2dC', 24, 40, 1, 1, 0, 8, None, [10], True),
           ('SVDConv2dM', 24, 40, 1, 1, 0, 8, None, 10, False), ('SVDConv2dM', 64, 64, 1, 1, 0, 8, None, [20], True)]
    return tt + tk + svd


def linear_cases():
    # numeric_example3.py shapes (in [5,7,9], out [4,8,16]) and a DeiT-like split; TK with ranks [r_out, r_in]
    return [('TTLinearM', 315, 512, [4, 8, 16, 5, 7, 9], [1, 2, 3, 4, 2, 2, 1], True),
            ('TTLinearR', 315, 512, [4, 8, 16, 5, 7, 9], [1, 2, 3, 4, 2, 2, 1], True),
            ('TTLinearM', 96, 144, [12, 12, 8, 12], [1, 10, 40, 9, 1], False),
            ('TTLinearR', 96, 144, [12, 12, 8, 12], [1, 10, 40, 9, 1], True),
            ('TKLinearM', 96, 144, None, [30, 20], True),
            ('TKLinearR', 96, 144, None, [30, 20], False)]


def main():
    torch.set_num_threads(8)
    ref = ref_shims.load_reference()
    g = torch.Generator().manual_seed(20211012)      # seed of numeric_example3.py:12
    cases, arrays = [], {}
    shared = {}

    def inputs(sig, make):
        """cases of one geometry share their dense weight / bias / input (smaller fixtures)"""
        if sig not in shared:
            shared[sig] = make()
        return shared[sig]

    for ci, (mod, cin, cout, k, stride, pad, hw, shapes, ranks, bias) in enumerate(conv_cases()):
        name = 'conv{}.weight'.format(ci)
        w, bb, x = inputs(('conv', cin, cout, k, hw), lambda: (
            torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5,
            torch.randn(cout, generator=g) * 0.1, torch.randn(2, cin, hw, hw, generator=g)))
        b = bb if bias else None
        hp = Hp({name: list(ranks) if not isinstance(ranks, int) else ranks},
                {name: list(shapes)} if shapes is not None else None)
        cls = getattr(ref.TTConv if mod.startswith('TT') else ref.TKConv if mod.startswith('TK') else ref.SVDConv, mod)
        layer = cls(cin, cout, k, stride=stride, padding=pad, bias=bias, hp_dict=hp, name=name,
                    dense_w=w.clone(), dense_b=b.clone() if b is not None else None)
        with torch.no_grad():
            y = layer(x)
        key = 'c{}'.format(ci)
        src = 'conv_{}_{}_{}_{}'.format(cin, cout, k, hw)
        arrays[src + '|w'] = w.numpy()
        arrays[src + '|x'] = x.numpy()
        arrays[src + '|b'] = bb.numpy()
        arrays[key + '|y'] = y.numpy().astype(np.float32)
        cases.append({'key': key, 'inputs': src, 'kind': 'conv', 'module': mod, 'in': cin, 'out': cout, 'kernel': k, 'stride': stride,
                      'padding': pad, 'bias': bias, 'tt_shapes': shapes,
                      'ranks': list(ranks) if not isinstance(ranks, int) else ranks,
                      'ranks_after': [int(v) for v in hp.ranks[name]] if not isinstance(hp.ranks[name], int) else hp.ranks[name],
                      'pinned': not mod.startswith('TK'),
                      'state_shapes': {n: list(p.shape) for n, p in layer.state_dict().items()}})
    for li, (mod, fin, fout, shapes, ranks, bias) in enumerate(linear_cases()):
        name = 'fc{}.weight'.format(li)
        w, bb, x = inputs(('fc', fin, fout), lambda: (torch.randn(fout, fin, generator=g) * 0.05,
                                                      torch.randn(fout, generator=g) * 0.1,
                                                      torch.randn(2, 7, fin, generator=g)))
        b = bb if bias else None
        hp = Hp({name: list(ranks)}, {name: list(shapes)} if shapes is not None else None)
        cls = getattr(ref.TTLinear if mod.startswith('TT') else ref.TKLinear, mod)
        layer = cls(fin, fout, bias=bias, hp_dict=hp, name=name, dense_w=w.clone(),
                    dense_b=b.clone() if b is not None else None)
        with torch.no_grad():
            y = layer(x)
        key = 'l{}'.format(li)
        src = 'fc_{}_{}'.format(fin, fout)
        arrays[src + '|w'] = w.numpy()
        arrays[src + '|x'] = x.numpy()
        arrays[src + '|b'] = bb.numpy()
        arrays[key + '|y'] = y.numpy().astype(np.float32)
        cases.append({'key': key, 'inputs': src, 'kind': 'linear', 'module': mod, 'in': fin, 'out': fout, 'bias': bias,
                      'tt_shapes': shapes, 'ranks': list(ranks), 'ranks_after': [int(v) for v in hp.ranks[name]],
                      'pinned': mod.startswith('TT'),
                      'state_shapes': {n: list(p.shape) for n, p in layer.state_dict().items()}})
    # ---- orthogonal.py:9-20 on a seeded toy model: value and gradients of the regulariser ----
    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.first_factor = torch.nn.Parameter(torch.randn(6, 20, generator=g) * 0.3)
            self.last_factor = torch.nn.Parameter(torch.randn(40, 9, generator=g) * 0.3)
            self.first_kernel = torch.nn.Parameter(torch.randn(5, 33, 1, 1, generator=g) * 0.3)
            self.left_kernel = torch.nn.Parameter(torch.randn(37, 37, 1, 1, generator=g) * 0.2)
            self.core_kernel = torch.nn.Parameter(torch.randn(4, 4, 3, 3, generator=g))

    toy = Toy()
    rho = 0.05
    loss = ref.orthogonal.append_double_l2_loss(toy, torch.zeros(()), rho, 'cpu')
    loss.backward()
    for n, p in toy.named_parameters():
        arrays['orth|' + n] = p.detach().numpy()
        if p.grad is not None:
            arrays['orth|grad|' + n] = p.grad.numpy()
    cases.append({'key': 'orth', 'kind': 'orth', 'module': 'orthogonal', 'rho': rho, 'loss': float(loss), 'pinned': True,
                  'params': [n for n, _ in toy.named_parameters()]})
    np.savez_compressed(os.path.join(GOLD, 'forward_modules.npz'), **arrays)
    with open(os.path.join(GOLD, 'forward_modules.json'), 'w') as f:
        json.dump(cases, f, sort_keys=True, indent=0)
    print('wrote', len(cases), 'cases;', sum(a.nbytes for a in arrays.values()) // 1024, 'KiB of arrays')


if __name__ == '__main__':
    main()
