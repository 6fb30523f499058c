"""Seeded synthetic weight sets for the BASELINE.json configs (no dataset / checkpoint access).

Each builder returns an ordered dict name -> fp32 CPU tensor with the parameter names and shapes of
the architecture the config names, initialised with that architecture's own init law:

  * resnet50   torchvision.models.resnet50(weights=None) conv init: kaiming_normal_(fan_out, relu)
               (what main.py:79-80 builds in --admm mode); only the 34 layers listed in
               hp_dicts/tt_resnet50_hp.py are materialised.
  * resnet32   resnet_cifar.py:15-19 `kaiming_normal_` (fan_in) on every conv of resnet_cifar.py:76-161.
  * deit_small timm VisionTransformer linear init trunc_normal_(std=.02); shapes from
               tt_deit_small_patch16_224_hp.py:7-9 / vit_tt.py:240-241.
  * tucker sweep: nn.Conv2d(C, C, 3) + kaiming_normal_ (config 5).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

import hp_tables


def _gen(seed):
    g = torch.Generator(device='cpu')
    g.manual_seed(seed)
    return g


def _normal(shape, std, g):
    return torch.randn(shape, generator=g, dtype=torch.float32) * std


def resnet50_weights(seed=0):
    g = _gen(seed)
    out = OrderedDict()
    table = hp_tables.tt_resnet50_general_3x()
    for name, shp in table.tt_shapes.items():
        if name.endswith('conv2.weight'):
            o = 1
            for s in shp[:2]:
                o *= s
            w = (o, o, 3, 3)
        else:
            w = (shp[0], shp[2], 1, 1)
        fan_out = w[0] * w[2] * w[3]
        out[name] = _normal(w, math.sqrt(2.0 / fan_out), g)
    return out


def resnet32_weights(seed=0):
    g = _gen(seed)
    out = OrderedDict()
    planes = {1: 16, 2: 32, 3: 64}
    for s in (1, 2, 3):
        for b in range(5):
            for c in (1, 2):
                cin = planes[s]
                if b == 0 and c == 1 and s > 1:
                    cin = planes[s - 1]
                shape = (planes[s], cin, 3, 3)
                out['layer{}.{}.conv{}.weight'.format(s, b, c)] = _normal(shape, math.sqrt(2.0 / (cin * 9)), g)
    return out


def deit_small_weights(seed=0):
    g = _gen(seed)
    out = OrderedDict()
    d = 384
    for blk in range(12):
        p = 'blocks.{}.'.format(blk)
        for suffix, shape in (('attn.qkv.weight', (3 * d, d)), ('attn.proj.weight', (d, d)),
                              ('mlp.fc1.weight', (4 * d, d)), ('mlp.fc2.weight', (d, 4 * d))):
            out[p + suffix] = _normal(shape, 0.02, g).clamp_(-2.0, 2.0)
    return out


def tucker_sweep_weight(channels, seed=0):
    g = _gen(seed + channels)
    return OrderedDict(weight=_normal((channels, channels, 3, 3), math.sqrt(2.0 / (channels * 9)), g))


def tucker_sweep_weights(seed=0, channels=(64, 128, 256, 512, 1024, 2048), fracs=(0.25, 0.5)):
    """BASELINE config 5 as ONE model: a C x C x 3 x 3 weight per (C, rank fraction) of the sweep -- independent tensors,
    so the sweep shards over GPUs like the layers of a network (sharding.LayerSharding)."""
    out = OrderedDict()
    for c in channels:
        for f in fracs:
            g = _gen(seed + c)
            out['c{}_r{}.weight'.format(c, int(c * f))] = _normal((c, c, 3, 3), math.sqrt(2.0 / (c * 9)), g)
    return out


class ParamBag(torch.nn.Module):
    """Minimal stand-in for a model: exposes `named_parameters()` with the given names.

    `ADMM` only ever calls `model.named_parameters()` (admm.py:35,43,81).  Dots in names are kept by
    storing the parameters in a plain dict rather than as registered attributes.
    """

    def __init__(self, weights, device='cpu'):
        super().__init__()
        self._bag = OrderedDict((n, torch.nn.Parameter(w.detach().clone().to(device))) for n, w in weights.items())

    def named_parameters(self, prefix='', recurse=True, remove_duplicate=True):
        for n, p in self._bag.items():
            yield n, p

    def parameters(self, recurse=True):
        return iter(self._bag.values())


CONFIGS = {
    # key -> (weights builder, hp table builder, format)
    'resnet32_tk': (resnet32_weights, lambda: hp_tables.tk_resnet32('3'), 'tk'),
    'resnet32_tk2': (resnet32_weights, lambda: hp_tables.tk_resnet32('2'), 'tk'),
    'resnet32_tt': (resnet32_weights, hp_tables.tt_resnet32_3x, 'tt'),
    'resnet50_tt': (resnet50_weights, hp_tables.tt_resnet50_general_3x, 'tt'),
    'resnet50_tt_special': (resnet50_weights, hp_tables.tt_resnet50_special_3x, 'tt'),
    'deit_small_tt': (deit_small_weights, hp_tables.tt_deit_small_2x, 'tt'),
}
