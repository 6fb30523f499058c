"""Rank / TT-shape tables for the BASELINE configs, built by rule instead of by literal listing.

The drop-in `admm.ADMM` accepts *any* object exposing `.ranks[name]` (and `.tt_shapes[name]` for
TT) -- in particular the reference's own `hp_dicts/*.py` classes keep working unchanged.  This module
only restates the tables that BASELINE.json's configs name, so that tests and `bench.py` run on a box
where `/root/reference` does not exist.  Equality with the reference tables is pinned by
`tests/golden/hp_tables.json` (dumped from the reference by `oracle/gen_golden.py`).

Tables (reference file:line):
  * tt_resnet50 general / special 3x   -- hp_dicts/tt_resnet50_hp.py:82-155 / :6-79
  * tt_resnet32 3x                     -- hp_dicts/tt_resnet32_hp.py:10-207 (ranks, tt_shapes only)
  * tk_resnet32 1.5x / 2x / 3x / 5x    -- hp_dicts/tk_resnet32_hp.py:7,40,74,107
  * tt_deit_small_patch16_224 2x       -- hp_dicts/tt_deit_small_patch16_224_hp.py:6-111
"""
from __future__ import annotations


class HpTable:
    """Plain holder with the reference's attribute protocol (`ranks`, `tt_shapes`)."""

    def __init__(self, label, ranks, tt_shapes=None):
        self.label = label
        self.ranks = ranks
        if tt_shapes is not None:
            self.tt_shapes = tt_shapes

    def fresh(self):
        """Deep copy -- `ttd.ten2tt` clips rank lists in place (ttd.py:18-19)."""
        import copy
        return HpTable(self.label, copy.deepcopy(self.ranks),
                       copy.deepcopy(getattr(self, 'tt_shapes', None)))

    def __repr__(self):
        return 'HpTable({}, {} layers)'.format(self.label, len(self.ranks))


_R50_BLOCKS = {1: 3, 2: 4, 3: 6, 4: 3}
_R50_PLANES = {1: 64, 2: 128, 3: 256, 4: 512}


def _r50_names_3x3():
    return ['layer{}.{}.conv2.weight'.format(s, b) for s in (1, 2, 3, 4) for b in range(_R50_BLOCKS[s])]


def _r50_1x1(stage, which):
    return ['layer{}.{}.{}.weight'.format(stage, b, which) for b in range(_R50_BLOCKS[stage])]


def _r50_in_channels(stage, block):
    # bottleneck conv1 input: previous stage's 4*planes for block 0, own 4*planes otherwise
    if block == 0:
        return _R50_PLANES[stage - 1] * 4 if stage > 1 else 64
    return _R50_PLANES[stage] * 4


def tt_resnet50_general_3x():
    shapes, ranks = {}, {}
    split = {1: ([8, 8], [8, 8]), 2: ([16, 8], [8, 16]), 3: ([16, 16], [16, 16]), 4: ([32, 16], [16, 32])}
    edge = {1: 8, 2: 15, 3: 15, 4: 30}
    mid0 = {1: 64, 2: 80, 3: 82, 4: 105}
    midn = {1: 55, 2: 70, 3: 82, 4: 105}
    for s in (1, 2, 3, 4):
        for b in range(_R50_BLOCKS[s]):
            n = 'layer{}.{}.conv2.weight'.format(s, b)
            shapes[n] = split[s][0] + [9] + split[s][1]
            m = mid0[s] if b == 0 else midn[s]
            ranks[n] = [1, edge[s], m, m, edge[s], 1]
    r11 = {3: 75, 4: 130}
    for s in (3, 4):
        p = _R50_PLANES[s]
        for b in range(_R50_BLOCKS[s]):
            n = 'layer{}.{}.conv1.weight'.format(s, b)
            shapes[n] = [p, 1, _r50_in_channels(s, b)]
            ranks[n] = [1, r11[s], r11[s], 1]
        for b in range(_R50_BLOCKS[s]):
            n = 'layer{}.{}.conv3.weight'.format(s, b)
            shapes[n] = [4 * p, 1, p]
            ranks[n] = [1, r11[s], r11[s], 1]
    # reference ordering: all 3x3, then layer3 conv1, layer3 conv3, layer4 conv1, layer4 conv3
    return HpTable('tt_resnet50_general_3x', _ordered(ranks, _r50_order()), _ordered(shapes, _r50_order()))


def tt_resnet50_special_3x():
    shapes, ranks = {}, {}
    mid0 = {1: 60, 2: 80, 3: 82, 4: 100}
    midn = {1: 50, 2: 70, 3: 82, 4: 100}
    for s in (1, 2, 3, 4):
        p = _R50_PLANES[s]
        for b in range(_R50_BLOCKS[s]):
            n = 'layer{}.{}.conv2.weight'.format(s, b)
            shapes[n] = [p, 9, p]
            m = mid0[s] if b == 0 else midn[s]
            ranks[n] = [1, m, m, 1]
    lo, hi = {3: 40, 4: 70}, {3: 85, 4: 160}
    for s in (3, 4):
        p = _R50_PLANES[s]
        for b in range(_R50_BLOCKS[s]):
            n = 'layer{}.{}.conv1.weight'.format(s, b)
            shapes[n] = [p, 1, _r50_in_channels(s, b)]
            ranks[n] = [1, lo[s], hi[s], 1]
            n = 'layer{}.{}.conv3.weight'.format(s, b)
            shapes[n] = [4 * p, 1, p]
            ranks[n] = [1, hi[s], lo[s], 1]
    return HpTable('tt_resnet50_special_3x', _ordered(ranks, _r50_order()), _ordered(shapes, _r50_order()))


def _r50_order():
    return (_r50_names_3x3() + _r50_1x1(3, 'conv1') + _r50_1x1(3, 'conv3')
            + _r50_1x1(4, 'conv1') + _r50_1x1(4, 'conv3'))


def _ordered(d, order):
    return {k: d[k] for k in order}


def _r32_names():
    return ['layer{}.{}.conv{}.weight'.format(s, b, c) for s in (1, 2, 3) for b in range(5) for c in (1, 2)]


def tt_resnet32_3x():
    shapes, ranks = {}, {}
    l3_mid = {'layer3.0.conv2': 27, 'layer3.1.conv1': 27, 'layer3.1.conv2': 27, 'layer3.2.conv1': 27,
              'layer3.2.conv2': 28, 'layer3.3.conv1': 28, 'layer3.3.conv2': 29, 'layer3.4.conv1': 24,
              'layer3.4.conv2': 15}
    for n in _r32_names():
        key = n[:-len('.weight')]
        if key.startswith('layer1'):
            shapes[n], ranks[n] = [16, 9, 16], [1, 16, 16, 1]
        elif key == 'layer2.0.conv1':
            shapes[n], ranks[n] = [8, 4, 9, 4, 4], [1, 8, 32, 16, 4, 1]
        elif key.startswith('layer2'):
            shapes[n], ranks[n] = [8, 4, 9, 4, 8], [1, 8, 16, 16, 8, 1]
        elif key == 'layer3.0.conv1':
            shapes[n], ranks[n] = [8, 8, 9, 4, 8], [1, 8, 40, 24, 8, 1]
        else:
            m = l3_mid[key]
            shapes[n], ranks[n] = [8, 8, 9, 8, 8], [1, 8, m, m, 8, 1]
    return HpTable('tt_resnet32_3x', ranks, shapes)


def tk_resnet32(ratio):
    """`ratio` in {'1p5', '2', '3', '5'} (utils.py:258-400 spells 1.5 as '1p5')."""
    ratio = str(ratio)
    ranks = {}
    for n in _r32_names():
        s, b, c = int(n[5]), int(n[7]), int(n[13])
        first = (b == 0 and c == 1)
        if ratio == '2':
            r = {1: [8, 8], 2: [16, 16], 3: [32, 32]}[s]
        elif ratio == '1p5':
            if s == 1:
                r = [16, 16]
            elif s == 2:
                r = [30, 16] if first else [30, 30]
            else:
                if first:
                    r = [40, 30]
                elif (b, c) in ((0, 2), (3, 1), (3, 2), (4, 1), (4, 2)):
                    r = [38, 38]
                else:
                    r = [39, 39]
        elif ratio == '3':
            if s == 1:
                r = [16, 16]
            elif s == 2:
                r = [24, 20] if first else [20, 20]
            else:
                r = [32, 25] if first else [25, 23]
        elif ratio == '5':
            if s == 1:
                r = [16, 16] if b == 0 else [11, 11]
            elif s == 2:
                r = [16, 13] if first else [13, 13]
            else:
                if first:
                    r = [24, 18]
                elif b in (0, 1):
                    r = [18, 18]
                elif b == 2:
                    r = [17, 17]
                else:
                    r = [16, 16]
        else:
            raise KeyError('tk_resnet32 ratio {!r}'.format(ratio))
        ranks[n] = list(r)
    return HpTable('tk_resnet32_{}x'.format(ratio), ranks)


def tt_deit_small_2x():
    shapes, ranks = {}, {}
    for blk in range(12):
        first = blk == 0
        mid = 320 if first else 256
        a, bq, f = (22, 23, 42) if first else (18, 18, 30)
        q = 35 if first else 25
        p = 'blocks.{}.'.format(blk)
        shapes[p + 'attn.qkv.weight'], ranks[p + 'attn.qkv.weight'] = (36, 32, 16, 24), (1, q, mid, bq, 1)
        shapes[p + 'attn.proj.weight'], ranks[p + 'attn.proj.weight'] = (24, 16, 16, 24), (1, a, mid, a, 1)
        shapes[p + 'mlp.fc1.weight'], ranks[p + 'mlp.fc1.weight'] = (48, 32, 16, 24), (1, f, mid, a, 1)
        shapes[p + 'mlp.fc2.weight'], ranks[p + 'mlp.fc2.weight'] = (24, 16, 32, 48), (1, a, mid, f, 1)
    return HpTable('tt_deit_small_patch16_224_2x', ranks, shapes)


def tucker_sweep(channels, frac):
    """Config 5: one C x C x 3 x 3 conv, ranks r_out = r_in = frac * C."""
    r = int(channels * frac)
    return HpTable('tk_sweep_C{}_r{}'.format(channels, r), {'weight': [r, r]})


def tucker_sweep_all(channels=(64, 128, 256, 512, 1024, 2048), fracs=(0.25, 0.5)):
    """Config 5 as one table: names as in workloads.tucker_sweep_weights."""
    return HpTable('tk_sweep_all', {'c{}_r{}.weight'.format(c, int(c * f)): [int(c * f), int(c * f)]
                                    for c in channels for f in fracs})


ALL = {
    'tt_resnet50_general_3x': tt_resnet50_general_3x,
    'tt_resnet50_special_3x': tt_resnet50_special_3x,
    'tt_resnet32_3x': tt_resnet32_3x,
    'tk_resnet32_1p5x': lambda: tk_resnet32('1p5'),
    'tk_resnet32_2x': lambda: tk_resnet32('2'),
    'tk_resnet32_3x': lambda: tk_resnet32('3'),
    'tk_resnet32_5x': lambda: tk_resnet32('5'),
    'tt_deit_small_2x': tt_deit_small_2x,
}
