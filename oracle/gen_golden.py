"""Generate tests/golden/* by executing the UNMODIFIED reference (/root/reference) in this container.

    python -m oracle.gen_golden            # from the repo root

The reference cannot travel to the GPU box, so its outputs on the seeded BASELINE workloads are
committed as small fixtures: for every listed layer the Frobenius norm, the sum and a fixed 32-point
probe of the reference Z (after `update(update_u=False)` and after a second full `update()`), the
penalty value, plus full Z arrays for a few small layers and known-answer vectors for ten2tt/tt2ten.
The TT / SVD / dual / penalty numbers come from the real reference code.  The Tucker numbers come from
the reference's admm.py running on the restated tensorly `partial_tucker` (oracle/port.py) and are
labelled `pinned: false`.
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

import hp_tables  # noqa: E402
import workloads  # noqa: E402
from oracle import ref_shims  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
N_PROBE = 32


def probe_index(numel, seed=0):
    rng = np.random.RandomState(seed + numel % 9973)
    return rng.randint(0, numel, size=N_PROBE)


def summarize(z):
    z = np.asarray(z, dtype=np.float32)
    flat = z.reshape(-1)
    return {'fro': float(np.sqrt(np.sum(flat.astype(np.float64) ** 2))),
            'sum': float(np.sum(flat.astype(np.float64))),
            'probe': [float(v) for v in flat[probe_index(flat.size)]]}


REF_TABLES = {
    'resnet32_tk': ('tk_resnet32_hp', 'HyperParamsDictRatio3x'),
    'resnet32_tk2': ('tk_resnet32_hp', 'HyperParamsDictRatio2x'),
    'resnet32_tt': ('tt_resnet32_hp', 'HyperParamsDictRatio3x'),
    'resnet50_tt': ('tt_resnet50_hp', 'HyperParamsDictGeneralRatio3x'),
    'resnet50_tt_special': ('tt_resnet50_hp', 'HyperParamsDictSpecialRatio3x'),
    'deit_small_tt': ('tt_deit_small_patch16_224_hp', 'HyperParamsDictRatio2x'),
}


class _Fresh:
    """Deep copy of a reference hp class (ten2tt mutates the class-level rank lists)."""

    def __init__(self, cls):
        import copy
        self.ranks = copy.deepcopy(dict(cls.ranks))
        if hasattr(cls, 'tt_shapes'):
            self.tt_shapes = copy.deepcopy(dict(cls.tt_shapes))


def run_config(ref, key, full_layers):
    wb, _, fmt = workloads.CONFIGS[key]
    weights = wb()
    hp = _Fresh(ref.hp_class(*REF_TABLES[key]))
    model = workloads.ParamBag(weights)
    admm = ref.admm.ADMM(model, 1e-3, hp, fmt, 'cpu')
    admm.update(update_u=False)
    z0 = {n: admm.z[n].numpy().copy() for n in admm.z}
    admm.update()          # U_1 = W - Z_0
    admm.update()          # Z_2 = Proj(W + U_1), U_2 = U_1 + W - Z_2
    loss = admm.append_admm_loss(torch.zeros((), dtype=torch.float32))
    out = {'format': fmt, 'rho': 1e-3, 'pinned': fmt != 'tk', 'layers': {},
           'penalty_after_3_updates': float(loss),
           'ranks_after': {n: [int(r) for r in (hp.ranks[n] if not isinstance(hp.ranks[n], int) else [hp.ranks[n]])]
                           for n in hp.ranks}}
    arrays = {}
    for n in admm.z:
        out['layers'][n] = {'z0': summarize(z0[n]), 'z2': summarize(admm.z[n].numpy()),
                            'u2': summarize(admm.u[n].numpy())}
        if n in full_layers:
            arrays[n + '|z0'] = z0[n]
            arrays[n + '|z2'] = admm.z[n].numpy()
            arrays[n + '|u2'] = admm.u[n].numpy()
    return out, arrays


def ttd_kats(ref):
    """Known-answer vectors for ten2tt / tt2ten incl. the in-place rank clip (ttd.py:18-19)."""
    rng = np.random.RandomState(20210915)   # seed of numeric_example2.py:12
    cases = []
    arrays = {}
    specs = [((8, 4, 9, 4, 4), [1, 8, 16, 10, 4, 1]),       # numeric_example2.py-like shapes
             ((5, 7, 9), [1, 9, 4, 1]),                     # r1 = 9 > 5 -> clipped to 5
             ((4, 8, 16), [1, 2, 3, 1]),                    # numeric_example3.py out shapes
             ((12, 3, 6), [1, 12, 6, 1])]                   # full rank -> exact reconstruction
    for ci, (shape, ranks) in enumerate(specs):
        x = rng.randn(*shape).astype(np.float32)
        r = list(ranks)
        cores = ref.ttd.ten2tt(x, list(shape), r)
        rec = ref.ttd.tt2ten(cores, shape)
        arrays['case{}|x'.format(ci)] = x
        arrays['case{}|rec'.format(ci)] = rec.astype(np.float32)
        cases.append({'shape': list(shape), 'ranks_in': list(ranks), 'ranks_out': [int(v) for v in r],
                      'core_shapes': [list(c.shape) for c in cores]})
    return cases, arrays


def hp_dump(ref):
    out = {}
    pairs = {'tt_resnet50_general_3x': ('tt_resnet50_hp', 'HyperParamsDictGeneralRatio3x'),
             'tt_resnet50_special_3x': ('tt_resnet50_hp', 'HyperParamsDictSpecialRatio3x'),
             'tt_resnet32_3x': ('tt_resnet32_hp', 'HyperParamsDictRatio3x'),
             'tk_resnet32_1p5x': ('tk_resnet32_hp', 'HyperParamsDictRatio1p5x'),
             'tk_resnet32_2x': ('tk_resnet32_hp', 'HyperParamsDictRatio2x'),
             'tk_resnet32_3x': ('tk_resnet32_hp', 'HyperParamsDictRatio3x'),
             'tk_resnet32_5x': ('tk_resnet32_hp', 'HyperParamsDictRatio5x'),
             'tt_deit_small_2x': ('tt_deit_small_patch16_224_hp', 'HyperParamsDictRatio2x')}
    for key, (mod, cls) in pairs.items():
        c = ref.hp_class(mod, cls)
        out[key] = {'ranks': {n: [int(v) for v in r] for n, r in c.ranks.items()}}
        if hasattr(c, 'tt_shapes'):
            out[key]['tt_shapes'] = {n: [int(v) for v in s] for n, s in c.tt_shapes.items()}
    return out


def main():
    torch.set_num_threads(8)
    ref = ref_shims.load_reference()
    os.makedirs(GOLD, exist_ok=True)
    with open(os.path.join(GOLD, 'hp_tables.json'), 'w') as f:
        json.dump(hp_dump(ref), f, sort_keys=True)
    full = {'resnet32_tt': ['layer1.0.conv1.weight', 'layer2.0.conv1.weight', 'layer3.4.conv2.weight'],
            'resnet32_tk': ['layer1.0.conv1.weight', 'layer2.0.conv1.weight', 'layer3.0.conv1.weight'],
            'resnet32_tk2': [], 'resnet50_tt': ['layer1.1.conv2.weight'], 'resnet50_tt_special': [],
            'deit_small_tt': []}
    summary = {}
    for key in full:
        print('running reference on', key, flush=True)
        out, arrays = run_config(ref, key, set(full[key]))
        summary[key] = out
        if arrays:
            np.savez_compressed(os.path.join(GOLD, key + '_arrays.npz'), **arrays)
    with open(os.path.join(GOLD, 'reference_summary.json'), 'w') as f:
        json.dump(summary, f, sort_keys=True)
    cases, arrays = ttd_kats(ref)
    np.savez_compressed(os.path.join(GOLD, 'ttd_kats.npz'), **arrays)
    with open(os.path.join(GOLD, 'ttd_kats.json'), 'w') as f:
        json.dump(cases, f, sort_keys=True)
    print('golden fixtures written to', GOLD)


if __name__ == '__main__':
    main()
