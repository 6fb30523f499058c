"""Host-side logic (plans, dispatch, rank clipping, autograd wiring, sharding) driven end to end
through the C-ABI emulator (tests/fake_tta.py) and compared with the CPU oracle.  No GPU needed."""
import numpy as np
import pytest
import torch

import hp_tables
import projector
import sharding
import workloads
from helpers import rel_fro
from oracle import port


def _subset(weights, names):
    return {n: weights[n] for n in names}


@pytest.mark.parametrize('key,names', [
    ('resnet32_tt', None),
    ('resnet50_tt', ['layer1.0.conv2.weight', 'layer2.1.conv2.weight', 'layer3.0.conv1.weight',
                     'layer3.0.conv3.weight']),
    ('resnet50_tt_special', ['layer3.0.conv1.weight', 'layer3.0.conv3.weight', 'layer1.0.conv2.weight']),
    ('deit_small_tt', ['blocks.0.attn.qkv.weight', 'blocks.1.mlp.fc2.weight']),
])
def test_admm_tt_flow_matches_oracle(emulated_backend, key, names):
    from admm import ADMM
    wb, hb, fmt = workloads.CONFIGS[key]
    weights = wb()
    if names:
        weights = _subset(weights, names)
    hp, hp_o = hb(), hb()
    model = workloads.ParamBag(weights)
    a = ADMM(model, 1e-3, hp, fmt, 'cpu', log=True)
    o = port.OracleADMM({n: w.numpy() for n, w in weights.items()}, 1e-3, hp_o, fmt)
    for n in weights:                       # admm.py:32-40
        assert torch.equal(a.z[n], weights[n]) and float(a.u[n].abs().sum()) == 0.0
    a.update(update_u=False)
    o.update(update_u=False)
    for n in weights:
        assert float(a.u[n].abs().sum()) == 0.0
        assert rel_fro(a.z[n].numpy(), o.z[n]) <= 1e-5, n
    a.update()
    o.update()
    a.update()
    o.update()
    for n in weights:
        assert rel_fro(a.z[n].numpy(), o.z[n]) <= 2e-5, n
        # U is compared on the scale of the weights (full-rank layers have U = rounding noise)
        assert np.linalg.norm(a.u[n].numpy() - o.u[n]) <= 2e-5 * np.linalg.norm(o.z[n]), n
        assert len(a.logger[n]) == 2
        assert abs(a.logger[n][-1] - o.diff_norm[n]) <= 1e-4 * np.linalg.norm(o.z[n])
    # in-place rank clip reaches the caller's table for conv weights only (admm.py:94 vs :105)
    for n in weights:
        assert [int(v) for v in hp.ranks[n]] == [int(v) for v in hp_o.ranks[n]], n
    # penalty and its gradient (admm.py:80-85)
    base = torch.tensor(0.25, requires_grad=True)
    loss = a.append_admm_loss(base * 2.0)
    assert abs(float(loss.detach()) - (0.5 + o.penalty())) <= 1e-5 * (0.5 + o.penalty())
    (loss * 3.0).backward()                  # GradScaler-style scaled loss (engines.py:315)
    assert abs(float(base.grad) - 6.0) < 1e-6
    for n, p in model.named_parameters():
        # gradient rho*(W - Z + U) compared on the scale rho*||W|| (noise-only for full-rank layers)
        assert np.linalg.norm(p.grad.numpy() - 3.0 * o.penalty_grad(n)) <= 1e-4 * 3e-3 * np.linalg.norm(o.w[n]), n


def test_format_none_and_bad_rank_raise(emulated_backend):
    from admm import ADMM
    weights = _subset(workloads.resnet32_weights(), ['layer1.0.conv1.weight'])
    with pytest.raises(Exception, match='Tensor format should be specified'):
        ADMM(workloads.ParamBag(weights), 1e-3, hp_tables.tt_resnet32_3x(), 'none', 'cpu')
    bag = workloads.ParamBag({'b': torch.zeros(3, 4, 5)})
    hp = hp_tables.HpTable('x', {'b': [1, 2, 1]}, {'b': [3, 20]})
    a = ADMM(bag, 1e-3, hp, 'tt', 'cpu')
    with pytest.raises(Exception, match='unsupported layer'):
        a.update()


def test_svd_format_and_single_rank_dispatch(emulated_backend):
    from admm import ADMM
    g = torch.Generator().manual_seed(3)
    weights = {'fc.weight': torch.randn(24, 40, generator=g), 'pw.weight': torch.randn(32, 16, 1, 1, generator=g)}
    hp = hp_tables.HpTable('svd', {'fc.weight': 5, 'pw.weight': [7]})
    a = ADMM(workloads.ParamBag(weights), 1e-3, hp, 'svd', 'cpu')
    a.update()
    assert rel_fro(a.z['fc.weight'].numpy(), port.project_linear_svd(weights['fc.weight'].numpy(), 5)) <= 1e-5
    assert rel_fro(a.z['pw.weight'].numpy(), port.project_conv_svd(weights['pw.weight'].numpy(), [7])) <= 1e-5
    assert tuple(a.z['pw.weight'].shape) == (32, 16, 1, 1)
    # helper methods keep the reference's numpy conventions (admm.py:129-149)
    out = a.prune_conv_rank_svd(weights['pw.weight'], 'pw.weight')
    assert out.shape == (32, 16, 1, 1)
    assert rel_fro(out, port.project_conv_svd(weights['pw.weight'].numpy(), [7])) <= 1e-5


def test_ttd_dropin_conventions(emulated_backend):
    import ttd
    rng = np.random.RandomState(5)
    x = rng.randn(5, 7, 9).astype(np.float32)
    ranks = [1, 9, 4, 1]
    cores = ttd.ten2tt(x, [5, 7, 9], ranks)
    assert ranks == [1, 5, 4, 1]            # clipped in place (ttd.py:18-19)
    assert [c.shape for c in cores] == [(1, 5, 5), (5, 7, 4), (4, 9, 1)]
    ref_ranks = [1, 9, 4, 1]
    ref = port.tt_contract(port.tt_svd(x, [5, 7, 9], ref_ranks), x.shape)
    assert rel_fro(ttd.tt2ten(cores, x.shape), ref) <= 1e-5


def test_tt_step_flops_accounting():
    """ResNet-50 TT-general totals 18.63 / 7.02 / 5.44 GFLOP (gram / proj / recon), 17.64 eig."""
    hp = hp_tables.tt_resnet50_general_3x()
    tot = np.zeros(4)
    for n in hp.ranks:
        tot += np.array(projector.tt_step_flops(hp.tt_shapes[n], hp.ranks[n]), dtype=np.float64)
    assert abs(tot[0] / 1e9 - 18.63) < 0.05 and abs(tot[1] / 1e9 - 7.02) < 0.05
    assert abs(tot[3] / 1e9 - 5.44) < 0.05 and abs(tot[2] / 1e9 - 17.64) < 0.05


def test_lpt_sharding_is_deterministic_and_balanced():
    hp = hp_tables.tt_resnet50_general_3x()
    w = workloads.resnet50_weights()
    costs = [sharding.layer_cost('tt', w[n].shape, hp.ranks[n], hp.tt_shapes[n]) for n in hp.ranks]
    for world in (1, 2, 4, 8):
        owner, load = sharding.lpt_assign(costs, world)
        owner2, _ = sharding.lpt_assign(costs, world)
        assert owner == owner2 and set(owner) == set(range(world))
        assert max(load) <= sum(costs) / world + max(costs)
    _, load8 = sharding.lpt_assign(costs, 8)
    assert max(load8) / (sum(costs) / 8) < 1.35


@pytest.mark.parametrize('key', ['resnet32_tk', 'resnet32_tk2'])
def test_admm_tucker_flow_matches_oracle(emulated_backend, key):
    """Tucker-2 HOOI (restated tensorly semantics; parity unpinned at the tensorly boundary)."""
    from admm import ADMM
    wb, hb, fmt = workloads.CONFIGS[key]
    weights = wb()
    names = list(weights)[::3]
    weights = _subset(weights, names)
    hp, hp_o = hb(), hb()
    a = ADMM(workloads.ParamBag(weights), 1e-3, hp, fmt, 'cpu')
    o = port.OracleADMM({n: w.numpy() for n, w in weights.items()}, 1e-3, hp_o, fmt)
    for upd in (False, True, True):
        a.update(update_u=upd)
        o.update(update_u=upd)
    plan = a._plans[0][0]
    for n in weights:
        assert rel_fro(a.z[n].numpy(), o.z[n]) <= 2e-5, n
        assert np.linalg.norm(a.u[n].numpy() - o.u[n]) <= 2e-5 * np.linalg.norm(o.z[n]), n
        if plan.errors[n][-1] > 1e-3:          # lossy layer: the stopping rule is well conditioned
            assert plan.hooi_sweeps[n] == o.sweeps[n], (n, plan.hooi_sweeps[n], o.sweeps[n])
    # 2-D weights take the same path (admm.py:121-127)
    g = torch.Generator().manual_seed(7)
    lin = {'fc.weight': torch.randn(40, 56, generator=g)}
    hp_l = hp_tables.HpTable('tk_lin', {'fc.weight': [9, 11]})
    b = ADMM(workloads.ParamBag(lin), 1e-3, hp_l, 'tk', 'cpu')
    b.update()
    assert rel_fro(b.z['fc.weight'].numpy(), port.project_tk(lin['fc.weight'].numpy(), [9, 11])) <= 1e-5
    out = b.prune_linear_rank_tk(lin['fc.weight'].numpy(), 'fc.weight')
    assert rel_fro(out, port.project_tk(lin['fc.weight'].numpy(), [9, 11])) <= 1e-5


def test_layer_modules_autograd_path_on_cpu(emulated_backend):
    """Drop-in layer modules: constructor surface, state-dict names and the torch (training) path,
    checked on CPU against the oracle identities of SURVEY 3.4."""
    import TKConv
    import TKLinear
    import TTConv
    import TTLinear
    g = torch.Generator().manual_seed(21)
    hp = hp_tables.tt_deit_small_2x()
    name = 'blocks.1.attn.proj.weight'
    w = torch.randn(384, 384, generator=g) * 0.02
    lin = TTLinear.TTLinearM(384, 384, bias=True, hp_dict=hp.fresh(), name=name, dense_w=w, dense_b=torch.zeros(384))
    x = torch.randn(5, 384, generator=g, requires_grad=True)
    z = torch.from_numpy(port.project_linear_tt(w.numpy(), hp.tt_shapes[name], list(hp.ranks[name])))
    assert rel_fro(lin(x).detach().numpy(), torch.nn.functional.linear(x, z).detach().numpy()) <= 1e-4
    assert [tuple(c.shape) for c in lin.tt_cores] == [(1, 24, 18), (18, 16, 256), (256, 16, 18), (18, 24, 1)]
    hp32 = hp_tables.tt_resnet32_3x()
    cname = 'layer2.0.conv1.weight'
    wc = torch.randn(32, 16, 3, 3, generator=g) * 0.1
    conv = TTConv.TTConv2dM(16, 32, 3, stride=2, padding=1, bias=False, hp_dict=hp32.fresh(), name=cname, dense_w=wc)
    xc = torch.randn(2, 16, 8, 8, generator=g, requires_grad=True)
    zc = torch.from_numpy(port.project_conv_tt(wc.numpy(), hp32.tt_shapes[cname], list(hp32.ranks[cname])))
    ref = torch.nn.functional.conv2d(xc, zc, stride=2, padding=1)
    assert rel_fro(conv(xc).detach().numpy(), ref.detach().numpy()) <= 1e-4
    assert tuple(conv.core_kernel.shape) == (32, 16, 3, 3) and len(conv.in_tt_cores) == 2 and len(conv.out_tt_cores) == 2
    hpk = hp_tables.tk_resnet32('3')
    for cls in (TKConv.TKConv2dC, TKConv.TKConv2dM, TKConv.TKConv2dR):
        tk = cls(16, 32, 3, stride=2, padding=1, bias=False, hp_dict=hpk, name=cname, dense_w=wc)
        zk = torch.from_numpy(port.project_tk(wc.numpy(), hpk.ranks[cname]))
        refk = torch.nn.functional.conv2d(xc, zk, stride=2, padding=1)
        assert rel_fro(tk(xc).detach().numpy(), refk.detach().numpy()) <= 1e-4, cls.__name__
    with pytest.raises(ValueError):
        TKConv.TKConv2dC(16, 32, 3, groups=2, hp_dict=hpk, name=cname)
    hpl = hp_tables.HpTable('tkl', {'fc.weight': [6, 7]})
    wl = torch.randn(20, 24, generator=g)
    tkl = TKLinear.TKLinearM(24, 20, bias=True, hp_dict=hpl, name='fc.weight', dense_w=wl, dense_b=torch.zeros(20))
    xl = torch.randn(3, 24, generator=g, requires_grad=True)
    zl = torch.from_numpy(port.project_tk(wl.numpy(), [6, 7]))
    assert rel_fro(tkl(xl).detach().numpy(), torch.nn.functional.linear(xl, zl).detach().numpy()) <= 1e-4
    # inference path without a GPU must fail loudly, not fall back (emulator off)
    import tta_runtime as rt
    rt.set_backend_for_tests(None)
    with pytest.raises(rt.TtaError):
        with torch.no_grad():
            lin(x.detach())


def test_update_from_host_equals_update(emulated_backend):
    """ADMM.update_from_host(host_w, host_z): same Z / U as copying the weights in, update(), copying Z out."""
    from admm import ADMM
    wb, hb, fmt = workloads.CONFIGS['resnet32_tt']
    weights = _subset(wb(), ['layer1.0.conv1.weight', 'layer2.0.conv1.weight', 'layer3.1.conv2.weight'])
    a = ADMM(workloads.ParamBag({n: torch.zeros_like(w) for n, w in weights.items()}), 1e-3, hb(), fmt, 'cpu')
    b = ADMM(workloads.ParamBag(weights), 1e-3, hb(), fmt, 'cpu')
    host_z = {n: torch.empty_like(w) for n, w in weights.items()}
    for _ in range(2):
        a.update_from_host(weights, host_z)
        b.update()
    for n in weights:
        assert torch.equal(a.z[n], b.z[n]) and torch.equal(a.u[n], b.u[n]) and torch.equal(host_z[n], b.z[n])


def test_tt_layer_groups_partition():
    """Layer groups of ADMM.update(): a partition of the layers, classes by chain shape, most critical first."""
    import projector
    wb, hb, _ = workloads.CONFIGS['resnet50_tt']
    hp, weights = hb(), wb()
    layers = [projector.TTLayer(n, weights[n].shape, hp.tt_shapes[n], hp.ranks[n]) for n in hp.ranks]
    groups = projector.tt_layer_groups(layers)
    assert sorted(i for g in groups for i in g) == list(range(len(layers)))
    assert 2 <= len(groups) <= 6
    names = [[layers[i].name for i in g] for g in groups]
    assert all('layer4' in n and 'conv2' in n for n in names[0]) and len(names[0]) == 3     # the 480 -> 512 chain
    small = [g for g in names if any(n.startswith('layer1') for n in g)]
    assert len(small) == 1 and all(n.startswith(('layer1', 'layer2')) for n in small[0])
    # DeiT-small: 48 equal-shaped chains, chunked into GPU-fulls
    wb, hb, _ = workloads.CONFIGS['deit_small_tt']
    hp, weights = hb(), wb()
    layers = [projector.TTLayer(n, weights[n].shape, hp.tt_shapes[n], list(hp.ranks[n])) for n in hp.ranks]
    groups = projector.tt_layer_groups(layers)
    assert sorted(i for g in groups for i in g) == list(range(48)) and len(groups) <= 6


def test_state_dict_roundtrip_and_adjust_rho(emulated_backend):
    """Additive checkpointing of the ADMM state (the reference drops u / z on resume, engines.py:216-245) and
    `adjust_rho` (admm.py:87-89: rho = factor * init_rho once epoch > int(0.85 * epochs))."""
    from admm import ADMM
    wb, hb, fmt = workloads.CONFIGS['resnet32_tt']
    names = ['layer3.0.conv2.weight', 'layer3.1.conv1.weight']       # lossy layers: U moves on every update
    weights = _subset(wb(), names)
    a = ADMM(workloads.ParamBag(weights), 1e-3, hb(), fmt, 'cpu')
    a.update(update_u=False)
    a.update()
    a.adjust_rho(90, 100)
    assert a.rho == pytest.approx(5e-3) and a.init_rho == 1e-3
    state = a.state_dict()
    assert set(state) == {'rho', 'init_rho', 'u', 'z'} and set(state['u']) == set(names)
    snap_u = {n: a.u[n].clone() for n in names}
    snap_z = {n: a.z[n].clone() for n in names}
    a.update()                                             # moves u and z away from the snapshot
    assert any(not torch.equal(a.u[n], snap_u[n]) for n in names)
    for n in names:                                        # the snapshot is a copy, not a view of the live state
        assert torch.equal(state['u'][n], snap_u[n]) and torch.equal(state['z'][n], snap_z[n])

    b = ADMM(workloads.ParamBag(weights), 1e-3, hb(), fmt, 'cpu')   # a resumed run: U = 0, Z = W (admm.py:32-40)
    b.load_state_dict(state)
    assert b.rho == pytest.approx(5e-3) and b.init_rho == 1e-3
    for n in names:
        assert torch.equal(b.u[n], snap_u[n]) and torch.equal(b.z[n], snap_z[n])
    # the restored state drives the next update and the penalty exactly like the original object
    a.load_state_dict(state)
    a.update()
    b.update()
    for n in names:
        assert torch.equal(a.u[n], b.u[n]) and torch.equal(a.z[n], b.z[n])
    la = a.append_admm_loss(torch.zeros(()))
    lb = b.append_admm_loss(torch.zeros(()))
    assert float(la) == float(lb) and float(la) > 0.0


@pytest.mark.parametrize('epoch,epochs,expect', [(0, 100, 1.0), (85, 100, 1.0), (86, 100, 5.0), (9, 10, 5.0),
                                                 (8, 10, 1.0)])
def test_adjust_rho_threshold(emulated_backend, epoch, epochs, expect):
    from admm import ADMM
    weights = _subset(workloads.resnet32_weights(), ['layer1.0.conv1.weight'])
    a = ADMM(workloads.ParamBag(weights), 2e-3, hp_tables.tt_resnet32_3x(), 'tt', 'cpu')
    a.adjust_rho(epoch, epochs)
    assert a.rho == pytest.approx(expect * 2e-3)
    a.adjust_rho(epochs, epochs, factor=7)
    assert a.rho == pytest.approx(7 * 2e-3)


def test_ten2tt_cores_are_orthonormal_like_the_reference(emulated_backend):
    """ttd.py:21-25 hands out U[:, :r] of each SVD: every core but the last, reshaped to (r_i * s_i, r_{i+1}),
    has orthonormal columns -- also on steps that keep every singular triplet (where the projection plans of
    admm.py skip the eigensolve).  A tuple of ranks raises only when a clip would write into it (ttd.py:18-19)."""
    import ttd
    rng = np.random.RandomState(4)
    x = rng.randn(8, 6, 10).astype(np.float32)
    for ranks_in in ([1, 8, 10, 1], [1, 5, 7, 1]):
        ranks = list(ranks_in)
        cores = ttd.ten2tt(x, [8, 6, 10], ranks)
        ref = port.tt_svd(x, [8, 6, 10], list(ranks_in))
        assert ranks == [1, min(ranks_in[1], 8), min(ranks_in[2], 10), 1]
        for c, rc in zip(cores, ref):
            assert c.shape == rc.shape
        for c in cores[:-1]:
            m = c.reshape(-1, c.shape[-1]).astype(np.float64)
            assert np.max(np.abs(m.T @ m - np.eye(m.shape[1]))) <= 1e-5
        assert rel_fro(ttd.tt2ten(cores, x.shape), port.tt_contract(ref, x.shape)) <= 1e-5
    ttd.ten2tt(x, [8, 6, 10], (1, 5, 7, 1))                      # no clip needed: a tuple is fine
    with pytest.raises(TypeError):
        ttd.ten2tt(x, [8, 6, 10], (1, 9, 7, 1))                  # clip 9 -> 8 must write into the tuple


def _orth_reference(model, rho):
    """orthogonal.py:9-20 restated with torch ops (the oracle of the regulariser)."""
    total = torch.zeros((), dtype=torch.float64)
    for name, p in model.named_parameters():
        if any(k in name for k in ('first_kernel', 'last_kernel', 'first_factor', 'last_factor', 'left_kernel')):
            m = torch.squeeze(p).double()
            g = m @ m.t() if m.shape[0] < m.shape[1] else m.t() @ m
            total = total + 0.5 * rho * torch.norm(g - torch.eye(g.shape[0], dtype=torch.float64), p=2) ** 2
    return total


def test_orthogonal_regulariser_matches_reference_formula(emulated_backend):
    """`append_double_l2_loss` (orthogonal.py:9-20): value and gradient for wide, tall and 4-D (1x1 kernel) factors,
    unselected parameters untouched, AMP-style scaled backward."""
    import orthogonal
    g = torch.Generator().manual_seed(7)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.first_factor = torch.nn.Parameter(torch.randn(6, 20, generator=g) * 0.3)         # wide: F F^T
            self.last_factor = torch.nn.Parameter(torch.randn(40, 9, generator=g) * 0.3)          # tall: F^T F
            self.first_kernel = torch.nn.Parameter(torch.randn(5, 33, 1, 1, generator=g) * 0.3)   # TKConv2dC
            self.left_kernel = torch.nn.Parameter(torch.randn(37, 37, 1, 1, generator=g) * 0.2)   # square: F^T F
            self.core_kernel = torch.nn.Parameter(torch.randn(4, 4, 3, 3, generator=g))           # not selected

    net = Net()
    rho = 0.05
    base = torch.tensor(0.5, requires_grad=True)
    loss = orthogonal.append_double_l2_loss(net, base * 2.0, rho, 'cpu')
    ref = _orth_reference(net, rho)
    assert abs(float(loss.detach()) - (1.0 + float(ref))) <= 1e-5 * (1.0 + float(ref))
    (loss * 3.0).backward()
    assert abs(float(base.grad) - 6.0) < 1e-6
    grads = torch.autograd.grad(ref, [net.first_factor, net.last_factor, net.first_kernel, net.left_kernel])
    for p, gr in zip([net.first_factor, net.last_factor, net.first_kernel, net.left_kernel], grads):
        assert torch.allclose(p.grad.double(), 3.0 * gr.double(), rtol=1e-4, atol=1e-6)
    assert net.core_kernel.grad is None
    # a model without factor parameters: the loss is returned unchanged
    lin = torch.nn.Linear(3, 3)
    l0 = torch.zeros(())
    assert orthogonal.append_double_l2_loss(lin, l0, rho, 'cpu') is l0


def test_orthogonal_regulariser_against_reference_golden(emulated_backend):
    """Host-side wiring of `append_double_l2_loss` against the reference's own run (tests/golden/forward_modules.*)."""
    import json
    import os
    import orthogonal
    from helpers import GOLDEN
    with open(os.path.join(GOLDEN, 'forward_modules.json')) as f:
        case = [c for c in json.load(f) if c['kind'] == 'orth'][0]
    arr = np.load(os.path.join(GOLDEN, 'forward_modules.npz'))

    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            for n in case['params']:
                setattr(self, n, torch.nn.Parameter(torch.from_numpy(arr['orth|' + n]).clone()))

    toy = Toy()
    loss = orthogonal.append_double_l2_loss(toy, torch.zeros(()), case['rho'], 'cpu')
    assert abs(float(loss.detach()) - case['loss']) <= 1e-5 * abs(case['loss'])
    loss.backward()
    for n, p in toy.named_parameters():
        key = 'orth|grad|' + n
        if key in arr.files:
            assert np.allclose(p.grad.numpy(), arr[key], rtol=1e-4, atol=1e-6), n
        else:
            assert p.grad is None


@pytest.mark.parametrize('shape,ranks', [((192, 384), [40, 72]), ((192, 384), [72, 40]), ((16, 8, 3, 3), [12, 8]),
                                         ((8, 64, 1, 1), [8, 40]), ((64, 64, 3, 3), [32, 32]), ((6, 50), [9, 9])])
def test_tucker_rank_clip_matches_the_oracle_factor_widths(shape, ranks):
    """A truncated SVD returns at most min(shape) vectors, so HOOI may end with narrower factors than the requested
    ranks; TKLayer resolves that statically and must agree with the oracle's factors (admm.py:116,124)."""
    rng = np.random.default_rng(5)
    w = rng.standard_normal(shape).astype(np.float32)
    w3 = w.reshape(shape[0], shape[1], -1)
    _, factors = port.partial_tucker2(w3, list(ranks))
    layer = projector.TKLayer('w', shape, list(ranks))
    assert (layer.r0, layer.r1) == (factors[0].shape[1], factors[1].shape[1])


def test_wave_makespan_with_whole_gpu_problems():
    """Eigenproblems wider than the cluster solvers occupy the whole GPU and run after the clustered ones."""
    t_small = projector.wave_makespan_ms([512, 256])
    assert projector.wave_makespan_ms([640]) == pytest.approx(projector.eig_time_ms(640))
    assert projector.wave_makespan_ms([512, 256, 640]) == pytest.approx(t_small + projector.eig_time_ms(640))


def test_first_step_gram_reads_w_and_u_in_place(emulated_backend):
    """The Gram task of the first TT step points at the parameter and the dual tensor themselves (tta_gram_task.a / a2)
    whenever its rows are made of whole output channels; later steps, and first steps that are column Grams of a k x k
    convolution, read the plan's own buffers."""
    import tta_runtime as rt
    hp = hp_tables.tt_resnet50_general_3x()
    names = ['layer3.1.conv2.weight', 'layer3.0.conv1.weight']       # a 3 x 3 (first step truncates) and a 1 x 1 convolution
    w = {n: t for n, t in workloads.resnet50_weights(seed=0).items() if n in names}
    layers = [projector.TTLayer(n, tuple(w[n].shape), hp.tt_shapes[n], list(hp.ranks[n])) for n in names]
    plan = projector.TTProjectionPlan(layers, 'cpu')
    ws = [w[n].clone() for n in names]
    us = [torch.randn_like(t) * 0.01 for t in ws]
    zs = [torch.empty_like(t) for t in ws]
    plan.bind(ws, us, zs)
    seen = 0
    for wave in plan.waves:
        tab = wave['gram'].host
        for q, li in enumerate(wave['idx']):
            row = tab[q]
            if int(row['a']) == ws[li].data_ptr():
                assert int(row['a2']) == us[li].data_ptr()
                seen += 1
            else:
                assert int(row['a2']) == 0
    assert seen == len(names)          # exactly the first real step of every layer


@pytest.mark.parametrize('k,red', [(8, 40), (32, 73728), (480, 4608), (512, 1024), (2048, 18432), (100, 7)])
def test_gram_splits_bounds(k, red):
    n = projector.gram_splits(k, red)
    tiles = (k + 127) // 128
    assert 1 <= n <= 148
    assert n == 1 or n * 256 <= red + 255                       # at least 256 reduction indices per slice
    assert n * tiles * (tiles + 1) // 2 <= 148 + tiles * (tiles + 1) // 2      # about one CTA per SM
