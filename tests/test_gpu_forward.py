"""Decomposed-layer forward building blocks and drop-in layer modules on a B200.
Tolerance (BASELINE.json north_star): forward outputs within 1e-2 relative error in bf16."""
import numpy as np
import pytest
import torch

import tta_runtime as rt

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
FWD_TOL = 1e-2


def _rel(a, b):
    a = a.double()
    b = b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


@pytest.mark.parametrize('M,N,K', [(128, 64, 64), (256, 128, 128), (300, 320, 368), (1000, 1120, 320),
                                   (128, 800, 256), (77, 23, 24), (4096, 352, 1344), (130, 129, 72),
                                   (20000, 1120, 320), (5000, 384, 1536), (300, 264, 40)])
@pytest.mark.parametrize('out_f32', [False, True])
def test_tcgen05_gemm_matches_torch(M, N, K, out_f32):
    g = torch.Generator(device='cpu').manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g).to(DEV).to(torch.bfloat16)
    b = torch.randn(N, K, generator=g).to(DEV).to(torch.bfloat16)
    bias = torch.randn(N, generator=g).to(DEV)
    c = torch.full((M, N), 7.0, device=DEV, dtype=torch.float32 if out_f32 else torch.bfloat16)
    rt.gemm_bf16_tc(a, b, c, M, N, K, bias=bias)
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t() + bias
    tol = 2e-6 * K if out_f32 else 6e-3          # fp32 out: accumulation order only; bf16 out: output rounding
    assert _rel(c.float(), ref) <= tol, (_rel(c.float(), ref), tol)


def test_tcgen05_gemm_strided_operands():
    g = torch.Generator(device='cpu').manual_seed(5)
    M, N, K = 512, 192, 96
    a_full = torch.randn(M, K + 40, generator=g).to(DEV).to(torch.bfloat16)
    b_full = torch.randn(N, K + 8, generator=g).to(DEV).to(torch.bfloat16)
    c_full = torch.zeros(M, N + 16, device=DEV, dtype=torch.bfloat16)
    rt.gemm_bf16_tc(a_full, b_full, c_full, M, N, K, lda=K + 40, ldb=K + 8, ldc=N + 16)
    ref = a_full[:, :K].float() @ b_full[:, :K].float().t()
    assert _rel(c_full[:, :N].float(), ref) <= 6e-3
    assert float(c_full[:, N:].abs().max()) == 0.0


def test_small_gemm_and_cast():
    g = torch.Generator(device='cpu').manual_seed(6)
    T, m1, r1, m0 = 50, 32, 35, 36
    z = torch.randn(T * m1, r1, generator=g).to(DEV).to(torch.bfloat16)
    core = torch.randn(m0, r1, generator=g).to(DEV).to(torch.bfloat16)
    bias = torch.randn(m0 * m1, generator=g).to(DEV)
    y = torch.zeros(T, m0 * m1, device=DEV)
    # y[t, o0*m1 + o1] = sum_a z[(t,o1), a] core[o0, a] + bias[o0*m1 + o1]
    rt.small_gemm(z, core, y, T * m1, m0, r1, m_inner=m1, s_outer=m0 * m1, s_inner=1, s_col=m1, bias=bias,
                  bias_inner=1, bias_col=m1)
    ref = torch.einsum('tia,oa->toi', z.float().view(T, m1, r1), core.float()).reshape(T, m0 * m1) + bias
    assert _rel(y, ref) <= 1e-5
    x = torch.randn(1000 * 16, 24, generator=g).to(DEV)
    h = torch.randn(23, 24, generator=g).to(DEV).to(torch.bfloat16)
    u = torch.zeros(1000 * 16, 23, device=DEV, dtype=torch.bfloat16)
    rt.small_gemm(x, h, u, x.shape[0], 23, 24)
    assert _rel(u.float(), x @ h.float().t()) <= 6e-3
    xb = torch.empty(x.numel(), device=DEV, dtype=torch.bfloat16)
    rt.cast_bf16(x.reshape(-1), xb)
    assert torch.equal(xb, x.reshape(-1).to(torch.bfloat16))


def _deit_hp():
    import hp_tables
    return hp_tables.tt_deit_small_2x()


@pytest.mark.parametrize('name,fin,fout', [('blocks.0.attn.qkv.weight', 384, 1152), ('blocks.1.attn.proj.weight', 384, 384),
                                            ('blocks.0.mlp.fc1.weight', 384, 1536), ('blocks.1.mlp.fc2.weight', 1536, 384)])
def test_ttlinear_m_forward_matches_dense(name, fin, fout):
    """SURVEY 3.4 identity: TTLinearM(dense_w=W)(x) == F.linear(x, Proj_TT(W)); bf16 tolerance 1e-2."""
    import TTLinear
    from oracle import port
    hp = _deit_hp()
    g = torch.Generator(device='cpu').manual_seed(3)
    w = torch.randn(fout, fin, generator=g) * 0.02
    b = torch.randn(fout, generator=g) * 0.1
    layer = TTLinear.TTLinearM(fin, fout, bias=True, hp_dict=hp.fresh(), name=name, dense_w=w, dense_b=b).to(DEV)
    x = torch.randn(4, 197, fin, generator=g).to(DEV)
    with torch.no_grad():
        y = layer(x)
    z = torch.from_numpy(port.project_linear_tt(w.numpy(), hp.tt_shapes[name], list(hp.ranks[name]))).to(DEV)
    ref = torch.nn.functional.linear(x, z, b.to(DEV))
    assert y.shape == ref.shape
    assert _rel(y, ref) <= FWD_TOL, _rel(y, ref)
    assert layer._fused2 is True                    # the DeiT-small tables take the fused two-factor tcgen05 kernel
    with torch.no_grad():
        yb = layer(x.to(torch.bfloat16))            # bf16 activations in -> bf16 out, no cast pass
    assert yb.dtype == torch.bfloat16 and _rel(yb.float(), ref) <= FWD_TOL
    layer._fused2 = False                           # the four-step chain (skinny contractions + tcgen05 GEMMs)
    with torch.no_grad():
        yc = layer(x)
    assert _rel(yc, ref) <= FWD_TOL, _rel(yc, ref)
    layer._fused2 = True
    # training path (autograd) agrees with the fused path and produces gradients for every core
    xg = x.clone().requires_grad_(True)
    yt = layer(xg)
    assert _rel(yt.detach(), ref) <= 2e-3           # torch conv runs TF32 by default
    yt.square().mean().backward()
    assert all(c.grad is not None for c in layer.tt_cores) and xg.grad is not None
    layer_r = TTLinear.TTLinearR(fin, fout, bias=True, hp_dict=hp.fresh(), name=name, dense_w=w, dense_b=b).to(DEV)
    with torch.no_grad():
        yr = layer_r(x)
    assert _rel(yr, ref) <= FWD_TOL
    assert set(dict(layer.named_parameters())) == {'tt_cores.0', 'tt_cores.1', 'tt_cores.2', 'tt_cores.3', 'bias'}


R32_CASES = [('layer1.0.conv1.weight', 16, 16, 1, 32), ('layer2.0.conv1.weight', 16, 32, 2, 32),
             ('layer2.1.conv2.weight', 32, 32, 1, 16), ('layer3.0.conv1.weight', 32, 64, 2, 16),
             ('layer3.2.conv2.weight', 64, 64, 1, 8)]


@pytest.mark.parametrize('name,cin,cout,stride,hw', R32_CASES)
def test_ttconv2d_m_forward_matches_dense(name, cin, cout, stride, hw):
    """TTConv2dM(dense_w=W)(x) == F.conv2d(x, Proj_TT(W)), bf16 tolerance 1e-2."""
    import hp_tables
    import TTConv
    from oracle import port
    hp = hp_tables.tt_resnet32_3x()
    g = torch.Generator(device='cpu').manual_seed(11)
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    layer = TTConv.TTConv2dM(cin, cout, 3, stride=stride, padding=1, bias=True, hp_dict=hp.fresh(), name=name,
                             dense_w=w, dense_b=b).to(DEV)
    x = torch.randn(8, cin, hw, hw, generator=g).to(DEV)
    with torch.no_grad():
        y = layer(x)
    z = torch.from_numpy(port.project_conv_tt(w.numpy(), hp.tt_shapes[name], list(hp.ranks[name]))).to(DEV)
    ref = torch.nn.functional.conv2d(x, z, b.to(DEV), stride=stride, padding=1)
    assert y.shape == ref.shape
    assert _rel(y, ref) <= FWD_TOL, _rel(y, ref)
    xg = x.clone().requires_grad_(True)
    yt = layer(xg)                               # autograd path
    assert _rel(yt.detach(), ref) <= 2e-3           # torch conv runs TF32 by default
    yt.mean().backward()
    assert layer.core_kernel.grad is not None and xg.grad is not None
    names = set(dict(layer.named_parameters()))
    assert 'core_kernel' in names and 'bias' in names and 'in_tt_cores.0' in names and 'out_tt_cores.0' in names


def test_ttconv2d_r_and_value_errors():
    import hp_tables
    import TTConv
    hp = hp_tables.tt_resnet32_3x()
    name = 'layer2.1.conv2.weight'
    g = torch.Generator(device='cpu').manual_seed(12)
    w = torch.randn(32, 32, 3, 3, generator=g) * 0.1
    layer = TTConv.TTConv2dR(32, 32, 3, padding=1, bias=True, hp_dict=hp.fresh(), name=name, dense_w=w,
                             dense_b=torch.zeros(32)).to(DEV)
    x = torch.randn(4, 32, 16, 16, generator=g).to(DEV)
    with torch.no_grad():
        y = layer(x)
        ref = torch.nn.functional.conv2d(x, layer._recover_weight(), layer.bias, padding=1)
    assert _rel(y, ref) <= FWD_TOL
    assert tuple(layer.conv_core.shape) == (16, 9, 16)
    with pytest.raises(ValueError):
        TTConv.TTConv2dM(32, 32, 3, groups=2, hp_dict=hp.fresh(), name=name)
    with pytest.raises(ValueError):
        TTConv.TTConv2dR(32, 32, 3, padding_mode='reflect', hp_dict=hp.fresh(), name=name)


@pytest.mark.parametrize('variant', ['C', 'M', 'R'])
@pytest.mark.parametrize('name,cin,cout,stride,hw', [R32_CASES[1], R32_CASES[4]])
def test_tkconv2d_forward_matches_dense(variant, name, cin, cout, stride, hw):
    """TKConv2d{C,M,R}(dense_w=W)(x) == F.conv2d(x, Proj_TK(W)) (restated-tensorly oracle; unpinned)."""
    import hp_tables
    import TKConv
    from oracle import port
    hp = hp_tables.tk_resnet32('3')
    g = torch.Generator(device='cpu').manual_seed(13)
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    cls = getattr(TKConv, 'TKConv2d' + variant)
    layer = cls(cin, cout, 3, stride=stride, padding=1, bias=True, hp_dict=hp, name=name, dense_w=w, dense_b=b).to(DEV)
    x = torch.randn(8, cin, hw, hw, generator=g).to(DEV)
    with torch.no_grad():
        y = layer(x)
    z = torch.from_numpy(port.project_tk(w.numpy(), hp.ranks[name])).to(DEV)
    ref = torch.nn.functional.conv2d(x, z, b.to(DEV), stride=stride, padding=1)
    assert _rel(y, ref) <= FWD_TOL, _rel(y, ref)
    xg = x.clone().requires_grad_(True)
    assert _rel(layer(xg).detach(), ref) <= 2e-3


@pytest.mark.parametrize('variant', ['M', 'R'])
def test_tklinear_forward_matches_dense(variant):
    import hp_tables
    import TKLinear
    from oracle import port
    g = torch.Generator(device='cpu').manual_seed(14)
    w = torch.randn(192, 384, generator=g) * 0.05
    b = torch.randn(192, generator=g) * 0.1
    hp = hp_tables.HpTable('tk_lin', {'fc.weight': [40, 72]})
    layer = getattr(TKLinear, 'TKLinear' + variant)(384, 192, bias=True, hp_dict=hp, name='fc.weight', dense_w=w,
                                                    dense_b=b).to(DEV)
    x = torch.randn(3, 50, 384, generator=g).to(DEV)
    with torch.no_grad():
        y = layer(x)
    z = torch.from_numpy(port.project_tk(w.numpy(), [40, 72])).to(DEV)
    ref = torch.nn.functional.linear(x, z, b.to(DEV))
    assert _rel(y, ref) <= FWD_TOL, _rel(y, ref)
    # r_in = 72 exceeds the width of the partially projected unfolding (40): a truncated SVD hands back 40 vectors, so the
    # refreshed factor has 40 columns (oracle/port.partial_tucker2; TKLayer resolves the clip statically)
    assert tuple(layer.first_factor.shape) == (40, 384) and tuple(layer.last_factor.shape) == (192, 40)
    assert tuple(layer.core_tensor.shape) == (40, 40)
    if variant == 'M':                     # fused training path: same gradients as the torch op chain (bf16 rounding)
        grads = []
        for fused in (False, True):
            layer.fused_training = fused
            layer.zero_grad()
            xg = x.clone().requires_grad_(True)
            layer(xg).square().sum().backward()
            grads.append([xg.grad.clone(), layer.first_factor.grad.clone(), layer.core_tensor.grad.clone(),
                          layer.last_factor.grad.clone(), layer.bias.grad.clone()])
        for a, b_ in zip(grads[1], grads[0]):
            assert _rel(a, b_) <= 3e-2, _rel(a, b_)


def test_tklinear_head_with_odd_output_width():
    """A 10-class head: out_features is not a multiple of 4, the fused kernel's output pitch is padded and sliced."""
    import hp_tables
    import TKLinear
    g = torch.Generator(device='cpu').manual_seed(21)
    w = torch.randn(10, 64, generator=g) * 0.1
    b = torch.randn(10, generator=g) * 0.1
    hp = hp_tables.HpTable('tk_head', {'fc.weight': [8, 24]})
    layer = TKLinear.TKLinearM(64, 10, bias=True, hp_dict=hp, name='fc.weight', dense_w=w, dense_b=b).to(DEV)
    x = torch.randn(33, 64, generator=g).to(DEV)
    with torch.no_grad():
        y = layer(x)
        ref = layer._forward_torch(x) if hasattr(layer, '_forward_torch') else torch.nn.functional.linear(
            torch.nn.functional.linear(torch.nn.functional.linear(x, layer.first_factor), layer.core_tensor), layer.last_factor,
            layer.bias)
    assert tuple(y.shape) == (33, 10)
    assert _rel(y, ref) <= FWD_TOL, _rel(y, ref)


@pytest.mark.parametrize('B,Cin,H,W,Ra,Rb,Cout,KS,stride,pad', [
    (3, 16, 32, 32, 16, 16, 16, 3, 1, 1),      # ttm_resnet32 layer1
    (2, 16, 32, 32, 16, 32, 32, 3, 2, 1),      # stride-2 transition
    (4, 64, 8, 8, 27, 29, 64, 3, 1, 1),        # layer3 ranks, ranks not multiples of 4
    (2, 7, 13, 9, 5, 6, 11, 3, 1, 1),          # ragged everything
    (2, 12, 10, 10, 8, 8, 20, 1, 1, 0),        # 1x1 core
    (1, 32, 19, 17, 10, 12, 24, 3, 2, 0),      # stride 2, no padding
    (1, 64, 56, 56, 40, 40, 64, 3, 1, 1),      # bigger image: several tiles per image, tile shrink
])
def test_ttconv_fused_kernel_matches_conv_chain(B, Cin, H, W, Ra, Rb, Cout, KS, stride, pad):
    """tta_ttconv_fused_fwd == conv1x1 -> conv kxk -> conv1x1 (+bias) of torch in fp32 (tolerance: fp32 rounding)."""
    import torch.nn.functional as F
    g = torch.Generator(device='cpu').manual_seed(B * 131 + Cin * 17 + H)
    x = torch.randn(B, Cin, H, W, generator=g).to(DEV)
    a_in = (torch.randn(Ra, Cin, generator=g) / Cin ** 0.5).to(DEV)
    kern = (torch.randn(Rb, Ra, KS, KS, generator=g) / (Ra * KS * KS) ** 0.5).to(DEV)
    a_out = (torch.randn(Cout, Rb, generator=g) / Rb ** 0.5).to(DEV)
    bias = torch.randn(Cout, generator=g).to(DEV)
    Ho, Wo = (H + 2 * pad - KS) // stride + 1, (W + 2 * pad - KS) // stride + 1
    for b in (bias, None):
        y = torch.full((B, Cout, Ho, Wo), float('nan'), device=DEV)
        rt.ttconv_fused_fwd(x, a_in, kern, a_out, b, y, B, Cin, H, W, Ra, Rb, Cout, KS, stride, pad)
        torch.cuda.synchronize()
        with torch.backends.cudnn.flags(allow_tf32=False):
            ref = F.conv2d(F.conv2d(F.conv2d(x.double(), a_in.double()[:, :, None, None]), kern.double(), None, stride, pad),
                           a_out.double()[:, :, None, None], b.double() if b is not None else None)
        assert torch.isfinite(y).all()
        assert _rel(y, ref) <= 2e-6, _rel(y, ref)


@pytest.mark.parametrize('B,Cin,H,W,Ra,Rb,Cout', [
    (3, 16, 32, 32, 16, 32, 32),       # ttm_resnet32 stride-2 transition
    (2, 32, 16, 16, 29, 31, 64),
    (2, 9, 13, 10, 5, 6, 11),          # odd sizes
])
def test_ttconv_tensor_core_kernel_stride2(B, Cin, H, W, Ra, Rb, Cout):
    import torch.nn.functional as F
    assert rt.ttconv_tc_supported(Cin, Ra, Rb, Cout, 3, 2, 1)
    g = torch.Generator(device='cpu').manual_seed(B * 7 + Cin)
    x = torch.randn(B, Cin, H, W, generator=g).to(DEV)
    a_in = (torch.randn(Ra, Cin, generator=g) / Cin ** 0.5).to(DEV)
    kern = (torch.randn(Rb, Ra, 3, 3, generator=g) / (Ra * 9) ** 0.5).to(DEV)
    a_out = (torch.randn(Cout, Rb, generator=g) / Rb ** 0.5).to(DEV)
    bias = torch.randn(Cout, generator=g).to(DEV)
    ref = F.conv2d(F.conv2d(F.conv2d(x.double(), a_in.double()[:, :, None, None]), kern.double(), None, 2, 1),
                   a_out.double()[:, :, None, None], bias.double())
    y = torch.full(tuple(ref.shape), float('nan'), device=DEV)
    rt.ttconv_tc_fwd(x, rt.ttconv_tc_pack(a_in, kern, a_out, bias), y, B, Cin, H, W, Ra, Rb, Cout, 3, 2, 1)
    torch.cuda.synchronize()
    assert torch.isfinite(y).all()
    assert _rel(y, ref) <= 1e-2, _rel(y, ref)


@pytest.mark.parametrize('B,Cin,H,W,Ra,Rb,Cout', [
    (3, 16, 32, 32, 16, 16, 16),       # ttm_resnet32 layer1
    (5, 32, 16, 16, 20, 24, 32),       # layer2, ranks not multiples of 16
    (4, 64, 8, 8, 27, 29, 64),         # layer3 ranks
    (2, 7, 13, 9, 5, 6, 11),           # ragged everything
    (1, 64, 56, 56, 40, 40, 64),       # bigger image: the taps reach across several 128-position tiles
    (1, 16, 3, 3, 16, 16, 16),         # one chunk, mostly padding
    (130, 16, 32, 32, 16, 16, 16),     # chunks of several tiles, many images
])
def test_ttconv_tensor_core_kernel_matches_conv_chain(B, Cin, H, W, Ra, Rb, Cout):
    """tta_ttconv_tc_fwd (bf16 tcgen05, intermediates in shared memory) == conv1x1 -> conv3x3 -> conv1x1 (+bias) of torch
    in fp64, within the 1e-2 bar of the decomposed-layer forwards (three bf16 roundings: ~4e-3)."""
    import torch.nn.functional as F
    assert rt.ttconv_tc_supported(Cin, Ra, Rb, Cout, 3, 1, 1)
    assert not rt.ttconv_tc_supported(Cin, Ra, Rb, Cout, 3, 1, 0) and not rt.ttconv_tc_supported(128, Ra, Rb, Cout, 3, 1, 1)
    g = torch.Generator(device='cpu').manual_seed(B * 131 + Cin * 17 + H)
    x = torch.randn(B, Cin, H, W, generator=g).to(DEV)
    a_in = (torch.randn(Ra, Cin, generator=g) / Cin ** 0.5).to(DEV)
    kern = (torch.randn(Rb, Ra, 3, 3, generator=g) / (Ra * 9) ** 0.5).to(DEV)
    a_out = (torch.randn(Cout, Rb, generator=g) / Rb ** 0.5).to(DEV)
    bias = torch.randn(Cout, generator=g).to(DEV)
    for b in (bias, None):
        y = torch.full((B, Cout, H, W), float('nan'), device=DEV)
        blob = rt.ttconv_tc_pack(a_in, kern, a_out, b)
        rt.ttconv_tc_fwd(x, blob, y, B, Cin, H, W, Ra, Rb, Cout, 3, 1, 1)
        torch.cuda.synchronize()
        ref = F.conv2d(F.conv2d(F.conv2d(x.double(), a_in.double()[:, :, None, None]), kern.double(), None, 1, 1),
                       a_out.double()[:, :, None, None], b.double() if b is not None else None)
        assert torch.isfinite(y).all()
        assert _rel(y, ref) <= 1e-2, _rel(y, ref)
        # every image and every border pixel individually (a wrong tap offset or a leak across the padding shows here)
        err = (y.double() - ref).flatten(1).norm(dim=1) / ref.flatten(1).norm(dim=1)
        assert float(err.max()) <= 1.5e-2, float(err.max())
        edge = torch.cat([(y.double() - ref)[..., 0, :].flatten(), (y.double() - ref)[..., -1, :].flatten(),
                          (y.double() - ref)[..., :, 0].flatten(), (y.double() - ref)[..., :, -1].flatten()])
        edge_ref = torch.cat([ref[..., 0, :].flatten(), ref[..., -1, :].flatten(), ref[..., :, 0].flatten(), ref[..., :, -1].flatten()])
        assert float(edge.norm() / edge_ref.norm()) <= 1.5e-2


@pytest.mark.parametrize('M,K1,N1,N2', [(128, 64, 64, 64), (256, 384, 320, 1152), (1000, 384, 256, 384),
                                        (300, 1536, 320, 384), (77, 72, 40, 50), (4096, 384, 320, 1536),
                                        (513, 200, 23, 1000), (129, 384, 384, 96), (20000, 384, 256, 1152),
                                        (200, 64, 32, 5), (300, 128, 48, 37)])
@pytest.mark.parametrize('out_f32', [False, True])
def test_lowrank2_fused_forward_matches_torch(M, K1, N1, N2, out_f32):
    """y = bf16(x W1^T) W2^T + bias in one TMA-fed tcgen05 kernel (persistent over row tiles; ragged M, K1, N1, N2)."""
    g = torch.Generator(device='cpu').manual_seed(M + 3 * K1 + 5 * N1 + 7 * N2)
    x = torch.randn(M, K1, generator=g).to(DEV).to(torch.bfloat16)
    w1 = (torch.randn(N1, K1, generator=g) / K1 ** 0.5).to(DEV).to(torch.bfloat16)
    ld2 = (N1 + 7) // 8 * 8                       # leading dimensions must be multiples of 8 (16-byte rows for TMA)
    w2_full = torch.full((N2, ld2), float('nan'), dtype=torch.bfloat16, device=DEV)
    w2_full[:, :N1] = (torch.randn(N2, N1, generator=g) / N1 ** 0.5).to(DEV).to(torch.bfloat16)
    w2 = w2_full[:, :N1]
    bias = torch.randn(N2, generator=g).to(DEV)
    ldy = (N2 + 7) // 8 * 8 + 8
    y = torch.full((M, ldy), 7.0, device=DEV, dtype=torch.float32 if out_f32 else torch.bfloat16)
    rt.lowrank2_fwd(x, w1, w2_full, bias, y, M, K1, N1, N2, ld2=ld2, ldy=ldy)
    torch.cuda.synchronize()
    v = (x.float() @ w1.float().t()).to(torch.bfloat16).float()
    ref = v @ w2.float().t() + bias
    tol = 2e-3 if out_f32 else 6e-3      # bf16 rounding of the intermediate can differ by one ulp at ties
    assert _rel(y[:, :N2].float(), ref) <= tol, _rel(y[:, :N2].float(), ref)
    gran = 4 if out_f32 else 8                                        # TMA stores work in 16-byte granules
    n2p = (N2 + gran - 1) // gran * gran
    assert float((y[:, n2p:].float() - 7.0).abs().max()) == 0.0      # nothing written past the granule that holds N2


@pytest.mark.parametrize('M,N,K,lda,ldb', [(128, 128, 64, 128, 128), (1152, 256, 50432, 1152, 256), (256, 384, 3000, 256, 384),
                                           (40, 72, 777, 48, 72), (320, 1536, 4097, 320, 1536), (129, 65, 130, 136, 72)])
def test_gemm_bf16_tn_weight_gradient_kernel(M, N, K, lda, ldb):
    """tta_gemm_bf16_tn: c = a^T b with the reduction index slow in both operands (TMA boxes as MN-major tcgen05 operands,
    split-K) against torch in fp64; bf16 inputs, fp32 accumulation -> 1e-5 grade of the fp32 result."""
    g = torch.Generator(device='cpu').manual_seed(M + N + K)
    a = torch.randn(K, lda, generator=g).to(torch.bfloat16).to(DEV)
    b = torch.randn(K, ldb, generator=g).to(torch.bfloat16).to(DEV)
    out = torch.full((M, N), float('nan'), device=DEV)
    rt.gemm_bf16_tn(a, b, out, M, N, K, lda=lda, ldb=ldb, ldc=N)
    torch.cuda.synchronize()
    ref = a[:, :M].double().t() @ b[:, :N].double()
    assert torch.isfinite(out).all()
    assert _rel(out, ref) <= 2e-5, _rel(out, ref)
    out2 = torch.empty_like(out)
    rt.gemm_bf16_tn(a, b, out2, M, N, K, lda=lda, ldb=ldb, ldc=N)
    assert torch.equal(out, out2)                      # fixed summation order: bit-reproducible


def test_tensor_core_kernels_stay_inside_their_outputs():
    """Sentinel guards around the outputs of the round-2 tensor-core kernels (no memcheck tool on the GPU pool): nothing
    outside the output tensor may change."""
    g = torch.Generator(device='cpu').manual_seed(77)
    guard = 4096
    # fused convolution: ragged sizes, stride 1 and 2
    for (B, Cin, H, W, Ra, Rb, Cout, stride) in ((2, 7, 13, 9, 5, 6, 11, 1), (3, 16, 15, 15, 16, 24, 32, 2), (1, 64, 8, 8, 27, 29, 64, 1)):
        x = torch.randn(B, Cin, H, W, generator=g).to(DEV)
        blob = rt.ttconv_tc_pack(torch.randn(Ra, Cin, generator=g).to(DEV), torch.randn(Rb, Ra, 3, 3, generator=g).to(DEV),
                                 torch.randn(Cout, Rb, generator=g).to(DEV), None)
        Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        n = B * Cout * Ho * Wo
        buf = torch.full((n + 2 * guard,), 12345.0, device=DEV)
        y = buf[guard:guard + n].view(B, Cout, Ho, Wo)
        rt.ttconv_tc_fwd(x, blob, y, B, Cin, H, W, Ra, Rb, Cout, 3, stride, 1)
        torch.cuda.synchronize()
        assert bool((buf[:guard] == 12345.0).all()) and bool((buf[guard + n:] == 12345.0).all())
        assert not bool((y == 12345.0).any())
    # weight-gradient GEMM: output pitch wider than N
    for (M, N, K) in ((40, 72, 777), (129, 65, 130)):
        a = torch.randn(K, (M + 7) // 8 * 8, generator=g).to(torch.bfloat16).to(DEV)
        b = torch.randn(K, (N + 7) // 8 * 8, generator=g).to(torch.bfloat16).to(DEV)
        ldc = N + 5
        buf = torch.full((guard + M * ldc + guard,), 12345.0, device=DEV)
        out = buf[guard:guard + M * ldc].view(M, ldc)
        rt.gemm_bf16_tn(a, b, out, M, N, K, lda=a.shape[1], ldb=b.shape[1], ldc=ldc)
        torch.cuda.synchronize()
        assert bool((buf[:guard] == 12345.0).all()) and bool((buf[guard + M * ldc:] == 12345.0).all())
        assert bool((out[:, N:] == 12345.0).all()) and not bool((out[:, :N] == 12345.0).any())


def test_lowrank2_rejects_bad_arguments():
    x = torch.zeros(128, 64, device=DEV, dtype=torch.bfloat16)
    w1 = torch.zeros(400, 64, device=DEV, dtype=torch.bfloat16)
    w2 = torch.zeros(64, 400, device=DEV, dtype=torch.bfloat16)
    y = torch.zeros(128, 64, device=DEV)
    with pytest.raises(rt.TtaError):
        rt.lowrank2_fwd(x, w1, w2, None, y, 128, 64, 400, 64)          # inner width beyond the TMEM budget
    with pytest.raises(rt.TtaError):
        rt.lowrank2_fwd(x, w1, w2, None, y, 128, 64, 64, 64, ldx=60)    # leading dimension not a multiple of 8


@pytest.mark.parametrize('name,fin,fout', [('blocks.0.attn.qkv.weight', 384, 1152), ('blocks.1.mlp.fc2.weight', 1536, 384)])
def test_ttlinear_fused_training_path(name, fin, fout):
    """forward + dX on the fused two-factor kernel, core gradients by the chain rule through the
    folded factors; compared with the fp32 torch op chain (bf16 operand rounding: 2e-2 on gradients)."""
    import TTLinear
    hp = _deit_hp()
    torch.manual_seed(11)
    layer = TTLinear.TTLinearM(fin, fout, bias=True, hp_dict=hp.fresh(), name=name).to(DEV)
    with torch.no_grad():
        layer.bias.normal_(0, 0.1)
    x = torch.randn(2, 197, fin, device=DEV)
    gy = torch.randn(2, 197, fout, device=DEV)

    def run(fused):
        layer.fused_training = fused
        layer.zero_grad()
        xg = x.clone().requires_grad_(True)
        y = layer(xg)
        (y * gy).sum().backward()
        return y.detach(), xg.grad.clone(), [c.grad.clone() for c in layer.tt_cores], layer.bias.grad.clone()

    y0, dx0, dc0, db0 = run(False)
    y1, dx1, dc1, db1 = run(True)
    assert _rel(y1, y0) <= FWD_TOL and _rel(dx1, dx0) <= 2e-2
    for a, b in zip(dc1, dc0):
        assert _rel(a, b) <= 2e-2, _rel(a, b)
    assert _rel(db1, db0) <= 1e-5
    with torch.autocast('cuda', dtype=torch.bfloat16):           # autocast selects the fused path by itself
        layer.fused_training = False
        layer.zero_grad()
        xg = x.clone().requires_grad_(True)
        y2 = layer(xg)
    (y2.float() * gy).sum().backward()
    assert _rel(y2.detach().float(), y0) <= FWD_TOL and _rel(xg.grad, dx0) <= 2e-2


# ---- outputs of the UNMODIFIED reference modules (tests/golden/forward_modules.*, oracle/gen_golden_forward.py) ----
def _forward_cases():
    import json
    import os
    from helpers import GOLDEN
    with open(os.path.join(GOLDEN, 'forward_modules.json')) as f:
        return json.load(f)


@pytest.mark.parametrize('case', [c for c in _forward_cases() if c['kind'] != 'orth'],
                         ids=lambda c: '{}-{}'.format(c['module'], c['key']))
def test_modules_match_reference_module_outputs(case):
    """Drop-in module built from the same dense_w / dense_b through the same constructor, same input: output within
    1e-2 relative (bf16 kernels) of what the reference's own module returned, same state-dict names and shapes.
    Pins TTConv2dR's untransposed decomposition (TTConv.py:284-288,321) against the reference itself; the TK*
    cases ran on the restated tensorly (pinned: false)."""
    import os
    import SVDConv
    import TKConv
    import TKLinear
    import TTConv
    import TTLinear
    import hp_tables
    from helpers import GOLDEN
    arr = np.load(os.path.join(GOLDEN, 'forward_modules.npz'))
    src, key = case['inputs'], case['key']
    w, x, y_ref = (torch.from_numpy(arr[src + '|w']), torch.from_numpy(arr[src + '|x']).to(DEV),
                   torch.from_numpy(arr[key + '|y']).to(DEV))
    b = torch.from_numpy(arr[src + '|b']) if case['bias'] else None
    name = 'layer.weight'
    ranks = list(case['ranks']) if not isinstance(case['ranks'], int) else case['ranks']
    hp = hp_tables.HpTable('golden', {name: ranks}, {name: list(case['tt_shapes'])} if case['tt_shapes'] else None)
    mod = {'TTConv2dM': TTConv, 'TTConv2dR': TTConv, 'TTLinearM': TTLinear, 'TTLinearR': TTLinear,
           'TKConv2dC': TKConv, 'TKConv2dM': TKConv, 'TKConv2dR': TKConv, 'TKLinearM': TKLinear,
           'TKLinearR': TKLinear, 'SVDConv2dR': SVDConv, 'SVDConv2dC': SVDConv, 'SVDConv2dM': SVDConv}[case['module']]
    cls = getattr(mod, case['module'])
    if case['kind'] == 'conv':
        layer = cls(case['in'], case['out'], case['kernel'], stride=case['stride'], padding=case['padding'],
                    bias=case['bias'], hp_dict=hp, name=name, dense_w=w, dense_b=b).to(DEV)
    else:
        layer = cls(case['in'], case['out'], bias=case['bias'], hp_dict=hp, name=name, dense_w=w, dense_b=b).to(DEV)
    shapes = {n: list(p.shape) for n, p in layer.state_dict().items()}
    assert shapes == case['state_shapes']                        # checkpoint compatibility with the reference
    after = hp.ranks[name]
    assert (after if isinstance(after, int) else [int(v) for v in after]) == case['ranks_after']
    with torch.no_grad():
        y = layer(x)
    assert y.shape == y_ref.shape
    assert _rel(y.float(), y_ref) <= FWD_TOL, (case['module'], _rel(y.float(), y_ref))
    # the autograd path (what fine-tuning runs) gives the same function
    xg = x.clone().requires_grad_(True)
    yg = layer(xg)
    assert _rel(yg.detach().float(), y_ref) <= FWD_TOL
    yg.float().pow(2).sum().backward()
    assert xg.grad is not None and all(p.grad is not None for p in layer.parameters())


def test_orthogonal_regulariser_matches_reference_run():
    """`append_double_l2_loss` on the GPU kernels (csrc/orth.cu) against the value and the gradients the reference's
    orthogonal.py produced on the same seeded factors (tests/golden/forward_modules.*, key `orth`)."""
    import os
    import orthogonal
    from helpers import GOLDEN
    case = [c for c in _forward_cases() if c['kind'] == 'orth'][0]
    arr = np.load(os.path.join(GOLDEN, 'forward_modules.npz'))

    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            for n in case['params']:
                setattr(self, n, torch.nn.Parameter(torch.from_numpy(arr['orth|' + n]).clone()))

    toy = Toy().to(DEV)
    base = torch.zeros((), device=DEV, requires_grad=True)
    loss = orthogonal.append_double_l2_loss(toy, base + 0.0, case['rho'], DEV)
    assert abs(float(loss.detach()) - case['loss']) <= 1e-5 * abs(case['loss'])
    (loss * 4.0).backward()                                   # a GradScaler-style scaled loss (engines.py:315)
    for n, p in toy.named_parameters():
        key = 'orth|grad|' + n
        if key in arr.files:
            assert _rel(p.grad.cpu(), 4.0 * torch.from_numpy(arr[key])) <= 1e-5, n
        else:
            assert p.grad is None, n
