"""Shared machinery of the decomposed-layer forwards (TTConv / TTLinear / TKConv / TKLinear drop-ins).

Inference path (no autograd): activations are kept row-major over "rows" (tokens or pixels) in bf16,
every contraction of a TT / Tucker chain is one kernel of libtta.so:

  * big contractions (K or N > 96): `tta_gemm_bf16_tc` -- tcgen05 + TMEM tensor-core GEMM;
  * skinny contractions (outer TT cores, K, N <= 96): `tta_small_gemm` (HBM-bound, one thread per row,
    output written through a split-row address map so chains end in their final layout);
  * k x k core convolution: `tta_im2col_bf16` + tensor-core GEMM.

Weights are re-packed to bf16 (K padded to a multiple of 8, TT cores permuted so each chain step is a
plain K-major GEMM) once and cached until a parameter's version counter changes.

Fused kernels on top of that: `tta_lowrank2_fwd` (two-factor linear layers, intermediate in TMEM), `tta_ttconv_tc_fwd`
(bf16 tcgen05 fused 1x1 -> 3x3 -> 1x1 convolution, weights packed once per weight change: FoldedConv.blob) with
`tta_ttconv_fused_fwd` (fp32 CUDA cores) for the geometries the tensor-core kernel does not serve.

Training path (autograd enabled and something requires grad): the linear layers go through `LowRank2Fn` when
`layer.fused_training` is set or autocast is active -- forward and dX on the fused two-factor kernel, V / dV on the
TMA GEMM, the weight gradients dW1 / dW2 on `tta_gemm_bf16_tn` (csrc/gemm_tn.cu); otherwise, and for the convolution
layers, the reference's own op chain restated with torch ops.
"""
from __future__ import annotations

import os

import torch

import tta_runtime as rt

SMALL_MAX = 96


def pad8(n):
    return (int(n) + 7) // 8 * 8


def needs_autograd(x, params):
    if not torch.is_grad_enabled():
        return False
    return x.requires_grad or any(p is not None and p.requires_grad for p in params)


class Workspace:
    """Zero-initialised scratch tensors cached by (tag, numel, dtype): padding columns are never
    written by the producers, so they stay zero across calls."""

    def __init__(self):
        self.bufs = {}

    def get(self, tag, numel, dtype, device):
        key = (tag, int(numel), dtype, str(device))
        t = self.bufs.get(key)
        if t is None:
            t = torch.zeros(int(numel), dtype=dtype, device=device)
            self.bufs[key] = t
        return t


class PackedWeight:
    """A (N x K) weight matrix cached as bf16 with row stride pad8(K) (zero padded)."""

    def __init__(self, builder, params):
        self.builder = builder
        self.params = [p for p in params if p is not None]
        self.key = None
        self.mat = None
        self.N = self.K = self.ld = 0

    def get(self):
        key = tuple((p.data_ptr(), p._version) for p in self.params)
        if key != self.key:
            with torch.no_grad():
                w = self.builder().detach().to(torch.float32)
                n, k = w.shape
                ld = pad8(k)
                m = torch.zeros(n, ld, dtype=torch.bfloat16, device=w.device)
                m[:, :k] = w.to(torch.bfloat16)
            self.mat, self.N, self.K, self.ld, self.key = m, n, k, ld, key
            self.small = m if ld == k else w.to(torch.bfloat16).contiguous()
        return self


_TTCONV_TC = os.environ.get('TTA_TTCONV_TC', '1') != '0'


class FoldedConv:
    """Cached fp32 operands of the fused factorised-convolution kernel (`tta_ttconv_fused_fwd`):
    a_in (r_a x C_in), kern (r_b, r_a, k, k), a_out (C_out x r_b); rebuilt when a parameter changes."""

    def __init__(self, builder, params):
        self.builder = builder
        self.params = [p for p in params if p is not None]
        self.key = None
        self.ops = None
        self.tc_ok = {}
        self._blob = None
        self._blob_key = None

    def blob(self, bias):
        """Weight image of the tensor-core kernel (tta_ttconv_tc_pack), rebuilt when a weight or the bias changes."""
        key = (self.key, (bias.data_ptr(), bias._version) if bias is not None else None)
        if key != self._blob_key:
            a_in, kern, a_out = self.ops
            self._blob = rt.ttconv_tc_pack(a_in, kern, a_out, bias)
            self._blob_key = key
        return self._blob

    def get(self):
        key = tuple((p.data_ptr(), p._version) for p in self.params)
        if key != self.key:
            with torch.no_grad():
                self.ops = tuple(t.detach().to(torch.float32).contiguous() for t in self.builder())
            self.key = key
        return self.ops


def fused_conv_supported(kernel_size, stride, padding, dilation):
    return (kernel_size[0] == kernel_size[1] and kernel_size[0] in (1, 3) and stride[0] == stride[1] and
            stride[0] in (1, 2) and padding[0] == padding[1] and padding[0] <= kernel_size[0] and
            tuple(dilation) == (1, 1))


def fused_conv(x, folded, bias, kernel_size, stride, padding):
    """y = last(1x1) . core(k x k) . first(1x1) (x) + bias in one kernel; x NCHW (any float dtype) -> NCHW fp32.
    (Kept lean: at ~10 us of GPU time per layer the Python side of the call is what bounds a network.)"""
    a_in, kern, a_out = folded.get()
    B, C, H, W = x.shape
    if x.dtype is not torch.float32 or not x.is_contiguous() or x.requires_grad:
        x = x.detach().to(torch.float32).contiguous()
    ks, s, p = kernel_size[0], stride[0], padding[0]
    cout = a_out.shape[0]
    y = torch.empty((B, cout, (H + 2 * p - ks) // s + 1, (W + 2 * p - ks) // s + 1), dtype=torch.float32, device=x.device)
    if bias is not None and (bias.dtype is not torch.float32 or not bias.is_contiguous()):
        bias = bias.detach().to(torch.float32).contiguous()
    # bf16 tcgen05 kernel for the geometry of the reference's tables (3 x 3, stride 1, pad 1, <= 64 channels / ranks);
    # the fp32 CUDA-core kernel otherwise (and with TTA_TTCONV_TC=0)
    tc_ok = folded.tc_ok.get((C, ks, s, p))
    if tc_ok is None:
        tc_ok = _TTCONV_TC and rt.ttconv_tc_supported(C, a_in.shape[0], kern.shape[0], cout, ks, s, p)
        folded.tc_ok[(C, ks, s, p)] = tc_ok
    if tc_ok:
        rt.ttconv_tc_fwd_raw(x.data_ptr(), folded.blob(bias).data_ptr(), y.data_ptr(), B, C, H, W, a_in.shape[0],
                             kern.shape[0], cout, ks, s, p)
    else:
        rt.ttconv_fused_fwd_raw(x.data_ptr(), a_in.data_ptr(), kern.data_ptr(), a_out.data_ptr(),
                                bias.data_ptr() if bias is not None else None, y.data_ptr(), B, C, H, W, a_in.shape[0],
                                kern.shape[0], cout, ks, s, p)
    return y


def contract(a, M, K, w, out, *, a_inner=1, a_outer=None, lda=None, m_inner=1, s_outer=None, s_inner=0, s_col=1,
             bias=None, bias_inner=0, bias_col=1):
    """out = a (M rows of K) . w^T with the skinny or the tensor-core kernel.

    Row i of `a` starts at (i // a_inner) * a_outer + (i % a_inner) * K (elements); result (i, j) goes to
    (i // m_inner) * s_outer + (i % m_inner) * s_inner + j * s_col.
    """
    w = w.get()
    assert w.K == K, (w.K, K)
    N = w.N
    if a_outer is None:
        a_outer = lda if (lda is not None and a_inner == 1) else K * a_inner
    if s_outer is None:
        s_outer = N
    plain_in = (a_inner == 1)
    plain_out = (m_inner == 1 and s_col == 1 and s_inner == 0)
    if K <= SMALL_MAX and N <= SMALL_MAX:
        rt.small_gemm(a, w.small, out, M, N, K, m_inner=m_inner, s_outer=s_outer, s_inner=s_inner, s_col=s_col, bias=bias,
                      bias_inner=bias_inner, bias_col=bias_col, a_inner=a_inner, a_outer=a_outer)
        return
    if not (plain_in and plain_out):
        raise NotImplementedError('tensor-core contraction needs plain row-major operands (K={}, N={})'.format(K, N))
    if a.dtype != torch.bfloat16:
        raise NotImplementedError('tensor-core contraction needs a bf16 activation')
    if a_outer % 8 or a_outer < w.ld:
        raise NotImplementedError('activation row stride {} not usable with K padded to {}'.format(a_outer, w.ld))
    # K is padded to w.ld: the activation's padding columns are zero by construction (Workspace)
    rt.gemm_bf16_tc(a, w.mat, out, M, N, w.ld, lda=a_outer, ldb=w.ld, ldc=s_outer,
                    bias=bias if bias_col == 1 and bias_inner == 0 else None)
    if bias is not None and not (bias_col == 1 and bias_inner == 0):
        raise NotImplementedError('bias map not supported on the tensor-core path')


class TTRowsEngine:
    """TT-matrix applied to row vectors: the in-core chain, an optional middle operator, the out-core
    chain (TTLinear.py:79-88, TTConv.py:133-147), token-major.

    in_cores[i]  : Parameter (rho_i, n_i, rho_{i+1}), rho_q = 1          (contracted last factor first)
    out_cores[i] : Parameter (r_i, m_i, r_{i+1}),   r_0 = 1
    Supports len(out_cores) <= 2 (all the reference's tables); otherwise the caller uses the op chain.
    """

    def __init__(self, in_cores, out_cores):
        self.in_cores = list(in_cores)
        self.out_cores = list(out_cores)
        self.ws = Workspace()
        self.w_in = [PackedWeight((lambda c=c: c.reshape(c.shape[0], -1)), [c]) for c in self.in_cores]
        p = len(self.out_cores)
        self.w_out = []
        if p >= 1:
            g0 = self.out_cores[0]
            self.w_out.append(PackedWeight((lambda c=g0: c.reshape(c.shape[1], c.shape[2])), [g0]))      # (m_0 x r_1)
        if p == 2:
            g1 = self.out_cores[1]
            self.w_out.append(PackedWeight((lambda c=g1: c.permute(1, 0, 2).reshape(-1, c.shape[2])), [g1]))  # ((o1,a1) x r_2)

    def supported(self):
        return len(self.out_cores) in (1, 2)

    def in_chain(self, x, R, ldx, device):
        """x: (R rows, row stride ldx) fp32 or bf16 -> (z, ld_z) with z (R x rho_0) bf16."""
        q = len(self.in_cores)
        if q == 0:
            return x, ldx
        shapes = [int(c.shape[1]) for c in self.in_cores]
        cur, ld_cur = x, ldx
        for i in range(q - 1, -1, -1):
            c = self.in_cores[i]
            rho_i, n_i, rho_n = int(c.shape[0]), int(c.shape[1]), int(c.shape[2])
            K, N = n_i * rho_n, rho_i
            lead = 1
            for j in range(i):
                lead *= shapes[j]
            ld_out = pad8(lead * N)
            out = self.ws.get(('in', i), R * ld_out, torch.bfloat16, device)
            if lead == 1:
                contract(cur, R, K, self.w_in[i], out, lda=ld_cur, a_outer=ld_cur, s_outer=ld_out)
            else:
                contract(cur, R * lead, K, self.w_in[i], out, a_inner=lead, a_outer=ld_cur, m_inner=lead,
                         s_outer=ld_out, s_inner=N, s_col=1)
            cur, ld_cur = out, ld_out
        return cur, ld_cur

    def out_chain(self, z, R, ldz, y, bias, device):
        """z (R x r_p, stride ldz) bf16 -> y (R x prod m) fp32 contiguous (+ bias)."""
        p = len(self.out_cores)
        if p == 1:
            g0 = self.out_cores[0]
            m0, r1 = int(g0.shape[1]), int(g0.shape[2])
            contract(z, R, r1, self.w_out[0], y, lda=ldz, a_outer=ldz, s_outer=m0, bias=bias, bias_inner=0, bias_col=1)
            return
        g0, g1 = self.out_cores
        m0, r1 = int(g0.shape[1]), int(g0.shape[2])
        m1, r2 = int(g1.shape[1]), int(g1.shape[2])
        ldv = pad8(m1 * r1)
        v = self.ws.get(('out', 1), R * ldv, torch.bfloat16, device)
        contract(z, R, r2, self.w_out[1], v, lda=ldz, a_outer=ldz, s_outer=ldv)
        # y[t, o0*m1 + o1] = sum_a v[t, o1, a] G0[o0, a]
        contract(v, R * m1, r1, self.w_out[0], y, a_inner=m1, a_outer=ldv, m_inner=m1, s_outer=m0 * m1, s_inner=1,
                 s_col=m1, bias=bias, bias_inner=1, bias_col=m1)


def conv_rows(engine_ws, x_nhwc, B, H, W, C, ldx, weight, ksize, stride, padding, dilation, device, tag='conv',
              out_dtype=torch.bfloat16, bias=None):
    """k x k convolution of an NHWC bf16 activation as im2col + tensor-core GEMM.
    `weight`: PackedWeight of the (C_out x kh*kw*C) matrix.  Returns (out, Ho, Wo, ld_out)."""
    kh, kw = ksize
    Ho = (H + 2 * padding[0] - dilation[0] * (kh - 1) - 1) // stride[0] + 1
    Wo = (W + 2 * padding[1] - dilation[1] * (kw - 1) - 1) // stride[1] + 1
    w = weight.get()
    rows = B * Ho * Wo
    if kh == 1 and kw == 1 and stride == (1, 1) and padding == (0, 0):
        cols, ldo = x_nhwc, ldx
    else:
        ldo = w.ld
        cols = engine_ws.get((tag, 'im2col'), rows * ldo, torch.bfloat16, device)
        rt.im2col_bf16(x_nhwc, cols, B, H, W, C, ldx, kh, kw, stride[0], stride[1], padding[0], padding[1],
                       dilation[0], dilation[1], Ho, Wo, ldo)
    ld_out = pad8(w.N) if out_dtype == torch.bfloat16 else w.N
    out = engine_ws.get((tag, 'out'), rows * ld_out, out_dtype, device)
    contract(cols, rows, w.K, weight, out, lda=ldo, a_outer=ldo, s_outer=ld_out, bias=bias)
    return out, Ho, Wo, ld_out


def tt_apply_torch(x2d, in_cores, out_cores):
    """Autograd-capable statement of the same TT-matrix product (training path): x2d (R x in) ->
    (R x out) with out-features in natural (o_0, ..., o_{p-1}) order."""
    R = x2d.shape[0]
    cur = x2d
    for c in reversed(list(in_cores)):
        k = c.shape[1] * c.shape[2]
        cur = cur.reshape(-1, k) @ c.reshape(c.shape[0], k).t()
    acc = cur.reshape(R, 1, -1)
    for g in reversed(list(out_cores)):
        acc = torch.einsum('tpb,aob->topa', acc, g).reshape(R, -1, g.shape[0])
    return acc.reshape(R, -1)


LOWRANK2_MAX_INNER = 384      # TMEM budget of tta_lowrank2_fwd: inner width + two output chunks <= 512 columns


def fold_in_cores(in_cores):
    """(r x in) matrix of the input-side cores: W1[a, (i_0..i_q)] = sum G_0[a,i_0,.] G_1[.,i_1,.] ... (TTLinear.py:79-83)."""
    w = in_cores[0]
    r = w.shape[0]
    for c in in_cores[1:]:
        w = w.reshape(-1, c.shape[0]) @ c.reshape(c.shape[0], -1)
    return w.reshape(r, -1)


def fold_out_cores(out_cores):
    """(out x r) matrix of the output-side cores, rows in natural (o_0, ..., o_{p-1}) order (TTLinear.py:84-88)."""
    w = out_cores[0]
    for c in out_cores[1:]:
        w = w.reshape(-1, c.shape[0]) @ c.reshape(c.shape[0], -1)
    return w.reshape(-1, out_cores[-1].shape[2])


def lowrank2_apply(ws, x2d, w1, w2, bias):
    """y = (x W1^T) W2^T + bias through the fused tcgen05 kernel.  x fp32 (cast to bf16 once) or bf16 (used
    as is); y has the dtype of x.  w1, w2: PackedWeight (bf16, row stride padded to 8)."""
    R, K1 = x2d.shape
    dev = x2d.device
    if x2d.dtype == torch.bfloat16:
        xb = x2d.contiguous()
        out_dtype = torch.bfloat16
    else:
        x32 = x2d.contiguous().to(torch.float32)
        xb = ws.get('xbf16', R * K1, torch.bfloat16, dev)
        rt.cast_bf16(x32.reshape(-1), xb)
        out_dtype = torch.float32
    N1, N2 = w1.N, w2.N
    # the kernel stores 16-byte granules: the row pitch is a multiple of 4 (fp32) / 8 (bf16) elements, the pad columns
    # are sliced off (e.g. a 10-class head)
    if out_dtype == torch.float32:
        ldy = N2 if N2 % 4 == 0 else (N2 + 3) // 4 * 4
    else:
        ldy = N2 if N2 % 8 == 0 else pad8(N2)
    y = torch.empty(R, ldy, dtype=out_dtype, device=dev)
    rt.lowrank2_fwd(xb, w1.mat, w2.mat, bias, y, R, K1, N1, N2, ldx=K1, ld1=w1.ld, ld2=w2.ld, ldy=ldy)
    return y if ldy == N2 else y[:, :N2]


def _pack_bf16(w):
    """(N x K) matrix -> bf16 with row stride pad8(K), zero padded (TMA wants 16-byte row pitches)."""
    n, k = w.shape
    ld = pad8(k)
    if ld == k:
        return w.detach().to(torch.bfloat16).contiguous(), ld
    m = torch.zeros(n, ld, dtype=torch.bfloat16, device=w.device)
    m[:, :k] = w.detach().to(torch.bfloat16)
    return m, ld


def _mm_f32(a, b):
    """bf16 x bf16 -> fp32 product (library GEMM): the token-dimension reductions of the weight gradients keep
    their fp32 accumulator instead of being rounded to bf16 on store."""
    try:
        return torch.mm(a, b, out_dtype=torch.float32)
    except (TypeError, RuntimeError, NotImplementedError):
        return (a @ b).float()


def _grad_tn(a, b, M, N, K, lda, ldb):
    """a^T b (fp32) for a (K x lda), b (K x ldb) bf16: own tcgen05 kernel (csrc/gemm_tn.cu); TTA_GRAD_MM=torch keeps the
    round-1 torch.mm path (cuBLAS) for comparison."""
    if _GRAD_TORCH or rt.backend_is_emulated():
        return _mm_f32(a[:, :M].t(), b[:, :N])
    out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    return rt.gemm_bf16_tn(a, b, out, M, N, K, lda=lda, ldb=ldb, ldc=N)


_GRAD_TORCH = os.environ.get('TTA_GRAD_MM', '') == 'torch'


class LowRank2Fn(torch.autograd.Function):
    """Training path of the two-factor layers (backward of the fused forwards):
    y = (x W1^T) W2^T + bias with the forward AND the input gradient on the fused TMA + tcgen05 kernel.

        dX  = (dY W2) W1         -- `tta_lowrank2_fwd` again, factors W2^T (N1 x N2) and W1^T (K1 x N1)
        dW2 = dY^T V,  V = x W1^T (recomputed by `tta_gemm_bf16_tc`, not stored)
        dW1 = dV^T x,  dV = dY W2 (`tta_gemm_bf16_tc`)
    The two weight gradients reduce over the token dimension (transposed-A products): they are plain library
    GEMMs (`torch.matmul`, bf16 operands, fp32 result).  Operands are rounded to bf16 as in the inference path;
    W1 / W2 are functions of the layer's cores built with autograd-tracked torch ops, so the core gradients
    follow by the chain rule.  Requires K1 % 8 == 0, N2 % 8 == 0, N1 <= LOWRANK2_MAX_INNER.
    """

    @staticmethod
    def forward(ctx, x2d, w1, w2, bias):
        rt.require_device(x2d)
        R, K1 = x2d.shape
        N1, N2 = w1.shape[0], w2.shape[0]
        xb = x2d.detach().contiguous().to(torch.bfloat16)
        w1b, ld1 = _pack_bf16(w1)
        w2b, ld2 = _pack_bf16(w2)
        out_dtype = torch.bfloat16 if x2d.dtype == torch.bfloat16 else torch.float32
        y = torch.empty(R, N2, dtype=out_dtype, device=x2d.device)
        b32 = bias.detach().to(torch.float32).contiguous() if bias is not None else None
        rt.lowrank2_fwd(xb, w1b, w2b, b32, y, R, K1, N1, N2, ldx=K1, ld1=ld1, ld2=ld2, ldy=N2)
        ctx.save_for_backward(xb, w1b, w2b)
        ctx.dims = (R, K1, N1, N2, ld1, ld2, x2d.dtype, w1.dtype, w2.dtype, bias is not None,
                    bias.dtype if bias is not None else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, w1b, w2b = ctx.saved_tensors
        R, K1, N1, N2, ld1, ld2, xdt, w1dt, w2dt, has_bias, bdt = ctx.dims
        dev = dy.device
        dyb = dy.detach().contiguous().to(torch.bfloat16)
        n1p = pad8(N1)
        w2t = w2b[:, :N1].t().contiguous()                       # (N1 x N2), row pitch N2 (a multiple of 8)
        w1t = torch.zeros(K1, n1p, dtype=torch.bfloat16, device=dev)
        w1t[:, :N1] = w1b[:, :K1].t()                            # (K1 x N1), row pitch pad8(N1)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(R, K1, dtype=torch.bfloat16 if xdt == torch.bfloat16 else torch.float32, device=dev)
            rt.lowrank2_fwd(dyb, w2t, w1t, None, dx, R, N2, N1, K1, ldx=N2, ld1=N2, ld2=n1p, ldy=K1)
            dx = dx.to(xdt)
        dw1 = dw2 = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            v = torch.zeros(R, n1p, dtype=torch.bfloat16, device=dev)
            rt.gemm_bf16_tc(xb, w1b, v, R, N1, K1, lda=K1, ldb=ld1, ldc=n1p)          # V = x W1^T
            dw2 = _grad_tn(dyb, v, N2, N1, R, N2, n1p).to(w2dt)                       # dW2 = dY^T V
            dv = torch.zeros(R, n1p, dtype=torch.bfloat16, device=dev)
            rt.gemm_bf16_tc(dyb, w2t, dv, R, N1, N2, lda=N2, ldb=N2, ldc=n1p)         # dV = dY W2
            dw1 = _grad_tn(dv, xb, N1, K1, R, n1p, K1).to(w1dt)                       # dW1 = dV^T X
        db = dy.sum(0).to(bdt) if has_bias and ctx.needs_input_grad[3] else None
        return dx, dw1, dw2, db


def lowrank2_trainable(in_features, out_features, r_mid):
    return in_features % 8 == 0 and out_features % 8 == 0 and r_mid <= LOWRANK2_MAX_INNER


def tt_chain_macs(in_cores, out_cores):
    """Multiply-accumulates per row of the factorised chain (TTLinear.py:79-88 order)."""
    macs = 0
    shapes = [int(c.shape[1]) for c in in_cores]
    for i in range(len(in_cores) - 1, -1, -1):
        c = in_cores[i]
        lead = 1
        for j in range(i):
            lead *= shapes[j]
        macs += lead * int(c.shape[1]) * int(c.shape[2]) * int(c.shape[0])
    p = 1
    for g in reversed(list(out_cores)):
        macs += p * int(g.shape[0]) * int(g.shape[1]) * int(g.shape[2])
        p *= int(g.shape[1])
    return macs


def split_tt(tt_shapes, out_channels, conv):
    """TTConv.py:49-68 / TTLinear.py:31-40: the first prefix of tt_shapes whose product equals the
    number of outputs is the 'out' part; for convs the next entry is kh*kw."""
    prod = 1
    for i, s in enumerate(tt_shapes):
        prod *= s
        if prod == out_channels:
            out_order = i + 1
            break
    else:
        raise ValueError('tt_shapes {} do not factor {} outputs'.format(tt_shapes, out_channels))
    in_order = len(tt_shapes) - out_order - (1 if conv else 0)
    return out_order, in_order


def to_rows(ws, x, tag='rows'):
    """NCHW fp32 -> pixel-major bf16 rows (B*H*W x pad8(C)); returns (rows, ld)."""
    B, C, H, W = x.shape
    ld = pad8(C)
    rows = ws.get((tag, 'nhwc'), B * H * W * ld, torch.bfloat16, x.device)
    rt.nchw_to_nhwc_bf16(x.contiguous().to(torch.float32), rows, B, C, H * W, ld)
    return rows, ld


def from_rows(rows, ld, B, C, Ho, Wo, bias, device):
    """pixel-major rows (bf16 or fp32) -> NCHW fp32 with the bias add fused."""
    y = torch.empty(B, C, Ho, Wo, dtype=torch.float32, device=device)
    rt.nhwc_to_nchw_f32(rows, y, bias, B, C, Ho * Wo, ld)
    return y


def conv_weight_matrix(w4d):
    """(O, I, kh, kw) -> (O, kh*kw*I) matching the im2col column order (kh, kw, c)."""
    return w4d.permute(0, 2, 3, 1).reshape(w4d.shape[0], -1)
