"""Drop-in replacement for the reference's `TKConv.py` (`TKConv2dC`, `TKConv2dM`, `TKConv2dR`;
TKConv.py:26-325): same constructor signature and `ValueError`s, parameter names / shapes
(`first_kernel (r_in, I, 1, 1)`, `core_kernel (r_out, r_in, k, k)`, `last_kernel (O, r_out, 1, 1)` for C;
`first_factor (r_in, I)`, `core_kernel` / `core_tensor`, `last_factor (O, r_out)` for M / R), `dense_w`
decomposition by Tucker-2 HOOI (`projector.tucker2_decompose`), xavier init.

forward (inference): pixel-major bf16 rows; 1x1 -> k x k (im2col) -> 1x1 as tcgen05 / skinny GEMMs of
libtta.so with the NCHW<->NHWC conversions and the bias fused into the layout kernels.
forward (autograd): torch ops.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch.nn import Module, Parameter, init
from torch.nn.modules.utils import _pair

import fwd_common as fc
import projector
import tta_runtime as rt


class _TKConvBase(Module):
    def _setup(self, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, padding_mode, hp_dict,
               name):
        if groups != 1:
            raise ValueError("groups must be 1 in this mode")
        if padding_mode != 'zeros':
            raise ValueError("padding_mode must be zero in this mode")
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.ranks = hp_dict.ranks[name]
        self.in_rank = self.ranks[1]
        self.out_rank = self.ranks[0]
        self.kernel_size = _pair(kernel_size)
        self.stride = _pair(stride)
        self.padding = _pair(padding)
        self.dilation = _pair(dilation)
        self.transposed = False
        self.output_padding = _pair(0)
        self.groups = groups
        self.padding_mode = padding_mode
        self._engine = None

    def _decompose(self, dense_w):
        core, (last, first) = projector.tucker2_decompose(dense_w, [self.out_rank, self.in_rank])
        # a requested rank above the mode size yields only `size` vectors (as tensorly's svd slice does)
        self.out_rank, self.in_rank = int(last.shape[1]), int(first.shape[1])
        return core.cpu(), last.cpu(), first.t().contiguous().cpu()

    def _fused(self, x, first2d, core4d, last2d, bias, params):
        """1x1 (I -> r_in), k x k (r_in -> r_out), 1x1 (r_out -> O): one fused fp32 kernel when the geometry is
        supported (k in {1, 3}, stride <= 2), else pixel-major bf16 rows through the GEMM kernels."""
        if fc.fused_conv_supported(self.kernel_size, self.stride, self.padding, self.dilation):
            if getattr(self, '_folded', None) is None:
                self._folded = fc.FoldedConv(lambda: (first2d(), core4d(), last2d()), params[:3])
            return fc.fused_conv(x, self._folded, bias, self.kernel_size, self.stride, self.padding)
        if self._engine is None:
            self._engine = (fc.Workspace(), fc.PackedWeight(first2d, [params[0]]),
                            fc.PackedWeight(lambda: fc.conv_weight_matrix(core4d()), [params[1]]),
                            fc.PackedWeight(last2d, [params[2]]))
        ws, w1, wk, w3 = self._engine
        with torch.no_grad():
            B, C, H, W = x.shape
            dev = x.device
            rows, ld = fc.to_rows(ws, x)
            R = B * H * W
            l1 = fc.pad8(self.in_rank)
            a1 = ws.get('a1', R * l1, torch.bfloat16, dev)
            fc.contract(rows, R, C, w1, a1, lda=ld, a_outer=ld, s_outer=l1)
            a2, Ho, Wo, l2 = fc.conv_rows(ws, a1, B, H, W, self.in_rank, l1, wk, self.kernel_size, self.stride, self.padding,
                                          self.dilation, dev)
            R2 = B * Ho * Wo
            y_rows = ws.get('yrows', R2 * self.out_channels, torch.float32, dev)
            fc.contract(a2, R2, self.out_rank, w3, y_rows, lda=l2, a_outer=l2, s_outer=self.out_channels)
            return fc.from_rows(y_rows, self.out_channels, B, self.out_channels, Ho, Wo, bias, dev)

    def extra_repr(self):
        return '{}, {}, kernel_size={}, stride={}, padding={}, ranks={}'.format(
            self.in_channels, self.out_channels, self.kernel_size, self.stride, self.padding, list(self.ranks))


class TKConv2dC(_TKConvBase):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 padding_mode='zeros', hp_dict=None, name=None, dense_w=None, dense_b=None):
        super().__init__()
        self._setup(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, padding_mode, hp_dict, name)
        self.first_kernel = Parameter(torch.empty(self.in_rank, self.in_channels, 1, 1))
        self.core_kernel = Parameter(torch.empty(self.out_rank, self.in_rank, *self.kernel_size))
        self.last_kernel = Parameter(torch.empty(self.out_channels, self.out_rank, 1, 1))
        if bias:
            self.bias = Parameter(torch.zeros(out_channels))
            if dense_b is not None:
                self.bias.data = dense_b
        else:
            self.register_parameter('bias', None)
        if dense_w is not None:
            core, last, first = self._decompose(dense_w)
            self.first_kernel.data = first.unsqueeze(-1).unsqueeze(-1)
            self.last_kernel.data = last.unsqueeze(-1).unsqueeze(-1)
            self.core_kernel.data = core
        else:
            self.reset_parameters()

    def reset_parameters(self):
        init.xavier_uniform_(self.first_kernel)
        init.xavier_uniform_(self.core_kernel)
        init.xavier_uniform_(self.last_kernel)

    def _params(self):
        return [self.first_kernel, self.core_kernel, self.last_kernel, self.bias]

    def forward(self, x):
        if torch.is_grad_enabled() and fc.needs_autograd(x, self._params()):
            return self.forward_features(x)[0]
        rt.require_device(x)
        return self._fused(x, lambda: self.first_kernel.reshape(self.in_rank, -1), lambda: self.core_kernel,
                           lambda: self.last_kernel.reshape(self.out_channels, -1), self.bias, self._params())

    def forward_features(self, x):
        features = []
        out = F.conv2d(x, self.first_kernel)
        features.append(out)
        out = F.conv2d(out, self.core_kernel, None, self.stride, self.padding, self.dilation, self.groups)
        features.append(out)
        out = F.conv2d(out, self.last_kernel, self.bias)
        features.append(out)
        return out, features

    def forward_flops(self, x):
        out = self.forward(x)
        _, _, H, W = x.shape
        _, _, Ho, Wo = out.shape
        k2 = self.kernel_size[0] * self.kernel_size[1]
        tk_flops = (H * W * self.in_rank * self.in_channels + Ho * Wo * self.core_kernel.numel()
                    + Ho * Wo * self.out_channels * self.out_rank) / 1e6
        base_flops = Ho * Wo * k2 * self.in_channels * self.out_channels / 1e6
        print('baseline # flops: {:.2f}M, tk # flops: {:.2f}M'.format(base_flops, tk_flops))
        return out, base_flops, tk_flops


class TKConv2dM(_TKConvBase):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 padding_mode='zeros', hp_dict=None, name=None, dense_w=None, dense_b=None):
        super().__init__()
        self._setup(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, padding_mode, hp_dict, name)
        self.first_factor = Parameter(torch.empty(self.in_rank, in_channels))
        self.core_kernel = Parameter(torch.empty(self.out_rank, self.in_rank, *self.kernel_size))
        self.last_factor = Parameter(torch.empty(out_channels, self.out_rank))
        if bias:
            self.bias = Parameter(torch.zeros(out_channels))
            if dense_b is not None:
                self.bias.data = dense_b
        else:
            self.register_parameter('bias', None)
        if dense_w is not None:
            core, last, first = self._decompose(dense_w)
            self.first_factor.data = first
            self.last_factor.data = last
            self.core_kernel.data = core
        else:
            self.reset_parameters()

    def reset_parameters(self):
        init.xavier_uniform_(self.first_factor)
        init.xavier_uniform_(self.last_factor)
        init.xavier_uniform_(self.core_kernel)

    def _params(self):
        return [self.first_factor, self.core_kernel, self.last_factor, self.bias]

    def forward(self, x):
        if torch.is_grad_enabled() and fc.needs_autograd(x, self._params()):
            out = F.linear(x.permute(0, 2, 3, 1), self.first_factor).permute(0, 3, 1, 2)
            out = F.conv2d(out, self.core_kernel, None, self.stride, self.padding, self.dilation, self.groups)
            return F.linear(out.permute(0, 2, 3, 1), self.last_factor, self.bias).permute(0, 3, 1, 2)
        rt.require_device(x)
        return self._fused(x, lambda: self.first_factor, lambda: self.core_kernel, lambda: self.last_factor, self.bias,
                           self._params())


class TKConv2dR(_TKConvBase):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 padding_mode='zeros', hp_dict=None, name=None, dense_w=None, dense_b=None):
        super().__init__()
        self._setup(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, padding_mode, hp_dict, name)
        self.kernel_shape = [out_channels, in_channels // groups, *self.kernel_size]
        self.filter_dim = self.kernel_size[0] * self.kernel_size[1]
        self.first_factor = Parameter(torch.empty(self.in_rank, in_channels))
        self.core_tensor = Parameter(torch.empty(self.out_rank, self.in_rank, *self.kernel_size))
        self.last_factor = Parameter(torch.empty(out_channels, self.out_rank))
        if bias:
            self.bias = Parameter(torch.zeros(out_channels))
            if dense_b is not None:
                self.bias.data = dense_b
        else:
            self.register_parameter('bias', None)
        if dense_w is not None:
            core, last, first = self._decompose(dense_w)
            self.first_factor.data = first
            self.last_factor.data = last
            self.core_tensor.data = core
        else:
            self.reset_parameters()

    def reset_parameters(self):
        init.xavier_uniform_(self.first_factor)
        init.xavier_uniform_(self.core_tensor)
        init.xavier_uniform_(self.last_factor)
        if self.bias is not None:
            bound = 1 / math.sqrt(self.in_channels * self.filter_dim)
            init.uniform_(self.bias, -bound, bound)

    def _recover_weight(self):
        # core x_0 last_factor x_1 first_factor^T   (tl.tucker_to_tensor, TKConv.py:314)
        w = torch.einsum('or,rskl->oskl', self.last_factor, self.core_tensor)
        return torch.einsum('oskl,si->oikl', w, self.first_factor)

    def _params(self):
        return [self.first_factor, self.core_tensor, self.last_factor, self.bias]

    def forward(self, x):
        if torch.is_grad_enabled() and fc.needs_autograd(x, self._params()):
            return F.conv2d(x, self._recover_weight(), self.bias, self.stride, self.padding, self.dilation, self.groups)
        rt.require_device(x)
        if self._engine is None:
            self._engine = (fc.Workspace(), fc.PackedWeight(lambda: fc.conv_weight_matrix(self._recover_weight()),
                                                           self._params()[:3]))
        ws, wk = self._engine
        with torch.no_grad():
            B, C, H, W = x.shape
            rows, ld = fc.to_rows(ws, x)
            y_rows, Ho, Wo, ldy = fc.conv_rows(ws, rows, B, H, W, C, ld, wk, self.kernel_size, self.stride, self.padding,
                                               self.dilation, x.device, out_dtype=torch.float32)
            return fc.from_rows(y_rows, ldy, B, self.out_channels, Ho, Wo, self.bias, x.device)
