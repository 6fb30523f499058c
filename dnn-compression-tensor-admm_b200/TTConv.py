"""Drop-in replacement for the reference's `TTConv.py` (`TTConv2dM`, `TTConv2dR`; TTConv.py:23-333):
same constructor signature, `ValueError`s, parameter names / shapes (`in_tt_cores.{i}`, `core_kernel`,
`out_tt_cores.{i}`, `bias` for M; `out_tt_cores.{i}`, `conv_core (r, k*k, r)`, `in_tt_cores.{i}` for R),
`dense_w` decomposition (ten2tt on (O, k*k, I), TTConv.py:96-109; the R variant's untransposed
quirk, :284-296) and xavier init.

forward (inference, TTConv2dM): NCHW -> pixel-major bf16 rows -> in-core chain -> k x k core convolution
(im2col + tcgen05 GEMM) -> out-core chain -> NCHW (+bias), every step a libtta.so kernel
(`fwd_common`).  forward (autograd): the same contraction with torch ops.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F
from torch.nn import Module, Parameter, ParameterList, init
from torch.nn.modules.utils import _pair

import fwd_common as fc
import tta_runtime as rt
from ttd import ten2tt


def _check_args(groups, padding_mode):
    if groups != 1:
        raise ValueError("groups must be 1 in this mode")
    if padding_mode != 'zeros':
        raise ValueError("padding_mode must be zero in this mode")


class _TTConvBase(Module):
    def _setup(self, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, padding_mode, hp_dict,
               name):
        self.kernel_size = _pair(kernel_size)
        self.stride = _pair(stride)
        self.padding = _pair(padding)
        self.dilation = _pair(dilation)
        self.transposed = False
        self.output_padding = _pair(0)
        self.groups = groups
        self.padding_mode = padding_mode
        self.tt_shapes = list(hp_dict.tt_shapes[name])
        self.tt_order = len(self.tt_shapes)
        self.out_tt_order, self.in_tt_order = fc.split_tt(self.tt_shapes, out_channels, conv=True)
        self.out_tt_shapes = self.tt_shapes[:self.out_tt_order]
        self.in_tt_shapes = self.tt_shapes[self.out_tt_order + 1:]
        assert in_channels == int(np.prod(self.in_tt_shapes))
        assert out_channels == int(np.prod(self.out_tt_shapes))
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.tt_ranks = list(hp_dict.ranks[name])
        self.out_tt_ranks = self.tt_ranks[:self.out_tt_order + 1]
        self.in_tt_ranks = self.tt_ranks[self.out_tt_order + 1:]
        self.filter_dim = self.kernel_size[0] * self.kernel_size[1]
        self._engine = None
        self._folded = None
        self._fusable = fc.fused_conv_supported(self.kernel_size, self.stride, self.padding, self.dilation)

    def get_ranks(self):
        return ', '.join(str(r) for r in self.tt_ranks)

    def extra_repr(self):
        return '{}, {}, kernel_size={}, stride={}, padding={}, tt_shapes={}, tt_ranks={}'.format(
            self.in_channels, self.out_channels, self.kernel_size, self.stride, self.padding, self.tt_shapes,
            self.tt_ranks)


class TTConv2dM(_TTConvBase):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 padding_mode='zeros', hp_dict=None, name=None, dense_w=None, dense_b=None):
        _check_args(groups, padding_mode)
        super().__init__()
        self._setup(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, padding_mode, hp_dict, name)
        self.in_tt_cores = ParameterList([Parameter(torch.empty(self.in_tt_ranks[i], self.in_tt_shapes[i],
                                                                self.in_tt_ranks[i + 1]))
                                          for i in range(self.in_tt_order)])
        self.core_kernel = Parameter(torch.empty(self.out_tt_ranks[-1], self.in_tt_ranks[0], *self.kernel_size))
        self.out_tt_cores = ParameterList([Parameter(torch.empty(self.out_tt_ranks[i], self.out_tt_shapes[i],
                                                                 self.out_tt_ranks[i + 1]))
                                           for i in range(self.out_tt_order)])
        if bias:
            self.bias = Parameter(torch.zeros(self.out_channels))
            if dense_b is not None:
                self.bias.data = dense_b
        else:
            self.register_parameter('bias', None)
        if dense_w is not None:
            w = dense_w.detach().cpu().numpy().reshape(self.out_channels, self.in_channels, self.filter_dim)
            cores = ten2tt(np.transpose(w, (0, 2, 1)), self.tt_shapes, self.tt_ranks)
            for i, c in enumerate(cores):
                t = torch.from_numpy(np.ascontiguousarray(c))
                if i < self.out_tt_order:
                    self.out_tt_cores[i].data = t
                elif i == self.out_tt_order:
                    self.core_kernel.data = t.permute(0, 2, 1).reshape(
                        self.out_tt_ranks[-1], self.in_tt_ranks[0], *self.kernel_size).contiguous()
                else:
                    self.in_tt_cores[i - self.out_tt_order - 1].data = t
        else:
            self.reset_parameters()

    def reset_parameters(self):
        for c in self.out_tt_cores:
            init.xavier_uniform_(c)
        for c in self.in_tt_cores:
            init.xavier_uniform_(c)
        init.xavier_uniform_(self.core_kernel)

    def _params(self):
        return list(self.in_tt_cores) + [self.core_kernel] + list(self.out_tt_cores) + [self.bias]

    def _forward_torch(self, x):
        B, C, H, W = x.shape
        rows = x.permute(0, 2, 3, 1).reshape(-1, C)
        z = fc.tt_apply_torch(rows, list(self.in_tt_cores), [])
        z = z.reshape(B, H, W, -1).permute(0, 3, 1, 2)
        z = F.conv2d(z, self.core_kernel, None, self.stride, self.padding, self.dilation, self.groups)
        _, R2, Ho, Wo = z.shape
        rows = z.permute(0, 2, 3, 1).reshape(-1, R2)
        y = fc.tt_apply_torch(rows, [], list(self.out_tt_cores))
        y = y.reshape(B, Ho, Wo, self.out_channels).permute(0, 3, 1, 2)
        if self.bias is not None:
            y = y + self.bias.view(1, -1, 1, 1)
        return y

    def forward(self, x):
        if torch.is_grad_enabled() and fc.needs_autograd(x, self._params()):
            return self._forward_torch(x)
        rt.require_device(x)
        if self._fusable:
            if self._folded is None:
                def build():
                    eye_in = torch.eye(self.in_channels, dtype=torch.float32, device=self.core_kernel.device)
                    eye_out = torch.eye(self.out_tt_ranks[-1], dtype=torch.float32, device=self.core_kernel.device)
                    a_in = fc.tt_apply_torch(eye_in, list(self.in_tt_cores), []).t()       # (r_a x C_in)
                    a_out = fc.tt_apply_torch(eye_out, [], list(self.out_tt_cores)).t()    # (C_out x r_b)
                    return a_in, self.core_kernel, a_out
                self._folded = fc.FoldedConv(build, [p for p in self._params() if p is not self.bias])
            return fc.fused_conv(x, self._folded, self.bias, self.kernel_size, self.stride, self.padding)
        if self._engine is None:
            eng = fc.TTRowsEngine(list(self.in_tt_cores), list(self.out_tt_cores))
            wk = fc.PackedWeight(lambda: fc.conv_weight_matrix(self.core_kernel), [self.core_kernel])
            self._engine = (eng, wk)
        eng, wk = self._engine
        if not eng.supported():
            return self._forward_torch(x)
        with torch.no_grad():
            B, C, H, W = x.shape
            dev = x.device
            rows, ld = fc.to_rows(eng.ws, x)
            z, ldz = eng.in_chain(rows, B * H * W, ld, dev)
            c, Ho, Wo, ldc = fc.conv_rows(eng.ws, z, B, H, W, self.in_tt_ranks[0], ldz, wk, self.kernel_size, self.stride,
                                          self.padding, self.dilation, dev)
            R2 = B * Ho * Wo
            y_rows = eng.ws.get('yrows', R2 * self.out_channels, torch.float32, dev)
            eng.out_chain(c, R2, ldc, y_rows, None, dev)
            return fc.from_rows(y_rows, self.out_channels, B, self.out_channels, Ho, Wo, self.bias, dev)

    def forward_flops(self, x):
        out = self.forward(x)
        _, _, H, W = x.shape
        _, _, Ho, Wo = out.shape
        tt = 0.0
        lead = int(np.prod(self.in_tt_shapes))
        for i in range(self.in_tt_order - 1, -1, -1):
            lead //= self.in_tt_shapes[i]
            tt += H * W * lead * self.in_tt_ranks[i] * self.in_tt_shapes[i] * self.in_tt_ranks[i + 1]
        tt += Ho * Wo * self.core_kernel.numel()
        tail = 1
        for i in range(self.out_tt_order - 1, -1, -1):
            tt += Ho * Wo * tail * self.out_tt_ranks[i] * self.out_tt_shapes[i] * self.out_tt_ranks[i + 1]
            tail *= self.out_tt_shapes[i]
        tt_flops = tt / 1e6
        base_flops = Ho * Wo * self.filter_dim * self.in_channels * self.out_channels / 1e6
        tt_params = sum(p.numel() for p in self._params() if p is not None and p is not self.bias)
        base_params = self.filter_dim * self.in_channels * self.out_channels
        print('baseline # params: {:.2f}K, tt # params: {:.2f}K'.format(base_params / 1000, tt_params / 1000))
        print('baseline # flops: {:.2f}M, tt # flops: {:.2f}M'.format(base_flops, tt_flops))
        return out, base_flops, tt_flops


class TTConv2dR(_TTConvBase):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 padding_mode='zeros', hp_dict=None, name=None, dense_w=None, dense_b=None):
        _check_args(groups, padding_mode)
        super().__init__()
        self._setup(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, padding_mode, hp_dict, name)
        self.kernel_shape = [out_channels, in_channels // groups, *self.kernel_size]
        self.out_tt_cores = ParameterList([Parameter(torch.empty(self.out_tt_ranks[i], self.out_tt_shapes[i],
                                                                 self.out_tt_ranks[i + 1]))
                                           for i in range(self.out_tt_order)])
        self.conv_core = Parameter(torch.empty(self.out_tt_ranks[-1], self.filter_dim, self.in_tt_ranks[0]))
        self.in_tt_cores = ParameterList([Parameter(torch.empty(self.in_tt_ranks[i], self.in_tt_shapes[i],
                                                                self.in_tt_ranks[i + 1]))
                                          for i in range(self.in_tt_order)])
        if bias:
            self.bias = Parameter(torch.zeros(self.out_channels))
            if dense_b is not None:
                self.bias.data = dense_b
        else:
            self.register_parameter('bias', None)
        if dense_w is not None:
            # TTConv.py:284-288: (O, I, k*k) is decomposed WITHOUT the (0,2,1) transpose while the shapes are
            # labelled out + [k*k] + in -- a self-consistent quirk that is reproduced as-is
            w = dense_w.detach().cpu().numpy().reshape(self.out_channels, self.in_channels, -1)
            shapes = self.out_tt_shapes + [w.shape[-1]] + self.in_tt_shapes
            cores = ten2tt(w, shapes, self.tt_ranks)
            for i, c in enumerate(cores):
                t = torch.from_numpy(np.ascontiguousarray(c))
                if i < self.out_tt_order:
                    self.out_tt_cores[i].data = t
                elif i == self.out_tt_order:
                    self.conv_core.data = t
                else:
                    self.in_tt_cores[i - self.out_tt_order - 1].data = t
        else:
            self.reset_parameters()

    def reset_parameters(self):
        for c in self.out_tt_cores:
            init.xavier_uniform_(c)
        init.xavier_uniform_(self.conv_core)
        for c in self.in_tt_cores:
            init.xavier_uniform_(c)
        if self.bias is not None:
            bound = 1 / math.sqrt(self.in_channels * self.filter_dim)
            init.uniform_(self.bias, -bound, bound)

    def _recover_weight(self):
        w = self.out_tt_cores[0]
        for i in range(1, self.out_tt_order):
            w = w.reshape(-1, self.out_tt_ranks[i]) @ self.out_tt_cores[i].reshape(self.out_tt_ranks[i], -1)
        w = w.reshape(-1, self.out_tt_ranks[-1]) @ self.conv_core.reshape(self.out_tt_ranks[-1], -1)
        for i in range(self.in_tt_order):
            w = w.reshape(-1, self.in_tt_ranks[i]) @ self.in_tt_cores[i].reshape(self.in_tt_ranks[i], -1)
        return w.reshape(self.kernel_shape)      # (O, k*k, I) memory read as (O, I, kh, kw): TTConv.py:321

    def _params(self):
        return list(self.out_tt_cores) + [self.conv_core] + list(self.in_tt_cores) + [self.bias]

    def forward(self, x):
        if torch.is_grad_enabled() and fc.needs_autograd(x, self._params()):
            return F.conv2d(x, self._recover_weight(), self.bias, self.stride, self.padding, self.dilation, self.groups)
        rt.require_device(x)
        if self._engine is None:
            self._engine = (fc.Workspace(), fc.PackedWeight(lambda: fc.conv_weight_matrix(self._recover_weight()),
                                                           [p for p in self._params() if p is not self.bias]))
        ws, wk = self._engine
        with torch.no_grad():
            B, C, H, W = x.shape
            rows, ld = fc.to_rows(ws, x)
            y_rows, Ho, Wo, ldy = fc.conv_rows(ws, rows, B, H, W, C, ld, wk, self.kernel_size, self.stride, self.padding,
                                               self.dilation, x.device, out_dtype=torch.float32)
            return fc.from_rows(y_rows, ldy, B, self.out_channels, Ho, Wo, self.bias, x.device)
