"""Drop-in replacement for the reference's `TKLinear.py` (`TKLinearM`, `TKLinearR`; TKLinear.py:23-122):
same constructor, parameters `first_factor (r_in, in)`, `core_tensor (r_out, r_in)`,
`last_factor (out, r_out)`, `bias`.  Inference forward (M) = one fused two-factor tcgen05 kernel (the core is
folded into the smaller side; three separate GEMMs when the inner rank exceeds its TMEM budget), (R) = one GEMM
with the rebuilt weight; autograd forward = the torch op chain."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch.nn import Module, Parameter, init

import fwd_common as fc
import projector
import tta_runtime as rt


class _TKLinearBase(Module):
    def __init__(self, in_features, out_features, bias=True, hp_dict=None, name=None, dense_w=None, dense_b=None):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.ranks = hp_dict.ranks[name]
        self.in_rank = self.ranks[1]
        self.out_rank = self.ranks[0]
        self.first_factor = Parameter(torch.empty(self.in_rank, self.in_features))
        self.core_tensor = Parameter(torch.empty(self.out_rank, self.in_rank))
        self.last_factor = Parameter(torch.empty(self.out_features, self.out_rank))
        if bias:
            self.bias = Parameter(torch.zeros(self.out_features))
            if dense_b is not None:
                self.bias.data = dense_b
        else:
            self.register_parameter('bias', None)
        if dense_w is not None:
            core, (last, first) = projector.tucker2_decompose(dense_w, [self.out_rank, self.in_rank])
            self.out_rank, self.in_rank = int(last.shape[1]), int(first.shape[1])
            self.first_factor.data = first.t().contiguous().cpu()
            self.last_factor.data = last.cpu()
            self.core_tensor.data = core.cpu()
        else:
            self.reset_parameters()
        self._engine = None
        self._fused = None

    def reset_parameters(self):
        init.kaiming_uniform_(self.first_factor, a=math.sqrt(5))
        init.kaiming_uniform_(self.core_tensor, a=math.sqrt(5))
        init.kaiming_uniform_(self.last_factor, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1 / math.sqrt(self.in_features)
            init.uniform_(self.bias, -bound, bound)

    def _recover_weight(self):
        return self.last_factor @ self.core_tensor @ self.first_factor

    def _params(self):
        return [self.first_factor, self.core_tensor, self.last_factor, self.bias]


class TKLinearM(_TKLinearBase):
    fused_training = False     # additive: route the autograd path through the fused kernels (bf16 operands)

    def forward(self, x):
        if (fc.needs_autograd(x, self._params()) and (self.fused_training or torch.is_autocast_enabled()) and x.is_cuda and
                fc.lowrank2_trainable(self.in_features, self.out_features, min(self.in_rank, self.out_rank))):
            # fused training path: forward and dX on the two-factor tcgen05 kernel, factor
            # gradients by the chain rule through the folded core (TKLinear.py:60-75)
            with torch.autocast('cuda', enabled=False):
                if self.in_rank <= self.out_rank:
                    w1, w2 = self.first_factor.float(), self.last_factor.float() @ self.core_tensor.float()
                else:
                    w1, w2 = self.core_tensor.float() @ self.first_factor.float(), self.last_factor.float()
                out_shape = list(x.shape)
                out_shape[-1] = self.out_features
                y = fc.LowRank2Fn.apply(x.reshape(-1, self.in_features), w1, w2, self.bias)
            return y.reshape(out_shape)
        if fc.needs_autograd(x, self._params()) or self.in_features % 8:
            out = F.linear(x, self.first_factor)
            out = F.linear(out, self.core_tensor)
            return F.linear(out, self.last_factor, self.bias)
        rt.require_device(x)
        # Two-factor form: the core folds into the smaller side (weights only, cached), then ONE fused tcgen05
        # kernel runs in -> r -> out with the rank-r intermediate on the SM (`tta_lowrank2_fwd`).
        r_mid = min(self.in_rank, self.out_rank)
        if r_mid <= fc.LOWRANK2_MAX_INNER:
            if self._fused is None:
                if self.in_rank <= self.out_rank:      # (last core) (first): inner width r_in
                    w1 = fc.PackedWeight(lambda: self.first_factor, [self.first_factor])
                    w2 = fc.PackedWeight(lambda: self.last_factor @ self.core_tensor, [self.last_factor, self.core_tensor])
                else:                                  # (last) (core first): inner width r_out
                    w1 = fc.PackedWeight(lambda: self.core_tensor @ self.first_factor, [self.core_tensor, self.first_factor])
                    w2 = fc.PackedWeight(lambda: self.last_factor, [self.last_factor])
                self._fused = (fc.Workspace(), w1, w2)
            ws, w1, w2 = self._fused
            with torch.no_grad():
                out_shape = list(x.shape)
                out_shape[-1] = self.out_features
                y = fc.lowrank2_apply(ws, x.reshape(-1, self.in_features), w1.get(), w2.get(), self.bias)
            return y.reshape(out_shape)
        if self._engine is None:
            self._engine = (fc.Workspace(), fc.PackedWeight(lambda: self.first_factor, [self.first_factor]),
                            fc.PackedWeight(lambda: self.core_tensor, [self.core_tensor]),
                            fc.PackedWeight(lambda: self.last_factor, [self.last_factor]))
        ws, w1, w2, w3 = self._engine
        with torch.no_grad():
            out_shape = list(x.shape)
            out_shape[-1] = self.out_features
            x2d = x.reshape(-1, self.in_features).contiguous().to(torch.float32)
            R, dev = x2d.shape[0], x.device
            xb = ws.get('x', R * self.in_features, torch.bfloat16, dev)
            rt.cast_bf16(x2d.reshape(-1), xb)
            l1, l2 = fc.pad8(self.in_rank), fc.pad8(self.out_rank)
            a1 = ws.get('a1', R * l1, torch.bfloat16, dev)
            fc.contract(xb, R, self.in_features, w1, a1, lda=self.in_features, a_outer=self.in_features, s_outer=l1)
            a2 = ws.get('a2', R * l2, torch.bfloat16, dev)
            fc.contract(a1, R, self.in_rank, w2, a2, lda=l1, a_outer=l1, s_outer=l2)
            y = torch.empty(R, self.out_features, dtype=torch.float32, device=dev)
            fc.contract(a2, R, self.out_rank, w3, y, lda=l2, a_outer=l2, s_outer=self.out_features, bias=self.bias)
        return y.reshape(out_shape)


class TKLinearR(_TKLinearBase):
    def forward(self, x):
        if fc.needs_autograd(x, self._params()) or self.in_features % 8:
            return F.linear(x, self._recover_weight(), self.bias)
        rt.require_device(x)
        if self._engine is None:
            self._engine = (fc.Workspace(), fc.PackedWeight(self._recover_weight,
                                                           [self.first_factor, self.core_tensor, self.last_factor]))
        ws, w = self._engine
        with torch.no_grad():
            out_shape = list(x.shape)
            out_shape[-1] = self.out_features
            x2d = x.reshape(-1, self.in_features).contiguous().to(torch.float32)
            R, dev = x2d.shape[0], x.device
            xb = ws.get('x', R * self.in_features, torch.bfloat16, dev)
            rt.cast_bf16(x2d.reshape(-1), xb)
            y = torch.empty(R, self.out_features, dtype=torch.float32, device=dev)
            fc.contract(xb, R, self.in_features, w, y, lda=self.in_features, a_outer=self.in_features,
                        s_outer=self.out_features, bias=self.bias)
        return y.reshape(out_shape)
