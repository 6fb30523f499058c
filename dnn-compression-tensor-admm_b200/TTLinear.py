"""Drop-in replacement for the reference's `TTLinear.py` (`TTLinearM`, `TTLinearR`): same constructor
signature, parameter names / shapes (`tt_cores.{i}` (r_i, s_i, r_{i+1}), `bias`) and xavier init
(TTLinear.py:23-160), so checkpoints and `vit_tt.py` work unchanged.

forward (inference): the TT chain runs token-major in bf16 on libtta.so -- skinny outer-core
contractions + tcgen05 tensor-core GEMMs for the big middle ones (`fwd_common.TTRowsEngine`).
forward (autograd): the same product with torch ops (fused backward is SURVEY 8(f) next).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F
from torch.nn import Module, Parameter, ParameterList, init

import fwd_common as fc
import tta_runtime as rt
from ttd import ten2tt


class _TTLinearBase(Module):
    def __init__(self, in_features, out_features, bias=True, hp_dict=None, name=None, dense_w=None, dense_b=None):
        super().__init__()
        self.tt_shapes = list(hp_dict.tt_shapes[name])
        self.tt_order = len(self.tt_shapes)
        self.out_tt_order, self.in_tt_order = fc.split_tt(self.tt_shapes, out_features, conv=False)
        self.out_tt_shapes = self.tt_shapes[:self.out_tt_order]
        self.in_tt_shapes = self.tt_shapes[self.out_tt_order:]
        assert in_features == int(np.prod(self.in_tt_shapes))
        assert out_features == int(np.prod(self.out_tt_shapes))
        self.in_features = in_features
        self.out_features = out_features
        self.tt_ranks = list(hp_dict.ranks[name])
        self.tt_cores = ParameterList([Parameter(torch.empty(self.tt_ranks[i], self.tt_shapes[i], self.tt_ranks[i + 1]))
                                       for i in range(self.tt_order)])
        if bias:
            self.bias = Parameter(torch.zeros(self.out_features))
            if dense_b is not None:
                self.bias.data = dense_b
        else:
            self.register_parameter('bias', None)
        if dense_w is not None:
            cores = ten2tt(dense_w.detach().cpu().numpy(), self.tt_shapes, self.tt_ranks)
            for i, c in enumerate(cores):
                self.tt_cores[i].data = torch.from_numpy(np.ascontiguousarray(c))
        else:
            self.reset_parameters()
        self._engine = None
        self._dense_engine = None
        self._dense_first = None
        self._fused2 = None
        self.fused_training = False     # additive: route the autograd path through the fused kernels (bf16 operands)

    def get_ranks(self):
        return ', '.join(str(r) for r in self.tt_ranks)

    def extra_repr(self):
        return 'in_features={}, out_features={}, tt_shapes={}, tt_ranks={}'.format(
            self.in_features, self.out_features, self.tt_shapes, self.tt_ranks)

    def _recover_weight(self):
        w = self.tt_cores[0]
        for i in range(1, self.tt_order):
            w = w.reshape(-1, self.tt_ranks[i]).mm(self.tt_cores[i].reshape(self.tt_ranks[i], -1))
        return w.reshape(self.out_features, self.in_features)


class TTLinearM(_TTLinearBase):
    def reset_parameters(self):
        for c in self.tt_cores:
            init.xavier_uniform_(c)

    def forward(self, x):
        params = list(self.tt_cores) + [self.bias]
        out_shape = list(x.shape)
        out_shape[-1] = self.out_features
        in_cores = list(self.tt_cores)[self.out_tt_order:]
        out_cores = list(self.tt_cores)[:self.out_tt_order]
        x2d = x.reshape(-1, self.in_features)
        if fc.needs_autograd(x, params):
            # Fused training path (forward and dX on the two-factor tcgen05 kernel, bf16 operands): taken under
            # autocast -- the reference trains under torch.cuda.amp.autocast (engines.py:285) -- or when
            # `fused_training` is set; otherwise the fp32 torch op chain below.
            if ((self.fused_training or torch.is_autocast_enabled()) and x.is_cuda and
                    fc.lowrank2_trainable(self.in_features, self.out_features, int(self.tt_ranks[self.out_tt_order]))):
                with torch.autocast('cuda', enabled=False):
                    w1 = fc.fold_in_cores([c.float() for c in in_cores])
                    w2 = fc.fold_out_cores([c.float() for c in out_cores])
                    y = fc.LowRank2Fn.apply(x2d, w1, w2, self.bias)
                return y.reshape(out_shape)
            y = fc.tt_apply_torch(x2d, in_cores, out_cores)
            if self.bias is not None:
                y = y + self.bias
            return y.reshape(out_shape)
        rt.require_device(x)
        if self._fused2 is None:
            # Two-factor form: the input-side cores fold into W1 (r x in) and the output-side cores into
            # W2 (out x r), r = the rank between the two sides (weights only, cached).  One fused tcgen05
            # kernel then runs in -> r -> out with the rank-r intermediate kept on the SM
            # (`tta_lowrank2_fwd`).  Taken when r fits its TMEM budget and the two dense factors cost no
            # more than 1.5x the MACs of the four-step chain (DeiT-small tables: 0.94x ... 1.25x).
            r_mid = int(self.tt_ranks[self.out_tt_order])
            macs2 = r_mid * (self.in_features + self.out_features)
            self._fused2 = bool(r_mid <= fc.LOWRANK2_MAX_INNER and self.in_features % 8 == 0 and
                                self.in_tt_order >= 1 and self.out_tt_order >= 1 and
                                macs2 <= 1.5 * fc.tt_chain_macs(in_cores, out_cores))
            if self._fused2:
                self._w1 = fc.PackedWeight(lambda: fc.fold_in_cores(list(self.tt_cores)[self.out_tt_order:]),
                                           list(self.tt_cores)[self.out_tt_order:])
                self._w2 = fc.PackedWeight(lambda: fc.fold_out_cores(list(self.tt_cores)[:self.out_tt_order]),
                                           list(self.tt_cores)[:self.out_tt_order])
                self._ws = fc.Workspace()
        if self._fused2:
            with torch.no_grad():
                y = fc.lowrank2_apply(self._ws, x2d, self._w1.get(), self._w2.get(), self.bias)
            return y.reshape(out_shape)
        if self._dense_first is None:
            # Contraction order: folding the cores into the (out x in) matrix first is a weights-only
            # computation; it is taken when the dense product is less work per token than the chain AND
            # clearly so (measured on B200: with fp32 activations in and out, one dense tcgen05 GEMM plus the
            # bf16 cast is slower than the chain for DeiT-small, 24.7 ms vs 19.3 ms over the 48 layers, even
            # where its MAC count is comparable).
            self._dense_first = (self.in_features % 8 == 0 and
                                 2 * self.in_features * self.out_features <= fc.tt_chain_macs(in_cores, out_cores))
        if self._dense_first:
            if self._dense_engine is None:
                self._dense_engine = (fc.Workspace(), fc.PackedWeight(self._recover_weight, list(self.tt_cores)))
            ws, w = self._dense_engine
            with torch.no_grad():
                x2d = x2d.contiguous().to(torch.float32)
                R = x2d.shape[0]
                xb = ws.get('xbf16', R * self.in_features, torch.bfloat16, x.device)
                rt.cast_bf16(x2d.reshape(-1), xb)
                y = torch.empty(R, self.out_features, dtype=torch.float32, device=x.device)
                fc.contract(xb, R, self.in_features, w, y, lda=self.in_features, a_outer=self.in_features,
                            s_outer=self.out_features, bias=self.bias)
            return y.reshape(out_shape)
        if self._engine is None:
            self._engine = fc.TTRowsEngine(in_cores, out_cores)
        eng = self._engine
        if not eng.supported():
            y = fc.tt_apply_torch(x2d, in_cores, out_cores)
            return (y + self.bias if self.bias is not None else y).reshape(out_shape)
        with torch.no_grad():
            x2d = x2d.contiguous().to(torch.float32)
            R = x2d.shape[0]
            last = in_cores[-1]
            if last.shape[1] * last.shape[2] > fc.SMALL_MAX or last.shape[0] > fc.SMALL_MAX:
                xb = eng.ws.get('xbf16', R * fc.pad8(self.in_features), torch.bfloat16, x.device)
                if fc.pad8(self.in_features) != self.in_features:
                    raise NotImplementedError('in_features must be a multiple of 8 for the tensor-core first step')
                rt.cast_bf16(x2d.reshape(-1), xb)
                x_in, ldx = xb, self.in_features
            else:
                x_in, ldx = x2d, self.in_features
            z, ldz = eng.in_chain(x_in, R, ldx, x.device)
            y = torch.empty(R, self.out_features, dtype=torch.float32, device=x.device)
            eng.out_chain(z, R, ldz, y, self.bias, x.device)
        return y.reshape(out_shape)


class TTLinearR(_TTLinearBase):
    def reset_parameters(self):
        for c in self.tt_cores:
            init.xavier_uniform_(c)
        if self.bias is not None:
            bound = 1 / math.sqrt(self.in_features)
            init.uniform_(self.bias, -bound, bound)

    def forward(self, x):
        params = list(self.tt_cores) + [self.bias]
        if fc.needs_autograd(x, params) or self.in_features % 8:
            return F.linear(x, self._recover_weight(), self.bias)
        rt.require_device(x)
        if self._engine is None:
            self._engine = (fc.Workspace(), fc.PackedWeight(self._recover_weight, list(self.tt_cores)))
        ws, w = self._engine
        with torch.no_grad():
            out_shape = list(x.shape)
            out_shape[-1] = self.out_features
            x2d = x.reshape(-1, self.in_features).contiguous().to(torch.float32)
            R = x2d.shape[0]
            xb = ws.get('xbf16', R * self.in_features, torch.bfloat16, x.device)
            rt.cast_bf16(x2d.reshape(-1), xb)
            y = torch.empty(R, self.out_features, dtype=torch.float32, device=x.device)
            fc.contract(xb, R, self.in_features, w, y, lda=self.in_features, a_outer=self.in_features,
                        s_outer=self.out_features, bias=self.bias)
        return y.reshape(out_shape)
