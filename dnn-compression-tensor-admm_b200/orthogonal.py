"""Drop-in replacement for the reference's `orthogonal.py` (`append_double_l2_loss(model, loss, rho, device)`,
orthogonal.py:9-20; called every optimiser step of a fine-tune run by engines.py:290-291,297-298).

For every parameter whose name contains `first_kernel`, `last_kernel`, `first_factor`, `last_factor` or
`left_kernel` the reference adds rho/2 * ||F F^T - I||_F^2 (F = squeeze(param), fewer rows than columns) or
rho/2 * ||F^T F - I||_F^2 (otherwise) through ~7 torch ops per factor plus their autograd twins.  Here all
factors of the model go through ONE batched forward launch (Gram, minus identity, squared norm; the residual is kept)
and ONE batched backward launch (2 rho R F), both in libtta.so (`csrc/orth.cu`).
"""
from __future__ import annotations

import numpy as np
import torch

import tta_runtime as rt

_KEYS = ('first_kernel', 'last_kernel', 'first_factor', 'last_factor', 'left_kernel')


def _selected(model):
    return [(n, p) for n, p in model.named_parameters() if any(k in n for k in _KEYS)]


def _geometry(p):
    """n vectors of length len with strides (si, st) over the squeezed 2-D factor (orthogonal.py:14-19)."""
    m = torch.squeeze(p)
    if m.dim() != 2:
        raise ValueError('orthogonality regulariser: parameter of shape {} does not squeeze to a matrix'.format(tuple(p.shape)))
    rows, cols = int(m.shape[0]), int(m.shape[1])
    if rows < cols:
        return rows, cols, cols, 1          # rows of F:    F F^T - I
    return cols, rows, 1, cols              # columns of F: F^T F - I


class _OrthFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, loss, rho, *params):
        dev = params[0].device
        tasks = np.zeros(len(params), dtype=rt.ORTH_TASK)
        res = []
        for i, p in enumerate(params):
            rt.require_device(p)
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise rt.TtaError('orthogonality regulariser: factors must be contiguous fp32 parameters')
            n, length, si, st = _geometry(p)
            r = torch.empty(n * n, dtype=torch.float32, device=dev)
            res.append(r)
            tasks[i] = (p.data_ptr(), r.data_ptr(), 0, si, st, n, length)
        acc = torch.zeros(1, dtype=torch.float64, device=dev)
        rt.orth_penalty_fwd(rt.TaskTable(tasks, dev), rho, acc)
        ctx.rho = float(rho)
        ctx.tasks = tasks
        ctx.res = res
        ctx.params = params
        return loss + acc.to(loss.dtype).reshape(loss.shape).to(loss.device)

    @staticmethod
    def backward(ctx, grad_out):
        params = ctx.params
        dev = params[0].device
        scale = grad_out.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        grads = [torch.empty_like(p) for p in params]
        tasks = ctx.tasks.copy()
        for i, g in enumerate(grads):
            tasks[i]['g'] = g.data_ptr()
        rt.orth_penalty_bwd(rt.TaskTable(tasks, dev), ctx.rho, scale, accumulate=False)
        return (grad_out, None) + tuple(grads)


def append_double_l2_loss(model, loss, rho, device):
    sel = _selected(model)
    if not sel:
        return loss
    return _OrthFn.apply(loss, rho, *[p for _, p in sel])
