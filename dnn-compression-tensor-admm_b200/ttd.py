"""Drop-in replacement for the reference's `ttd.py` (`from ttd import ten2tt, tt2ten`, used by
TTConv.py:20, TTLinear.py:20, admm.py:12) running on the B200 kernels of libtta.so.

Conventions kept (ttd.py:10-43): numpy in / numpy out, cores are a list of arrays shaped
(r_i, s_i, r_{i+1}), and `ten2tt` clips `tt_ranks` IN PLACE when an unfolding has fewer singular
values than requested (ttd.py:18-19).  Singular-vector signs / rotations inside a degenerate
subspace are not unique; the product of the cores (what every caller consumes) is.
"""
from __future__ import annotations

import numpy as np
import torch

import projector
import tta_runtime as rt


def _device():
    if rt.backend_is_emulated():
        return torch.device('cpu')
    if not torch.cuda.is_available():
        raise rt.TtaError('ttd: no CUDA device -- this implementation runs on B200 (sm_100a) only')
    return torch.device('cuda', torch.cuda.current_device())


def ten2tt(x, tt_shapes, tt_ranks):
    shapes = [int(s) for s in tt_shapes]
    x = np.ascontiguousarray(x, dtype=np.float32)
    numel = int(x.size)
    layer = projector.TTLayer('ten2tt', (shapes[0], numel // shapes[0]), shapes, list(tt_ranks))
    for i, r in enumerate(layer.ranks):   # in-place clip, ttd.py:18-19: only a clipped entry is written (a tuple
        if int(tt_ranks[i]) != r:         # raises TypeError exactly when the reference would)
            tt_ranks[i] = r
    dev = _device()
    xt = torch.from_numpy(x.reshape(shapes[0], -1)).to(dev).contiguous()
    zt = torch.empty_like(xt)
    # skip_full_rank off: a step that keeps every singular triplet still gets its eigensolve, so that core i
    # reshaped to (r_i * s_i, r_{i+1}) has orthonormal columns like the U of ttd.py:21-25 (the projection plans of
    # admm.py serve such steps by A = I * A, which gives the same product but identity / raw-carry cores)
    plan = projector.TTProjectionPlan([layer], dev, skip_full_rank=False)
    plan.run([xt], [None], [zt])
    return [c.detach().cpu().numpy().copy() for c in plan.cores(0)]


def tt2ten(tt_cores, tt_shapes):
    dev = _device()
    cores = [torch.as_tensor(np.ascontiguousarray(c, dtype=np.float32)).to(dev).contiguous() for c in tt_cores]
    acc = cores[0].reshape(-1, cores[0].shape[-1]).contiguous()
    for core in cores[1:]:
        r = core.shape[0]
        left = acc.reshape(-1, r).contiguous()
        right = core.reshape(r, -1).contiguous()
        out = torch.empty(left.shape[0], right.shape[1], dtype=torch.float32, device=dev)
        task = np.zeros(1, dtype=rt.GEMM_TASK)
        task[0] = (left.data_ptr(), right.data_ptr(), out.data_ptr(), 0, r, 1, right.shape[1], 1,
                   right.shape[1], left.shape[0], right.shape[1], r, 0)
        rt.gemm(rt.TaskTable(task, dev))
        acc = out
    return acc.cpu().numpy().reshape(tt_shapes)
