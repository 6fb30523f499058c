"""Layer-sharded Z-update across the ranks of one box.

Every listed layer's projection is independent (admm.py:43-44 has no cross-layer state) and, under
DDP, every rank holds identical W and U and calls `ADMM.update()` at the same point of every epoch
(engines.py:270-271).  The reference therefore repeats the whole projection on every rank; here

  1. layers are assigned to ranks by a deterministic longest-processing-time greedy on a per-layer
     cost model (identical on all ranks -- it depends on shapes only);
  2. each rank projects only its own layers, writing Z into its slab of one flat fp32 buffer;
  3. ONE all-gather over NCCL (NVLink 5 / NVSwitch) makes every Z resident on every rank;
  4. the fused U-update then runs locally on all layers (inputs identical => U bit-identical).

There is no other data-path collective.  With world_size == 1 (or torch.distributed not
initialised) this degenerates to "all layers local, no exchange".
"""
from __future__ import annotations

import torch

import projector

EIG_WEIGHT = 4.5   # Jacobi does ~40 k^3 flops at CUDA-core rate vs the 9 k^3 accounting figure


def layer_cost(kind, weight_shape, ranks, tt_shapes=None):
    if kind == 'tk':
        o, i = int(weight_shape[0]), int(weight_shape[1])
        kk = 1
        for v in weight_shape[2:]:
            kk *= int(v)
        r0, r1 = int(ranks[0]), int(ranks[1])
        sweeps = 8
        per_sweep = 2 * o * i * kk * (r0 + r1) + 2 * o * o * kk * r1 + 2 * i * i * kk * r0
        eig = EIG_WEIGHT * 9 * (o ** 3 + i ** 3)
        return float(sweeps * (per_sweep + eig))
    g, p, e, r = projector.tt_step_flops(tt_shapes, ranks)
    return float(g + p + r + EIG_WEIGHT * e)


def lpt_assign(costs, world_size):
    """Greedy LPT: heaviest layer first onto the least-loaded rank.  Returns owner rank per layer."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world_size
    owner = [0] * len(costs)
    for i in order:
        r = min(range(world_size), key=lambda q: (load[q], q))
        owner[i] = r
        load[r] += costs[i]
    return owner, load


def _dist_state():
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return None, 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


class LayerSharding:
    def __init__(self, admm, names, group=None):
        self.names = list(names)
        self.group = group
        self.dist, self.rank, self.world = _dist_state()
        self.flat_w = None
        if self.world == 1:
            self.local_names = list(self.names)
            self.owner = {n: 0 for n in self.names}
            return
        costs = []
        for n in self.names:
            p = admm._params[n]
            kind = admm._classify(n, p)
            if kind == 'svd':
                r = admm.hp_dict.ranks[n]
                r = r if isinstance(r, int) else r[0]
                costs.append(layer_cost('tt', p.shape, [1, r, 1], [int(p.shape[0]), int(p.shape[1])]))
            elif kind == 'tt':
                costs.append(layer_cost('tt', p.shape, admm.hp_dict.ranks[n], admm.hp_dict.tt_shapes[n]))
            else:
                costs.append(layer_cost('tk', p.shape, admm.hp_dict.ranks[n]))
        owner, self.load = lpt_assign(costs, self.world)
        self.owner = {n: o for n, o in zip(self.names, owner)}
        self.local_names = [n for n in self.names if self.owner[n] == self.rank]

        # flat Z buffer: [rank 0 slab | rank 1 slab | ...], slabs padded to the largest
        dev = admm._state_device()
        sizes = [0] * self.world
        self.offset = {}
        for n in self.names:
            r = self.owner[n]
            self.offset[n] = sizes[r]
            sizes[r] += admm._params[n].numel()
        self.slab = max(max(sizes), 1)
        self.flat = torch.zeros(self.slab * self.world, dtype=torch.float32, device=dev)
        for n in self.names:
            p = admm._params[n]
            start = self.owner[n] * self.slab + self.offset[n]
            view = self.flat[start:start + p.numel()].view(p.shape)
            view.copy_(admm.z[n])
            admm.z[n] = view
        admm._ew_cache = None

    def exchange_weights(self, params, local_names, remote_names):
        """All-gather of the ranks' weight slabs (same layout as the Z buffer): after `update_from_host` uploaded
        only the local layers' weights, every rank needs W of every layer for the dual update (admm.py:73)."""
        if self.world == 1:
            return
        if self.flat_w is None:
            self.flat_w = torch.zeros_like(self.flat)
        view = lambda n: self.flat_w[self.owner[n] * self.slab + self.offset[n]:
                                     self.owner[n] * self.slab + self.offset[n] + params[n].numel()].view(params[n].shape)
        for n in local_names:
            view(n).copy_(params[n].data)
        mine = self.flat_w[self.rank * self.slab:(self.rank + 1) * self.slab]
        if self.flat_w.is_cuda:
            self.dist.all_gather_into_tensor(self.flat_w, mine, group=self.group)
        else:
            outs = [self.flat_w[r * self.slab:(r + 1) * self.slab] for r in range(self.world)]
            self.dist.all_gather(outs, mine.clone(), group=self.group)
        for n in remote_names:
            params[n].data.copy_(view(n))

    def exchange(self, z):
        """One all-gather of the Z slabs (no-op for a single rank)."""
        if self.world == 1:
            return
        mine = self.flat[self.rank * self.slab:(self.rank + 1) * self.slab]
        if self.flat.is_cuda:
            self.dist.all_gather_into_tensor(self.flat, mine, group=self.group)
        else:  # gloo (CPU tests): list form
            outs = [self.flat[r * self.slab:(r + 1) * self.slab] for r in range(self.world)]
            self.dist.all_gather(outs, mine.clone(), group=self.group)
