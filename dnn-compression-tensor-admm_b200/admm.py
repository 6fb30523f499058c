"""Drop-in replacement for the reference's `admm.py` (same import surface: `from admm import ADMM`,
engines.py:32,241-246,270-271,288-296) whose Z-update, U-update and penalty run as batched sm_100a
kernels through libtta.so instead of per-layer numpy / LAPACK / tensorly calls on the CPU.

Behavioural contract kept from the reference (admm.py:15-149):
  * ctor signature, `Exception('ERROR: Tensor format should be specified!')` for format 'none'
    (:27-28) and `Exception('ERROR: unsupported layer in ADMM!')` for listed params that are neither
    2-D nor 4-D (:68-69);
  * `.u`, `.z` dicts name -> fp32 tensor on `device`, `.rho`, `.init_rho`, `.logger`, `.hp_dict`,
    `.format`, `.device`; `update(update_u=True)`, `append_admm_loss(loss)`, `adjust_rho(...)`;
  * dispatch by tensor rank / rank-list length (:47-67); the in-place rank clip of ttd.py:18-19 is
    written back into `hp_dict.ranks[name]` for conv weights (:94 passes the list by reference) and
    not for linear weights (:105 copies);
  * `prune_*_rank_*` helpers (numpy in / numpy out for tt & tk, torch in / numpy out for svd).

Differences (all additive): `update()` is one batched pass over all layers; under
`torch.distributed` (world_size > 1) the layers are sharded over the ranks by an LPT greedy on the
per-layer FLOP model and the projected Z tensors are exchanged with one all-gather (`sharding.py`);
`state_dict()/load_state_dict()` expose u/z (the reference silently drops them on resume).
"""
from __future__ import annotations

import numpy as np
import os
import time

import torch

import projector
import tta_runtime as rt


def _is_multi(rank_entry):
    return (not isinstance(rank_entry, int)) and len(rank_entry) > 1


def _svd_rank(rank_entry):
    return rank_entry if isinstance(rank_entry, int) else rank_entry[0]


class _PenaltyFn(torch.autograd.Function):
    """loss + rho/2 * sum_l ||W_l - Z_l + U_l||^2 with one fused forward and one fused backward launch."""

    @staticmethod
    def forward(ctx, loss, admm, *params):
        ctx.admm = admm
        acc = torch.zeros(1, dtype=torch.float64, device=admm._state_device())
        rt.penalty_fwd(admm._ew_table(), admm.rho, acc)
        ctx.rho = float(admm.rho)
        return loss + acc.to(loss.dtype).reshape(loss.shape).to(loss.device)

    @staticmethod
    def backward(ctx, grad_out):
        admm = ctx.admm
        dev = admm._state_device()
        scale = grad_out.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        flat = torch.empty(admm._total_numel, dtype=torch.float32, device=dev)
        tab = admm._ew_table(grad_flat=flat)
        rt.penalty_bwd(tab, ctx.rho, scale, accumulate=False)
        grads = []
        off = 0
        for name in admm._names:
            p = admm._params[name]
            grads.append(flat[off:off + p.numel()].view(p.shape))
            off += p.numel()
        return (grad_out, None) + tuple(grads)


def _priority_range():
    """(lowest, highest) stream priority of the device; numerically the highest priority is the smallest number."""
    try:
        lo, hi = torch.cuda.Stream.priority_range()
        return int(lo), int(hi)
    except Exception:
        return 0, -1


class ADMM:
    def __init__(self, model, rho, hp_dict, format, device, verbose=False, log=False):
        self.model = model
        self.init_rho = rho
        self.hp_dict = hp_dict
        self.format = format
        self.device = device
        self.verbose = verbose
        self.log = log
        if self.log:
            self.logger = {}

        if format == 'none':
            raise Exception('ERROR: Tensor format should be specified!')

        self.rho = self.init_rho

        self.u = {}
        self.z = {}
        self._names = []
        self._params = {}
        for name, param in self.model.named_parameters():
            if name in self.hp_dict.ranks:
                rt.require_device(param)
                self.u[name] = torch.zeros(param.shape, dtype=torch.float32, device=param.device)
                self.z[name] = param.data.detach().clone().to(torch.float32).contiguous()
                self._names.append(name)
                self._params[name] = param
                if self.log:
                    self.logger[name] = []
        self._total_numel = sum(self._params[n].numel() for n in self._names)
        self._plans = None
        self._ew_cache = None
        self.sweeps = {}
        self._xw_stream = None
        self._xw_done = None
        self._shard = None
        self._streams = None
        self._copy_stream = None
        self.concurrent_groups = True    # additive: False runs all TT layers as one lock-step plan

    # ------------------------------------------------------------------------------------------
    def _state_device(self):
        return self._params[self._names[0]].device if self._names else torch.device(self.device)

    def _classify(self, name, param):
        """admm.py:47-69 dispatch -> ('tt'|'tk'|'svd', layer description)."""
        ranks = self.hp_dict.ranks[name]
        nd = param.dim()
        if nd == 4:
            if self.format == 'tk' and _is_multi(ranks):
                return 'tk'
            if self.format == 'tt' and _is_multi(ranks):
                return 'tt'
            return 'svd'
        if nd == 2:
            if self.format == 'tk':
                return 'tk'
            if self.format == 'tt':
                return 'tt'
            return 'svd'
        raise Exception('ERROR: unsupported layer in ADMM!')

    def _build_plans(self, names):
        dev = self._state_device()
        tt_layers, tt_names, tk_layers, tk_names = [], [], [], []
        for name in names:
            p = self._params[name]
            kind = self._classify(name, p)
            if kind == 'tt':
                L = projector.TTLayer(name, p.shape, self.hp_dict.tt_shapes[name], self.hp_dict.ranks[name])
                if p.dim() == 4:
                    # ttd.py:18-19 mutates the caller's list; admm.py:94 passes hp_dict's own list
                    try:
                        self.hp_dict.ranks[name][:] = L.ranks
                    except TypeError:
                        pass  # immutable tuple in a user table: nothing to write back into
                tt_layers.append(L)
                tt_names.append(name)
            elif kind == 'svd':
                r = _svd_rank(self.hp_dict.ranks[name])
                if p.dim() == 4 and (p.shape[2] != 1 or p.shape[3] != 1):
                    raise Exception('ERROR: unsupported layer in ADMM!')
                o, i = int(p.shape[0]), int(p.shape[1])
                tt_layers.append(projector.TTLayer(name, (o, i), [o, i], [1, r, 1]))
                tt_names.append(name)
            else:
                tk_layers.append(projector.TKLayer(name, p.shape, self.hp_dict.ranks[name]))
                tk_names.append(name)
        plans = []
        if tt_layers:
            # independent layer groups, most critical chain first; each runs on its own stream in update()
            for g in (projector.tt_layer_groups(tt_layers) if self.concurrent_groups else [list(range(len(tt_layers)))]):
                plans.append((projector.TTProjectionPlan([tt_layers[i] for i in g], dev), [tt_names[i] for i in g]))
        if tk_layers:
            plans.append((projector.TKProjectionPlan(tk_layers, dev), tk_names))
        return plans

    def _ew_table(self, grad_flat=None):
        key = tuple((self._params[n].data_ptr(), self.z[n].data_ptr(), self.u[n].data_ptr()) for n in self._names)
        if grad_flat is None and self._ew_cache is not None and self._ew_cache[0] == key:
            return self._ew_cache[1]
        arr = np.zeros(len(self._names), dtype=rt.EW_TASK)
        off = 0
        for i, n in enumerate(self._names):
            p = self._params[n]
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise rt.TtaError('parameter {} must be contiguous fp32'.format(n))
            g = grad_flat.data_ptr() + 4 * off if grad_flat is not None else 0
            arr[i] = (p.data_ptr(), self.z[n].data_ptr(), self.u[n].data_ptr(), g, p.numel())
            off += p.numel()
        tab = rt.TaskTable(arr, self._state_device())
        if grad_flat is None:
            self._ew_cache = (key, tab)
        return tab

    # ------------------------------------------------------------------------------------------
    def update_from_host(self, host_w, host_z=None, update_u=True, gather_host_z=False):
        """Additive: `update()` for weights that live in (pinned) host memory.  `host_w[name]` is copied into the
        parameter and, when `host_z` is given, Z is copied back into `host_z[name]` -- group by group on the
        groups' own streams, so the transfers overlap the projection of the other groups (the critical group is
        uploaded first, and only its download is exposed at the end).  Synchronise the current stream before reading
        `host_z`.

        Under torch.distributed every rank moves only ITS layers over PCIe: it uploads the weights of the layers it
        projects, the other layers' weights arrive with a second all-gather over NVLink next to the one of Z (the dual
        update needs W of every layer on every rank), and it downloads only the Z of its own layers -- the ranks' host
        buffers together hold the result.  `gather_host_z=True` makes every rank download every layer's Z instead."""
        self.update(update_u, _host_in=host_w, _host_out=host_z, _gather_host_z=gather_host_z)

    def update(self, update_u=True, _host_in=None, _host_out=None, _gather_host_z=False):
        if not self._names:
            return
        import sharding
        if self._plans is None:
            self._shard = sharding.LayerSharding(self, self._names)
            self._plans = self._build_plans(self._shard.local_names)
        with torch.no_grad():
            local = set(self._shard.local_names)
            remote = [n for n in self._names if n not in local]
            # W of the layers other ranks project: one all-gather of the weight slabs (NVLink), no PCIe traffic.
            # Up to four ranks it is issued inside _run_plans, right after the uploads of the local layers, on a side stream,
            # and overlaps the projection (only the dual update needs the result): e2e 5.64 -> 5.34 ms at N = 2, 5.21 ->
            # 4.92 ms at N = 4.  With 8 ranks that was measured slower (5.25 -> 6.7 ms: the NCCL CTAs spin until the slowest
            # rank has uploaded, scattered over the GPCs, and a 16-CTA eigensolver cluster needs 16 free SMs of ONE GPC), so
            # there the exchange follows the all-gather of Z as in round 1.  TTA_XW_OVERLAP=0 / 1 overrides.  The order of
            # the two collectives is the same on every rank in either mode.
            need_w = self._shard.world > 1 and _host_in is not None      # every rank takes part, whatever it owns
            env = os.environ.get('TTA_XW_OVERLAP')
            overlap = need_w and ((env == '1') if env in ('0', '1') else self._shard.world <= 4)
            self._xw_done = None
            self._run_plans(_host_in, _host_out, remote if overlap else None)
            self._shard.exchange(self.z)
            if self._xw_done is not None:
                torch.cuda.current_stream(self._state_device()).wait_event(self._xw_done)
            if need_w and not overlap:
                self._shard.exchange_weights(self._params, self._shard.local_names, remote)
            if remote and _host_out is not None and _gather_host_z:
                self._copy(_host_out, remote, False)
            if update_u:
                want_norm = self.log or self.verbose
                sq = torch.zeros(len(self._names), dtype=torch.float64, device=self._state_device()) if want_norm else None
                rt.dual_update(self._ew_table(), sq)
                if want_norm:
                    norms = sq.sqrt().tolist()
                    for n, v in zip(self._names, norms):
                        if self.log:
                            self.logger[n].append(float(v))
                        if self.verbose:
                            print('*INFO: {} in ADMM, norm(w-z)={}'.format(n, v))

    def _exchange_weights(self, remote, after_events=None):
        """All-gather of W (sharding.exchange_weights); on a side stream behind `after_events` when given."""
        if remote is None:
            return
        if after_events is None:
            self._shard.exchange_weights(self._params, self._shard.local_names, remote)
            return
        dev = self._state_device()
        if self._xw_stream is None:
            self._xw_stream = torch.cuda.Stream(device=dev)
        for ev in after_events:
            self._xw_stream.wait_event(ev)
        with torch.cuda.stream(self._xw_stream):
            self._shard.exchange_weights(self._params, self._shard.local_names, remote)
            self._xw_done = torch.cuda.Event()
            self._xw_done.record(self._xw_stream)

    def _run_plans(self, host_in=None, host_out=None, remote=None):
        """Z-update of the local layers.  Several TT plans (layer groups) are enqueued on side streams -- the
        first two on high-priority streams -- forked from and joined back into the current stream, so that
        the Gram / refinement kernels of one group overlap the eigensolves of another and short chains do not
        wait for long ones; sweep counts are read back (the only host sync) after everything is enqueued."""
        args = lambda names: ([self._params[n].data for n in names], [self.u[n] for n in names],
                              [self.z[n] for n in names])
        async_plans = [(pl, nm) for pl, nm in self._plans if hasattr(pl, 'enqueue')]
        if len(async_plans) > 1 and not rt.backend_is_emulated() and all(pl.profile is None for pl, _ in async_plans):
            dev = self._state_device()
            if self._streams is None or len(self._streams) != len(async_plans):
                # descending priorities in the order of the groups (most critical first): when SMs free up, the block
                # scheduler serves the pending CTAs of the critical chain before those of the shorter chains
                lo, hi = _priority_range()
                prios = os.environ.get('TTA_GROUP_PRIOS')
                prios = [int(v) for v in prios.split(',')] if prios else list(range(hi, lo + 1))
                self._streams = [torch.cuda.Stream(device=dev, priority=max(hi, min(lo, prios[min(i, len(prios) - 1)])))
                                 for i in range(len(async_plans))]
            main = torch.cuda.current_stream(dev)
            fork = torch.cuda.Event()
            fork.record(main)
            t_host = time.perf_counter()
            self.enqueue_ms_per_group = []
            uploaded = []
            for (plan, names), st in zip(async_plans, self._streams):
                st.wait_event(fork)
                with torch.cuda.stream(st):
                    if host_in is not None:
                        for n in names:
                            self._params[n].data.copy_(host_in[n], non_blocking=True)
                        if remote is not None:
                            ev = torch.cuda.Event()
                            ev.record(st)
                            uploaded.append(ev)
                    plan.enqueue(*args(names))
                    if host_out is not None:
                        for n in names:
                            host_out[n].copy_(self.z[n], non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(st)
                main.wait_event(done)
                self.enqueue_ms_per_group.append((time.perf_counter() - t_host) * 1e3)     # diagnostics: host time so far
            sync_plans = [(pl, nm) for pl, nm in self._plans if not hasattr(pl, 'enqueue')]
            for plan, names in sync_plans:
                self._copy(host_in, names, True)
            if remote is not None:
                if sync_plans:          # their uploads are on the current stream
                    ev = torch.cuda.Event()
                    ev.record(main)
                    uploaded.append(ev)
                self._exchange_weights(remote, uploaded)
            for plan, names in self._plans:
                if hasattr(plan, 'enqueue'):
                    plan.collect()
                else:
                    plan.run(*args(names))
                    self._copy(host_out, names, False)
                self.sweeps.update(plan.sweeps)
            return
        for plan, names in self._plans:
            self._copy(host_in, names, True)
        self._exchange_weights(remote)
        for plan, names in self._plans:
            plan.run(*args(names))
            self._copy(host_out, names, False)
            self.sweeps.update(plan.sweeps)

    def _copy(self, host, names, to_device):
        if host is None:
            return
        for n in names:
            if to_device:
                self._params[n].data.copy_(host[n], non_blocking=True)
            else:
                host[n].copy_(self.z[n], non_blocking=True)

    def append_admm_loss(self, loss):
        if not self._names:
            return loss
        params = [self._params[n] for n in self._names]
        return _PenaltyFn.apply(loss, self, *params)

    def adjust_rho(self, epoch, epochs, factor=5):
        if epoch > int(0.85 * epochs):
            self.rho = factor * self.init_rho

    # -- additive: checkpointing of the ADMM state (the reference drops u/z on resume) -------------
    def state_dict(self):
        return {'rho': self.rho, 'init_rho': self.init_rho,
                'u': {n: t.detach().clone() for n, t in self.u.items()},
                'z': {n: t.detach().clone() for n, t in self.z.items()}}

    def load_state_dict(self, state):
        self.rho = state['rho']
        self.init_rho = state['init_rho']
        for n in self._names:
            self.u[n].copy_(state['u'][n])
            self.z[n].copy_(state['z'][n])

    # -- single-layer helpers with the reference's numpy conventions (admm.py:91-149) --------------
    def _project_one(self, v, layer, kind='tt'):
        dev = self._state_device()
        vt = torch.as_tensor(np.ascontiguousarray(v), dtype=torch.float32).to(dev).contiguous()
        zt = torch.empty_like(vt)
        plan = (projector.TTProjectionPlan if kind == 'tt' else projector.TKProjectionPlan)([layer], dev)
        plan.run([vt], [None], [zt])
        return zt.cpu().numpy()

    def prune_conv_rank_tt(self, z, name):
        L = projector.TTLayer(name, z.shape, self.hp_dict.tt_shapes[name], self.hp_dict.ranks[name])
        try:
            self.hp_dict.ranks[name][:] = L.ranks
        except TypeError:
            pass
        return self._project_one(z, L)

    def prune_linear_rank_tt(self, z, name):
        L = projector.TTLayer(name, z.shape, self.hp_dict.tt_shapes[name], list(self.hp_dict.ranks[name]))
        return self._project_one(z, L)

    def prune_conv_rank_tk(self, z, name):
        return self._project_one(z, projector.TKLayer(name, z.shape, self.hp_dict.ranks[name]), 'tk')

    def prune_linear_rank_tk(self, z, name):
        return self._project_one(z, projector.TKLayer(name, z.shape, self.hp_dict.ranks[name]), 'tk')

    def prune_conv_rank_svd(self, z, name):
        r = _svd_rank(self.hp_dict.ranks[name])
        m = z.detach().squeeze().cpu().numpy()
        o, i = m.shape
        out = self._project_one(m, projector.TTLayer(name, (o, i), [o, i], [1, r, 1]))
        return out[:, :, None, None]

    def prune_linear_rank_svd(self, z, name):
        r = _svd_rank(self.hp_dict.ranks[name])
        m = z.detach().cpu().numpy()
        o, i = m.shape
        return self._project_one(m, projector.TTLayer(name, (o, i), [o, i], [1, r, 1]))
