"""Host-side plans for the batched low-rank projections (Z-update of admm.py:42-69).

A plan is built once per (model, hp table): it owns the workspaces, precomputes every kernel's task
table for every "wave" (TT step i over all layers at once -- the layers are independent, admm.py:43)
and then replays them through the C ABI (`tta_runtime`) on each `run()`.

TT-SVD of one layer (ttd.py:10-31) per step, with A the (r_i*s_i) x n unfolding of the carry:
    m <= n :  G = A A^T  --eigh-->  E = top-r eigenvectors (rows),  core = E^T,  carry' = E A
    m >  n :  G = A^T A  --eigh-->  E = V_r^T, sigma;  core = A E^T diag(1/sigma),  carry' = diag(sigma) E
(the reference computes the same objects from a full LAPACK SVD; only the dominant r-dimensional
singular subspace enters the reconstruction, and that subspace is the dominant eigenspace of G).
Reconstruction (ttd.py:34-43) is the left-to-right chain of GEMMs, its last product written straight
into Z (through the (O,KK,I)->(O,I,KK) fold for k x k convolutions, admm.py:99).
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

import tta_runtime as rt


def _round_up(x, m):
    return (x + m - 1) // m * m


def _prod(xs):
    p = 1
    for v in xs:
        p *= int(v)
    return p


def clip_tt_ranks(shapes, ranks):
    """The in-place clip of ttd.py:18-19, resolved statically: r_{i+1} <- min(r_{i+1}, m, n)."""
    ranks = list(ranks)
    d = len(shapes)
    for i in range(d - 1):
        m = ranks[i] * shapes[i]
        n = _prod(shapes[i + 1:])
        ranks[i + 1] = min(ranks[i + 1], m, n)
    return ranks


def eig_geometry(k, pmax=None):
    """Column-state geometry (ld, kpad, bw) of the Jacobi solver for a k x k Gram matrix.

    k <= 32: one CTA of the column-rotation cluster solver (P = 1, bw = ceil(k/2) rounded to even).
    32 < k <= 512: gram-rotate-apply cluster solver -- blocks of bw = 16 columns, two blocks per CTA,
    P = ceil(k / 32) CTAs per problem (2..16), kpad = 32 * P.
    `pmax` (tests only) selects the geometry of the column-rotation solver for every k <= 512:
    kpad = 2 * P * bw with P the smallest power of two <= pmax such that bw = ceil(k / 2P) <= 16.
    k > 512 (Tucker sweep): multi-launch solver, blocks of 16 (8 if shared memory is short).
    """
    ld = _round_up(k, 4)
    if ld <= 512:
        if pmax is None and k > 32:
            p = (k + 31) // 32
            return ld, 32 * p, 16
        if pmax is None:
            pmax = 1
        p = 1
        while (k + 2 * p - 1) // (2 * p) > 16 and p < pmax:
            p *= 2
        bw = max(2, _round_up((k + 2 * p - 1) // (2 * p), 2))
        return ld, 2 * p * bw, bw
    bw = 16 if 2 * 16 * ld * 4 <= 200 * 1024 else 8
    return ld, _round_up(k, bw), bw


_EIG_TIME_TABLE = ((8, 0.03), (32, 0.10), (64, 0.31), (130, 0.74), (256, 1.45), (512, 4.2), (1024, 40.0), (2048, 320.0))


def eig_time_ms(k):
    """Measured duration of one Jacobi eigensolve on B200 (log-log interpolation; scheduling model only)."""
    tab = _EIG_TIME_TABLE
    if k <= tab[0][0]:
        return tab[0][1]
    for (k0, t0), (k1, t1) in zip(tab[:-1], tab[1:]):
        if k <= k1:
            f = (math.log(k) - math.log(k0)) / (math.log(k1) - math.log(k0))
            return math.exp(math.log(t0) + f * (math.log(t1) - math.log(t0)))
    return tab[-1][1] * (k / tab[-1][0]) ** 3


def eig_ctas(k):
    """CTAs (= SMs) one eigenproblem occupies."""
    return 1 if k <= 32 else min((k + 31) // 32, 16) if k <= 512 else 148


_GPC_SMS = (18, 18, 18, 18, 18, 18, 20, 20)      # B200: 148 SMs in 8 GPCs (a cluster lives inside one GPC)


def wave_makespan_ms(ks):
    """Simulated duration of one eigensolver wave: problems are started largest cluster first (the internal
    streams carry descending priorities) on the first GPC with enough free SMs; one CTA per SM."""
    if not ks:
        return 0.0
    whole = [k for k in ks if eig_ctas(k) > max(_GPC_SMS)]      # whole-GPU solves run one after another
    pending = sorted((k for k in ks if eig_ctas(k) <= max(_GPC_SMS)), reverse=True)
    serial = sum(eig_time_ms(k) for k in whole)
    if not pending:
        return serial
    free = list(_GPC_SMS)
    running = []                 # (end time, gpc, ctas)
    now, end = 0.0, 0.0
    while pending:
        placed = False
        for idx, k in enumerate(pending):
            need = eig_ctas(k)
            g = next((g for g in range(len(free)) if free[g] >= need), None)
            if g is not None:
                free[g] -= need
                t_end = now + eig_time_ms(k)
                running.append((t_end, g, need))
                end = max(end, t_end)
                pending.pop(idx)
                placed = True
                break
        if not placed:
            running.sort()
            t_end, g, need = running.pop(0)
            now = t_end
            free[g] += need
    return end + serial


TRD_MIN_K = 33          # k <= 32 stays on the one-CTA Jacobi solver (eig_cluster.cu)


def default_solver():
    """'trd' (fp64 tridiagonalisation route, csrc/trd.cu) or 'jacobi' (TTA_EIG_SOLVER=jacobi: the round-1 path)."""
    return 'jacobi' if os.environ.get('TTA_EIG_SOLVER', 'trd').lower().startswith('j') else 'trd'


_TRD_MAX_K = []


def uses_trd(k, solver):
    if solver != 'trd' or k < TRD_MIN_K:
        return False
    if not _TRD_MAX_K:
        _TRD_MAX_K.append(rt.symeig_max_k())
    return k <= _TRD_MAX_K[0]


class EigClusterFlag(Exception):
    """The fp64 solver met eigenvalues it cannot separate (status 1 of tta_symeig_task): rerun on the Jacobi solver."""


def refine_window(k, r):
    """Rows of S / T the refinement forms: the r selectable vectors plus a margin that is far wider
    than the ordering error of the fp32 eigenvalue estimates (include/tta.h, tta_refine_task)."""
    return min(k, r + max(16, (r + 7) // 8))


def gram_splits(k, red_len):
    """Slices of the reduction range per Gram problem (one CTA per 128 x 128 lower-triangular tile and slice on the
    tensor-core path, csrc/gram_tc.cu): enough tiles x slices for every SM, at least 256 reduction indices (8 k-blocks)
    per slice, at most 148 slices (the finish kernel sums them in fp64).  The fp64 CUDA-core kernels, which serve
    the operands TMA cannot address, take the same split (their tiles are 32 / 64 / 128 wide)."""
    tiles = (k + 127) // 128
    pairs = tiles * (tiles + 1) // 2
    return max(1, min((red_len + 255) // 256, (148 + pairs - 1) // pairs, 148))


def tt_step_flops(shapes, ranks):
    """SURVEY 8(d) accounting: (gram, proj, eig, recon) FLOPs of one layer (used for LPT sharding)."""
    ranks = clip_tt_ranks(shapes, ranks)
    d = len(shapes)
    gram = proj = eig = recon = 0
    for i in range(d - 1):
        m = ranks[i] * shapes[i]
        n = _prod(shapes[i + 1:])
        k, big = min(m, n), max(m, n)
        r = ranks[i + 1]
        gram += 2 * k * k * big
        proj += 2 * r * m * n * (2 if m > n else 1)
        eig += 9 * k ** 3
    for i in range(1, d):
        recon += 2 * _prod(shapes[:i]) * ranks[i] * shapes[i] * ranks[i + 1]
    return gram, proj, eig, recon


class _Phases:
    """Optional per-phase device timing (CUDA events on the launching stream); used by bench.py only.

    `sink` is None (no timing, no extra syncs) or a dict phase -> accumulated milliseconds.
    """

    def __init__(self, sink):
        self.sink = sink if (sink is not None and not rt.backend_is_emulated()) else None
        self.marks = []

    def mark(self, name):
        if self.sink is None:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.marks.append((name, ev))

    def finish(self):
        if self.sink is None:
            return
        end = torch.cuda.Event(enable_timing=True)
        end.record()
        end.synchronize()
        evs = self.marks + [('end', end)]
        for (name, a), (_, b) in zip(evs[:-1], evs[1:]):
            self.sink[name] = self.sink.get(name, 0.0) + a.elapsed_time(b)


class _Buf:
    """fp32/fp64 workspace tensor + raw address."""

    def __init__(self, numel, device, dtype=torch.float32):
        self.t = torch.empty(max(int(numel), 1), dtype=dtype, device=device)
        self.ptr = self.t.data_ptr()


class TTLayer:
    """Static description of one layer's projection.

    kind: 'conv'   4-D weight (O, I, kh, kw) viewed as (O, KK, I) (admm.py:96)
          'matrix' 2-D weight, tensorised row-major straight to `shapes` (admm.py:107)
    """

    def __init__(self, name, weight_shape, shapes, ranks):
        self.name = name
        self.weight_shape = tuple(int(v) for v in weight_shape)
        self.numel = _prod(self.weight_shape)
        if len(self.weight_shape) == 4:
            o, i, kh, kw = self.weight_shape
            self.O, self.I, self.KK = o, i, kh * kw
        elif len(self.weight_shape) == 2:
            self.O, self.I, self.KK = self.weight_shape[0], self.weight_shape[1], 1
        else:
            raise Exception('ERROR: unsupported layer in ADMM!')
        self.shapes = [int(s) for s in shapes]
        if _prod(self.shapes) != self.numel:
            raise ValueError('tt_shapes {} do not factor weight {} of {}'.format(self.shapes, self.weight_shape, name))
        if len(ranks) != len(self.shapes) + 1:
            raise ValueError('need len(ranks) == len(tt_shapes) + 1 for {}'.format(name))
        self.ranks = clip_tt_ranks(self.shapes, [int(r) for r in ranks])
        self.d = len(self.shapes)


def tt_layer_groups(layers, skip_full_rank=True, max_groups=6):
    """Partition TT layers into groups that run as independent plans on their own CUDA streams.

    One plan executes its waves in lock step (every eigenproblem of a wave must finish before the next
    Gram starts), so layers whose chains have very different lengths hold each other up, and the Gram /
    refinement / projection kernels between two eigensolver phases leave most SMs idle.  Groups are formed by
    (number of real steps, size class of the largest eigenproblem); a class whose big wave needs more than
    one GPU-full of SMs is chunked, and the classes of small problems (k <= 128) are merged.  Returned most
    critical first (longest modelled chain): the caller gives the first groups the high-priority streams.
    """
    def real_ks(L):
        ks = []
        for i in range(L.d - 1):
            m, n = L.ranks[i] * L.shapes[i], _prod(L.shapes[i + 1:])
            k = min(m, n)
            if not (skip_full_rank and L.ranks[i + 1] == k):
                ks.append(k)
        return ks

    mode = os.environ.get('TTA_GROUP_MODE', '')
    if mode.startswith('rr'):
        # experiment: G balanced groups, every class dealt round-robin (same wave structure in every group)
        G = int(mode[2:] or 2)
        order = sorted(range(len(layers)), key=lambda li: (-sum(eig_time_ms(k) * eig_ctas(k) for k in real_ks(layers[li])), li))
        out = [[] for _ in range(G)]
        for q, li in enumerate(order):
            out[q % G].append(li)
        return [sorted(g) for g in out if g]
    classes = {}
    for li, L in enumerate(layers):
        ks = real_ks(L)
        kmax = max(ks, default=0)
        key = (0, 0) if kmax <= 128 else (len(ks), int(math.ceil(math.log2(kmax))))
        classes.setdefault(key, []).append((li, ks))
    groups = []
    for key, members in classes.items():
        chunks = [[]]
        load = 0
        for li, ks in members:
            need = eig_ctas(max(ks, default=1))
            if chunks[-1] and key != (0, 0) and load + need > kNumSMsB200:
                chunks.append([])
                load = 0
            chunks[-1].append((li, ks))
            load += need
        groups.extend(chunks)
    chain = lambda g: max(sum(eig_time_ms(k) for k in ks) for _, ks in g) if g else 0.0
    groups.sort(key=lambda g: -chain(g))
    while len(groups) > max_groups:                       # merge the two least critical groups
        b = groups.pop()
        groups[-1] = groups[-1] + b
    return [[li for li, _ in g] for g in groups if g]


kNumSMsB200 = 148


class TTProjectionPlan:
    """Batched TT-SVD projection of a list of layers: Z_l = Proj_TT(W_l + U_l)."""

    def __init__(self, layers, device, tol=5e-7, max_sweeps=40, refine=True, skip_full_rank=True):
        self.layers = list(layers)
        self.device = torch.device(device)
        self.tol = float(tol)
        self.max_sweeps = int(max_sweeps)
        self.refine = bool(refine)
        # A step that keeps r == min(m, n) singular triplets truncates nothing: U S V^T is the
        # unfolding itself, so the step is served by the exact factorisation A = I * A (m <= n) or
        # A = A * I (m > n) without an eigensolve.  ttd.ten2tt passes skip_full_rank=False to hand out
        # orthonormal cores like the reference's SVD does; the product of the cores is the same.
        self.skip_full_rank = bool(skip_full_rank)
        self.sweeps = {}
        self.profile = None
        self.trace = None
        self._bound = None
        # Warm start: from the second update on, the Jacobi state of a step starts from X = G Q_prev (Q_prev = the
        # eigenvector estimate of the previous update, still in the refinement buffer `qt`) instead of X = G.  Any
        # orthogonal Q is a valid start of the one-sided iteration (it converges to G Q J = V Lambda all the same);
        # when consecutive updates see similar Gram matrices -- ADMM iterates, fine-tuning epochs -- G Q_prev is
        # nearly orthogonal already and the solver needs fewer sweeps.  Results are converged to the same
        # criterion either way.  TTA_WARM_START=0 disables it.
        self.warm_start = bool(refine) and os.environ.get('TTA_WARM_START', '1') != '0'
        # eigensolver of the steps with 32 < k <= tta_symeig_max_k(): 'trd' = fp64 Householder tridiagonalisation +
        # bisection + twisted factorisation (no refinement, no warm start needed), 'jacobi' = round-1 path.  A plan
        # whose trd solve reports inseparable eigenvalues (exactly repeated singular values) switches to 'jacobi'.
        self.solver = default_solver() if refine else 'jacobi'
        self.gram_in_place = os.environ.get('TTA_GRAM_IN_PLACE', '1') != '0'
        self._flags_ev = self._flags_host = self._flags_dev = None
        self._last = None
        self._warm_valid = False
        self._warm_used = False      # the update being collected was warm-started
        self._cold_sweeps = {}       # (wave, slot) -> sweeps of the last cold solve
        self._nowarm = {}            # (wave, slot) -> updates left of a cold-start penalty (the warm start did not pay)
        self._strikes = {}           # (wave, slot) -> consecutive warm updates whose sweep count barely dropped
        self._alloc()

    # -- workspace ---------------------------------------------------------------------------------
    def _alloc(self):
        dev = self.device
        self.ws = []
        self._eyes = {}
        for L in self.layers:
            w = {'T': _Buf(L.numel, dev), 'steps': [], 'acc': {}}
            carry = w['T']
            for i in range(L.d - 1):
                m = L.ranks[i] * L.shapes[i]
                n = _prod(L.shapes[i + 1:])
                k = min(m, n)
                r = L.ranks[i + 1]
                if self.skip_full_rank and r == k:
                    eye = self._eye(k)
                    # m <= n: core = I_m, carry' = A;   m > n: core = A, carry' = I_n
                    st = dict(m=m, n=n, k=k, r=r, identity=True, A=carry,
                              core=eye if m <= n else carry, carry=carry if m <= n else eye)
                    w['steps'].append(st)
                    carry = st['carry']
                    continue
                ld, kpad, bw = eig_geometry(k)
                nsplit = gram_splits(k, max(m, n))
                trd = uses_trd(k, self.solver)
                if trd:
                    ld, kpad, bw = _round_up(k, 4), k, 0
                st = dict(m=m, n=n, k=k, r=r, identity=False, ld=ld, kpad=kpad, bw=bw, nsplit=nsplit, A=carry, trd=trd,
                          X=None if trd else _Buf(ld * kpad, dev), part=_Buf(nsplit * k * k, dev, torch.float64),
                          E=_Buf(r * k, dev), core=_Buf(m * r, dev), carry=_Buf(r * n, dev))
                if m > n:
                    st['sigma'] = _Buf(r, dev)
                    st['isigma'] = _Buf(r, dev)
                if trd:
                    f64 = torch.float64
                    st['g64'] = _Buf(k * k, dev, f64)
                    st['work'] = _Buf(rt.symeig_work_doubles(k, r), dev, f64)
                    st['e64'] = _Buf(r * k, dev, f64)
                    st['lam'] = _Buf(r, dev, f64)
                elif self.refine:
                    f64 = torch.float64
                    wnd = refine_window(k, r)
                    st['wnd'] = wnd
                    st['g64'] = _Buf(k * k, dev, f64)
                    st['qt'] = _Buf(k * k, dev, f64)
                    for nm in ('y', 's', 't'):
                        st[nm] = _Buf(wnd * k, dev, f64)
                    st['c'] = _Buf(r * k, dev, f64)
                    st['e64'] = _Buf(r * k, dev, f64)
                    st['lam'] = _Buf(r, dev, f64)
                    st['lam0'] = _Buf(k, dev, f64)
                w['steps'].append(st)
                carry = st['carry']
            # reconstruction accumulators acc_j, j = 1..d-1 ; acc_0 is core_0
            for j in range(1, L.d):
                rows = _prod(L.shapes[:j + 1])
                w['acc'][j] = _Buf(rows * L.ranks[j + 1], dev)
            self.ws.append(w)

    def _eye(self, k):
        if k not in self._eyes:
            b = _Buf(k * k, self.device)
            b.t[:k * k].view(k, k).copy_(torch.eye(k, dtype=torch.float32, device=self.device))
            self._eyes[k] = b
        return self._eyes[k]

    def _schedule(self):
        """Wave index of every real (non-identity) step.  Steps of one layer keep their order; a layer
        with fewer real steps than the longest one may start at any offset.  Flexible layers are placed,
        costliest first, at the offset that minimises the sum of the simulated wave durations
        (`wave_makespan_ms`: one cluster of P CTAs per eigenproblem, a cluster lives inside one GPC) --
        e.g. the k = 512 problems of the 1x1 convolutions are split between the two waves that hold the big
        middle steps of the k x k convolutions instead of adding a wave-long critical path of their own."""
        real = [[i for i, st in enumerate(w['steps']) if not st['identity']] for w in self.ws]
        nwaves = max((len(r) for r in real), default=0)
        jobs = [[] for _ in range(max(nwaves, 1))]          # k of every eigenproblem of a wave
        ks = lambda li: [self.ws[li]['steps'][i]['k'] for i in real[li]]
        sched = [None] * len(self.ws)
        order = sorted(range(len(self.ws)),
                       key=lambda li: (-len(real[li]), -sum(eig_time_ms(k) * eig_ctas(k) for k in ks(li)), li))
        base = [0.0] * max(nwaves, 1)
        for li in order:
            n = len(real[li])
            if n == 0:
                sched[li] = {}
                continue
            best, best_key = 0, None
            for off in range(nwaves - n + 1):
                total = 0.0
                for wv in range(nwaves):
                    if off <= wv < off + n:
                        total += wave_makespan_ms(jobs[wv] + [ks(li)[wv - off]])
                    else:
                        total += base[wv]
                # ties: prefer the centre of the sequence (the long middle waves)
                key = (round(total, 4), abs(2 * off + n - nwaves))
                if best_key is None or key < best_key:
                    best, best_key = off, key
            sched[li] = {i: best + q for q, i in enumerate(real[li])}
            for q, k in enumerate(ks(li)):
                jobs[best + q].append(k)
                base[best + q] = wave_makespan_ms(jobs[best + q])
        return nwaves, sched

    def max_order(self):
        return max((L.d for L in self.layers), default=0)

    # -- task tables -------------------------------------------------------------------------------
    def bind(self, w_list, u_list, z_list):
        """(Re)build the task tables for concrete W / U / Z tensors (raw pointers are baked in)."""
        key = tuple((w.data_ptr(), (u.data_ptr() if u is not None else 0), z.data_ptr())
                    for w, u, z in zip(w_list, u_list, z_list))
        if key == self._bound:
            return
        dev = self.device
        for t in list(w_list) + [u for u in u_list if u is not None] + list(z_list):
            rt.require_device(t)
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise rt.TtaError('W/U/Z must be contiguous fp32 tensors')
        nL = len(self.layers)

        unfold = np.zeros(nL, dtype=rt.FOLD_TASK)
        for li, (L, w) in enumerate(zip(self.layers, self.ws)):
            unfold[li] = (w_list[li].data_ptr(), u_list[li].data_ptr() if u_list[li] is not None else 0,
                          w['T'].ptr, 0, L.O, L.I, L.KK, 0)
        self.t_unfold = rt.TaskTable(unfold, dev)

        self.waves = []
        nwaves, sched = self._schedule()
        for wv in range(nwaves):
            members = [(li, i) for li in range(nL) for i, w_ in sched[li].items() if w_ == wv]
            steps = [self.ws[li]['steps'][si] for li, si in members]
            idx = [li for li, _ in members]
            jq = [q for q, st in enumerate(steps) if not st['trd']]        # Jacobi (+ refinement) members
            tq = sorted((q for q, st in enumerate(steps) if st['trd']), key=lambda q: -steps[q]['k'])   # fp64 solver
            g = np.zeros(len(idx), dtype=rt.GRAM_TASK)
            mm = np.zeros(len(idx), dtype=rt.GEMM_TASK)
            sel = np.zeros(len(idx), dtype=rt.SELECT_TASK)
            for q, st in enumerate(steps):
                m, n, k, r = st['m'], st['n'], st['k'], st['r']
                a = st['A'].ptr
                x = st['X'].ptr if st['X'] is not None else 0
                g64 = st['g64'].ptr if 'g64' in st else 0
                # The Gram matrix of the first step is formed from W and U IN PLACE (tta_gram_task.a2): a row Gram does
                # not depend on the order of the reduction index, so the (O, I, KK) -> (O, KK, I) permute of admm.py:96
                # is irrelevant to it as long as a row of the unfolding is made of whole output channels; without a
                # spatial extent (KK == 1) there is no permute at all.  T = unfold(W + U) feeds the projection GEMM only.
                li, si = members[q]
                L = self.layers[li]
                ga, ga2 = a, 0
                if si == 0 and self.gram_in_place and (L.KK == 1 or (m <= n and n % (L.I * L.KK) == 0)):
                    ga = w_list[li].data_ptr()
                    ga2 = u_list[li].data_ptr() if u_list[li] is not None else 0
                if m <= n:   # row Gram A A^T
                    g[q] = (ga, st['part'].ptr, x, g64, n, 0, 1, k, 1, n, st['nsplit'], st['ld'], st['kpad'], ga2)
                    sel[q] = (x, st['E'].ptr, st['core'].ptr, 0, 0, 0, k, st['ld'], r, 0)
                    # carry' (r x n) = E (r x m) * A (m x n)
                    mm[q] = (st['E'].ptr, a, st['carry'].ptr, 0, m, 1, n, 1, n, r, n, m, 0)
                else:        # column Gram A^T A
                    g[q] = (ga, st['part'].ptr, x, g64, 1, 0, n, k, 1, m, st['nsplit'], st['ld'], st['kpad'], ga2)
                    sel[q] = (x, st['E'].ptr, 0, st['carry'].ptr, st['sigma'].ptr, st['isigma'].ptr,
                              k, st['ld'], r, 0)
                    # core (m x r) = A (m x n) * E^T (n x r) * diag(1/sigma)
                    mm[q] = (a, st['E'].ptr, st['core'].ptr, st['isigma'].ptr, n, 1, 1, n, r, m, r, n, 0)
            # -- Jacobi members --
            nj = len(jq)
            e = np.zeros(nj, dtype=rt.EIG_TASK)
            rf = np.zeros(nj, dtype=rt.REFINE_TASK)
            dg = [np.zeros(nj, dtype=rt.GEMM_TASK) for _ in range(4)]
            wg = np.zeros(nj, dtype=rt.GEMM_TASK)
            for jj, q in enumerate(jq):
                st = steps[q]
                k, r = st['k'], st['r']
                e[jj] = (st['X'].ptr, k, st['ld'], st['kpad'], st['bw'])
                if self.refine:
                    sl = sel[q]
                    wnd = st['wnd']
                    rf[jj] = (st['X'].ptr, st['qt'].ptr, st['s'].ptr, st['t'].ptr, st['c'].ptr, st['lam'].ptr,
                              st['lam0'].ptr, st['e64'].ptr, sl['e'], sl['et'], sl['se'], sl['sigma'],
                              sl['isigma'], k, st['ld'], r, wnd)
                    qt = st['qt'].ptr
                    dg[0][jj] = (qt, st['g64'].ptr, st['y'].ptr, 0, k, 1, k, 1, k, wnd, k, k, 0)    # y = qt[:wnd] * g64
                    dg[1][jj] = (st['y'].ptr, qt, st['s'].ptr, 0, k, 1, 1, k, k, wnd, k, k, 0)      # s = y * qt^T
                    dg[2][jj] = (qt, qt, st['t'].ptr, 0, k, 1, 1, k, k, wnd, k, k, 0)               # t = qt[:wnd] * qt^T
                    dg[3][jj] = (st['c'].ptr, qt, st['e64'].ptr, 0, k, 1, k, 1, k, r, k, k, 0)      # e64 = c * qt
                    # warm start: X (column j at x + j*ld, fp32) = G q_j, i.e. rows of qt * g64 (G symmetric)
                    # guarded by the smallest eigenvalue estimate of the previous solve (lam0 is sorted, descending):
                    # a null column is stored as a zero row of qt, and a rank-deficient start could not recover
                    # directions that become non-null later -- such a problem starts cold (X = G stays in place)
                    wg[jj] = (qt, st['g64'].ptr, st['X'].ptr, st['lam0'].ptr + 8 * (k - 1), k, 1, k, 1, st['ld'], k, k, k,
                              rt.GEMM_STORE_F32 | rt.GEMM_GUARD)
            # -- fp64 tridiagonalisation members (sorted by descending k: equal cluster sizes are adjacent) --
            nt = len(tq)
            sy = np.zeros(nt, dtype=rt.SYMEIG_TASK)
            fin = np.zeros(nt, dtype=rt.REFINE_TASK)
            status = torch.zeros(max(nt, 1), dtype=torch.int32, device=dev)
            for tt, q in enumerate(tq):
                st = steps[q]
                sl = sel[q]
                sy[tt] = (st['g64'].ptr, st['work'].ptr, st['lam'].ptr, st['e64'].ptr, status.data_ptr() + 4 * tt,
                          st['k'], st['r'])
                fin[tt] = (0, 0, 0, 0, 0, st['lam'].ptr, 0, st['e64'].ptr, sl['e'], sl['et'], sl['se'], sl['sigma'],
                           sl['isigma'], st['k'], st['ld'], st['r'], st['r'])
            wave = dict(idx=idx, jidx=[idx[q] for q in jq], tidx=[idx[q] for q in tq],
                        gram=rt.TaskTable(g, dev), eig=rt.TaskTable(e, dev),
                        select=rt.TaskTable(sel[jq] if nj else sel[:0], dev), gemm=rt.TaskTable(mm, dev),
                        symeig=rt.TaskTable(sy, dev), symfin=rt.TaskTable(fin, dev), status=status)
            if self.refine:
                wave['refine'] = rt.TaskTable(rf, dev)
                wave['dgemm_ys_t'] = rt.TaskTable(np.concatenate([dg[0], dg[2]]), dev)   # independent: one launch
                wave['dgemm_s'] = rt.TaskTable(dg[1], dev)
                wave['dgemm_e'] = rt.TaskTable(dg[3], dev)
                wave['warm_rows'] = wg
                keep = [q for q in range(nj) if (wv, q) not in self._nowarm]
                wave['warm_q'] = keep
                wave['warm'] = rt.TaskTable(wg[keep], dev)
            nbytes = rt.jacobi_scratch_bytes(wave['eig']) if nj else 0
            wave['scratch'] = torch.empty(max(nbytes // 4 + 16, 16), dtype=torch.int32, device=dev)
            self.waves.append(wave)

        # reconstruction chain: recon wave j multiplies acc_{j-1} by core_j (or the last carry)
        self.recon = []
        fold = []
        for j in range(1, self.max_order()):
            idx = [li for li, L in enumerate(self.layers) if L.d > j]
            mm = np.zeros(len(idx), dtype=rt.GEMM_TASK)
            for q, li in enumerate(idx):
                L, w = self.layers[li], self.ws[li]
                rows = _prod(L.shapes[:j])
                rj, rj1, sj = L.ranks[j], L.ranks[j + 1], L.shapes[j]
                left = w['steps'][0]['core'].ptr if j == 1 else w['acc'][j - 1].ptr
                last = (j == L.d - 1)
                right = w['steps'][j - 1]['carry'].ptr if last else w['steps'][j]['core'].ptr
                out = w['acc'][j].ptr
                if last and L.KK == 1:
                    out = z_list[li].data_ptr()      # no fold needed: (O, 1, I) == (O, I)
                ncols = sj * rj1
                mm[q] = (left, right, out, 0, rj, 1, ncols, 1, ncols, rows, ncols, rj, 0)
                if last and L.KK != 1:
                    fold.append((0, 0, w['acc'][j].ptr, z_list[li].data_ptr(), L.O, L.I, L.KK, 0))
            self.recon.append(rt.TaskTable(mm, dev))
        self.t_fold = rt.TaskTable(np.array(fold, dtype=rt.FOLD_TASK) if fold else np.zeros(0, dtype=rt.FOLD_TASK), dev)
        self._bound = key

    # -- execution ---------------------------------------------------------------------------------
    def run(self, w_list, u_list, z_list):
        self.enqueue(w_list, u_list, z_list)
        self.collect()

    def enqueue(self, w_list, u_list, z_list):
        """Enqueue the whole projection on the current stream; no host synchronisation."""
        self.bind(w_list, u_list, z_list)
        self._last = (w_list, u_list, z_list)
        ph = _Phases(self.profile)
        tr = self.trace            # None, or a list that receives (label, CUDA event) marks of this stream (diagnostics)

        def mark(label):
            if tr is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                tr.append((label, ev))
        warm = self.warm_start and self.refine and self._warm_valid
        self._warm_valid = False            # set again by collect() once this update is known to have converged
        self._warm_used = warm
        mark('start')
        ph.mark('unfold')
        rt.unfold_add(self.t_unfold)
        for wi, wave in enumerate(self.waves):
            ph.mark('gram')
            rt.gram(wave['gram'])
            mark('w{} eig begin'.format(wi))
            ph.mark('eig')
            if wave['symeig'].n:
                rt.symeig_top(wave['symeig'])
            if wave['eig'].n:
                if warm and wave['warm'].n:
                    rt.gemm_f64(wave['warm'])
                rt.jacobi_eigh_async(wave['eig'], wave['scratch'], self.tol, self.max_sweeps)
            mark('w{} eig end'.format(wi))
            ph.mark('select')
            if wave['symeig'].n:
                rt.refine_finalize(wave['symfin'])
            if wave['eig'].n:
                if self.refine:
                    rt.refine_prepare(wave['refine'])
                    rt.gemm_f64(wave['dgemm_ys_t'])
                    rt.gemm_f64(wave['dgemm_s'])
                    rt.refine_coeff(wave['refine'])
                    rt.gemm_f64(wave['dgemm_e'])
                    rt.refine_finalize(wave['refine'])
                else:
                    rt.select(wave['select'])
            ph.mark('project')
            rt.gemm(wave['gemm'])
        ph.mark('recon')
        for tab in self.recon:
            rt.gemm(tab)
        ph.mark('fold')
        if self.t_fold.n:
            rt.fold_store(self.t_fold)
        self._enqueue_flag_readback()
        mark('end')
        ph.finish()

    def _flag_parts(self):
        parts = []
        for w in self.waves:
            if w['eig'].n or w['symeig'].n:
                parts.append(w['scratch'][:6 * w['eig'].n])
                parts.append(w['status'][:w['symeig'].n])
        return parts

    def _enqueue_flag_readback(self):
        """Sweep counts / status words of every wave -> pinned host memory, on the plan's stream right behind its last
        kernel: collect() then only waits for an event instead of issuing a cat + D2H copy per plan after the join."""
        parts = self._flag_parts()
        if not parts or rt.backend_is_emulated():
            self._flags_ev = None
            return
        n = sum(int(p.numel()) for p in parts)
        if self._flags_host is None or self._flags_host.numel() != n:
            self._flags_dev = torch.empty(n, dtype=torch.int32, device=self.device)
            self._flags_host = torch.empty(n, dtype=torch.int32, pin_memory=True)
        torch.cat(parts, out=self._flags_dev)
        self._flags_host.copy_(self._flags_dev, non_blocking=True)
        self._flags_ev = torch.cuda.Event()
        self._flags_ev.record()

    def collect(self):
        """The one host synchronisation of an update: sweep counts / convergence status of every wave."""
        self.sweeps = {}
        live = [w for w in self.waves if w['eig'].n or w['symeig'].n]
        if not live:
            return
        if self._flags_ev is not None:
            self._flags_ev.synchronize()              # the copy was enqueued behind the last kernel of the plan
            flat = self._flags_host.numpy()
            self._flags_ev = None
        else:
            flat = torch.cat(self._flag_parts()).cpu().numpy()     # one D2H copy per plan
        off = 0
        flagged = False
        for wi, wave in enumerate(self.waves):
            nj, nt = wave['eig'].n, wave['symeig'].n
            if not (nj or nt):
                continue
            n6 = 6 * nj
            sw = rt.jacobi_results_from_host(flat[off:off + n6], wave['eig'], self.max_sweeps) if nj else []
            off += n6
            flagged = flagged or bool(np.any(flat[off:off + nt] != 0))
            off += nt
            for q, li in enumerate(wave['jidx']):
                self.sweeps.setdefault(self.layers[li].name, []).append(int(sw[q]))
            for li in wave['tidx']:
                self.sweeps.setdefault(self.layers[li].name, []).append(0)
            if self.refine and nj:
                # adaptive warm start: the G * Q_prev product costs 2 k^3 fp64 flops per problem; a problem whose
                # sweep count barely drops (unstable eigenvectors: flat spectrum, basis changes upstream in the TT
                # chain) goes back to the cold start: two such updates in a row -> 16 cold updates, then it is probed again
                warmed = set(wave['warm_q']) if self._warm_used else set()
                dirty = False
                for q in range(nj):
                    key = (wi, q)
                    if q not in warmed:
                        self._cold_sweeps[key] = int(sw[q])
                        if key in self._nowarm:
                            self._nowarm[key] -= 1
                            if self._nowarm[key] <= 0:
                                del self._nowarm[key]
                                dirty = True
                    elif int(sw[q]) > 0.7 * self._cold_sweeps.get(key, 1 << 30):
                        self._strikes[key] = self._strikes.get(key, 0) + 1
                        if self._strikes[key] >= 2:
                            self._strikes[key] = 0
                            self._nowarm[key] = 16
                            dirty = True
                    else:
                        self._strikes[key] = 0
                if dirty:
                    keep = [q for q in range(nj) if (wi, q) not in self._nowarm]
                    wave['warm_q'] = keep
                    wave['warm'] = rt.TaskTable(wave['warm_rows'][keep], self.device)
        if flagged:
            # exactly repeated singular values (structured weights): the twisted factorisation may have returned
            # parallel vectors for them.  This plan moves to the Jacobi solver for good and redoes the update.
            self.solver = 'jacobi'
            self._bound = None
            self._cold_sweeps, self._nowarm, self._strikes = {}, {}, {}
            self._alloc()
            self.enqueue(*self._last)
            return self.collect()
        self._warm_valid = True

    def cores(self, li):
        """Cores of layer `li` after `run()` as tensors shaped (r_i, s_i, r_{i+1}) (ten2tt's return)."""
        L, w = self.layers[li], self.ws[li]
        out = []
        for i in range(L.d - 1):
            st = w['steps'][i]
            out.append(st['core'].t[:st['m'] * st['r']].view(L.ranks[i], L.shapes[i], L.ranks[i + 1]))
        last = w['steps'][-1]['carry'] if L.d > 1 else w['T']
        out.append(last.t[:L.ranks[L.d - 1] * L.shapes[L.d - 1] * L.ranks[L.d]]
                   .view(L.ranks[L.d - 1], L.shapes[L.d - 1], L.ranks[L.d]))
        return out


# ==================================================================================================
# Tucker-2 (HOOI) projection                                   admm.py:113-127 -> tensorly partial_tucker
# ==================================================================================================
class EigBatch:
    """Gram -> dominant-r eigenvectors for a list of independent problems: the fp64 tridiagonalisation solver
    (33 <= k <= tta_symeig_max_k()) or Jacobi -> fp64 refinement -> selection (everything else).

    Each problem: dict(a=ptr, k=, r=, si=, sb=, sc=, nb=, nc=) describing the Gram operand as in
    include/tta.h (tta_gram_task); outputs E (r x k, rows = dominant eigenvectors) and optionally ET.
    """

    def __init__(self, device, tol=5e-7, max_sweeps=40, solver=None):
        self.device = device
        self.tol = tol
        self.max_sweeps = max_sweeps
        self.warm_start = os.environ.get('TTA_WARM_START', '1') != '0'
        self.solver = solver or default_solver()
        self.bufs = {}
        self.tables = {}

    def buffers(self, key, k, r, red_len, want_et):
        if key in self.bufs:
            return self.bufs[key]
        dev = self.device
        nsplit = gram_splits(k, red_len)
        f64 = torch.float64
        if uses_trd(k, self.solver):
            b = dict(k=k, r=r, ld=_round_up(k, 4), kpad=k, bw=0, nsplit=nsplit, trd=True, X=None,
                     part=_Buf(nsplit * k * k, dev, f64), E=_Buf(r * k, dev),
                     ET=_Buf(r * k, dev) if want_et else None, g64=_Buf(k * k, dev, f64),
                     work=_Buf(rt.symeig_work_doubles(k, r), dev, f64), e64=_Buf(r * k, dev, f64),
                     lam=_Buf(r, dev, f64))
        else:
            ld, kpad, bw = eig_geometry(k)
            wnd = refine_window(k, r)
            b = dict(k=k, r=r, ld=ld, kpad=kpad, bw=bw, nsplit=nsplit, wnd=wnd, trd=False, X=_Buf(ld * kpad, dev),
                     part=_Buf(nsplit * k * k, dev, f64), E=_Buf(r * k, dev),
                     ET=_Buf(r * k, dev) if want_et else None,
                     g64=_Buf(k * k, dev, f64), qt=_Buf(k * k, dev, f64), y=_Buf(wnd * k, dev, f64),
                     s=_Buf(wnd * k, dev, f64), t=_Buf(wnd * k, dev, f64), c=_Buf(r * k, dev, f64),
                     e64=_Buf(r * k, dev, f64), lam=_Buf(r, dev, f64), lam0=_Buf(k, dev, f64))
        self.bufs[key] = b
        return b

    def build(self, sig, problems):
        """problems: list of (bufs, gram operand dict).  Cached by `sig` (hashable)."""
        if sig in self.tables:
            return self.tables[sig]
        dev = self.device
        n = len(problems)
        g = np.zeros(n, dtype=rt.GRAM_TASK)
        for q, (b, op) in enumerate(problems):
            g[q] = (op['a'], b['part'].ptr, b['X'].ptr if b['X'] is not None else 0, b['g64'].ptr, op['si'], op['sb'],
                    op['sc'], b['k'], op['nb'], op['nc'], b['nsplit'], b['ld'], b['kpad'], op.get('a2', 0))
        jq = [q for q, (b, _) in enumerate(problems) if not b['trd']]
        tq = sorted((q for q, (b, _) in enumerate(problems) if b['trd']), key=lambda q: -problems[q][0]['k'])
        nj, nt = len(jq), len(tq)
        e = np.zeros(nj, dtype=rt.EIG_TASK)
        rf = np.zeros(nj, dtype=rt.REFINE_TASK)
        dg = [np.zeros(nj, dtype=rt.GEMM_TASK) for _ in range(4)]
        wg = np.zeros(nj, dtype=rt.GEMM_TASK)
        for jj, q in enumerate(jq):
            b = problems[q][0]
            k, r = b['k'], b['r']
            e[jj] = (b['X'].ptr, k, b['ld'], b['kpad'], b['bw'])
            wnd = b['wnd']
            rf[jj] = (b['X'].ptr, b['qt'].ptr, b['s'].ptr, b['t'].ptr, b['c'].ptr, b['lam'].ptr, b['lam0'].ptr,
                      b['e64'].ptr, b['E'].ptr, b['ET'].ptr if b['ET'] is not None else 0, 0, 0, 0, k, b['ld'], r, wnd)
            qt = b['qt'].ptr
            dg[0][jj] = (qt, b['g64'].ptr, b['y'].ptr, 0, k, 1, k, 1, k, wnd, k, k, 0)
            dg[1][jj] = (b['y'].ptr, qt, b['s'].ptr, 0, k, 1, 1, k, k, wnd, k, k, 0)
            dg[2][jj] = (qt, qt, b['t'].ptr, 0, k, 1, 1, k, k, wnd, k, k, 0)
            dg[3][jj] = (b['c'].ptr, qt, b['e64'].ptr, 0, k, 1, k, 1, k, r, k, k, 0)
            # warm start X = G Q_prev (see TTProjectionPlan), guarded by the smallest previous eigenvalue estimate
            wg[jj] = (qt, b['g64'].ptr, b['X'].ptr, b['lam0'].ptr + 8 * (k - 1), k, 1, k, 1, b['ld'], k, k, k,
                      rt.GEMM_STORE_F32 | rt.GEMM_GUARD)
        sy = np.zeros(nt, dtype=rt.SYMEIG_TASK)
        fin = np.zeros(nt, dtype=rt.REFINE_TASK)
        status = torch.zeros(max(nt, 1), dtype=torch.int32, device=dev)
        for tt, q in enumerate(tq):
            b = problems[q][0]
            sy[tt] = (b['g64'].ptr, b['work'].ptr, b['lam'].ptr, b['e64'].ptr, status.data_ptr() + 4 * tt, b['k'], b['r'])
            fin[tt] = (0, 0, 0, 0, 0, b['lam'].ptr, 0, b['e64'].ptr, b['E'].ptr, b['ET'].ptr if b['ET'] is not None else 0,
                       0, 0, 0, b['k'], b['ld'], b['r'], b['r'])
        tabs = dict(n=n, jq=jq, tq=tq, gram=rt.TaskTable(g, dev), eig=rt.TaskTable(e, dev), refine=rt.TaskTable(rf, dev),
                    d_yt=rt.TaskTable(np.concatenate([dg[0], dg[2]]), dev), d_s=rt.TaskTable(dg[1], dev),
                    d_e=rt.TaskTable(dg[3], dev), warm=rt.TaskTable(wg, dev),
                    symeig=rt.TaskTable(sy, dev), symfin=rt.TaskTable(fin, dev), status=status,
                    big=any(problems[q][0]['k'] > 512 for q in jq))
        nbytes = rt.jacobi_scratch_bytes(tabs['eig']) if nj else 0
        tabs['scratch'] = torch.empty(max(nbytes // 4 + 16, 16), dtype=torch.int32, device=dev)
        self.tables[sig] = tabs
        return tabs

    def _refine(self, tabs):
        rt.refine_prepare(tabs['refine'])
        rt.gemm_f64(tabs['d_yt'])
        rt.gemm_f64(tabs['d_s'])
        rt.refine_coeff(tabs['refine'])
        rt.gemm_f64(tabs['d_e'])
        rt.refine_finalize(tabs['refine'])

    def enqueue(self, tabs, warm=False):
        """Same as `run` without the host synchronisation; sweep counts via `results(tabs)` after a sync.
        warm: start the Jacobi iteration from G * (eigenvectors of the previous solve in the same buffers) -- the
        Gram matrices of successive HOOI sweeps differ less and less, so the later sweeps need 2-3 Jacobi sweeps."""
        rt.gram(tabs['gram'])
        if tabs['symeig'].n:
            rt.symeig_top(tabs['symeig'])
            rt.refine_finalize(tabs['symfin'])
        if not tabs['eig'].n:
            return
        if warm and self.warm_start:
            rt.gemm_f64(tabs['warm'])
        if tabs['big']:      # k > 512: the multi-launch solver synchronises per sweep and reports through the call
            tabs['_sweeps'] = rt.jacobi_eigh(tabs['eig'], tabs['scratch'], self.tol, self.max_sweeps)
        else:
            rt.jacobi_eigh_async(tabs['eig'], tabs['scratch'], self.tol, self.max_sweeps)
        self._refine(tabs)

    def _merge(self, tabs, jac_sweeps):
        """Sweep counts in problem order (0 for the fp64 solver); raises EigClusterFlag on a reported cluster."""
        if tabs['symeig'].n:
            if bool(tabs['status'][:tabs['symeig'].n].ne(0).any().item()):
                raise EigClusterFlag()
        out = np.zeros(tabs['n'], dtype=np.int32)
        for jj, q in enumerate(tabs['jq']):
            out[q] = int(jac_sweeps[jj])
        return out

    def snapshot_flags(self, *tabs_list):
        """Device-side copy of what `results` would check (fp64 solver status words, Jacobi convergence words) for solves
        whose buffers are about to be reused without a host synchronisation in between."""
        snap = []
        for tabs in tabs_list:
            if tabs['symeig'].n:
                snap.append(('status', tabs, tabs['status'][:tabs['symeig'].n].clone()))
            if tabs['eig'].n and not tabs['big']:
                snap.append(('jacobi', tabs, tabs['scratch'][:6 * tabs['eig'].n].clone()))
        return snap

    def check_flags(self, snapshots):
        for snap in snapshots:
            for kind, tabs, t in snap:
                if kind == 'status':
                    if bool(t.ne(0).any().item()):
                        raise EigClusterFlag()
                else:
                    rt.jacobi_results_from_host(t.cpu().numpy(), tabs['eig'], self.max_sweeps)

    def results(self, tabs):
        if not tabs['eig'].n:
            return self._merge(tabs, [])
        if tabs['big']:
            return self._merge(tabs, tabs['_sweeps'])
        return self._merge(tabs, rt.jacobi_results(tabs['eig'], tabs['scratch'], self.max_sweeps))

    def run(self, tabs):
        rt.gram(tabs['gram'])
        if tabs['symeig'].n:
            rt.symeig_top(tabs['symeig'])
            rt.refine_finalize(tabs['symfin'])
        sweeps = []
        if tabs['eig'].n:
            sweeps = rt.jacobi_eigh(tabs['eig'], tabs['scratch'], self.tol, self.max_sweeps)
            self._refine(tabs)
        return self._merge(tabs, sweeps)


class TKLayer:
    """Tucker-2 projection of a conv (O, I, kh, kw) or linear (O, I) weight on modes (0, 1).

    ranks = [r_out, r_in] as in hp_dicts/tk_*.py (admm.py:115,123 pass them as `rank=`).
    """

    def __init__(self, name, weight_shape, ranks):
        self.name = name
        self.weight_shape = tuple(int(v) for v in weight_shape)
        if len(self.weight_shape) == 4:
            self.O, self.I, self.KK = self.weight_shape[0], self.weight_shape[1], self.weight_shape[2] * self.weight_shape[3]
        elif len(self.weight_shape) == 2:
            self.O, self.I, self.KK = self.weight_shape[0], self.weight_shape[1], 1
        else:
            raise Exception('ERROR: unsupported layer in ADMM!')
        self.numel = self.O * self.I * self.KK
        if isinstance(ranks, int) or len(ranks) != 2:
            raise ValueError('Tucker-2 needs ranks [r_out, r_in] for {}'.format(name))
        # A truncated SVD returns at most min(shape) vectors (tensorly slices U[:, :rank]): the HOSVD initialisation
        # clips r_out to min(O, I*KK) and r_in to min(I, O*KK), and every HOOI half-sweep clips the factor it
        # refreshes to the width of the partially projected unfolding (O x r_in*KK, I x r_out*KK) -- e.g. a linear
        # layer with ranks [30, 20] ends with factors of 20 columns.  The fixed point is reached statically here.
        r0 = min(int(ranks[0]), self.O, self.I * self.KK)
        r1 = min(int(ranks[1]), self.I, self.O * self.KK)
        while True:
            n0 = min(r0, r1 * self.KK)
            n1 = min(r1, n0 * self.KK)
            if (n0, n1) == (r0, r1):
                break
            r0, r1 = n0, n1
        self.r0, self.r1 = r0, r1


class TKProjectionPlan:
    """Batched Tucker-2 HOOI projection: HOSVD initialisation + alternating sweeps with tensorly's
    stopping rule (`iteration > 1 and |err[-2] - err[-1]| < tol`, n_iter_max = 100, tol = 1e-4), every
    layer stopping at its own sweep.  All contractions are plain 2-D GEMMs / Grams because the
    tensor is kept in the (O, KK, I) layout produced by the unfold kernel:

        mode-1 products contract the fastest index   (O*KK x I) . E1^T
        mode-0 products contract the slowest index   E0 . (O x KK*I)
    """

    def __init__(self, layers, device, tol=5e-7, max_sweeps=40, n_iter_max=100, hooi_tol=10e-5):
        self.layers = list(layers)
        self.device = torch.device(device)
        self.n_iter_max = int(n_iter_max)
        self.hooi_tol = float(hooi_tol)
        self.tol, self.max_sweeps = tol, max_sweeps
        self.sweeps = {}
        self.hooi_sweeps = {}
        self.errors = {}
        self.profile = None
        self.gram_in_place = os.environ.get('TTA_GRAM_IN_PLACE', '1') != '0'
        self._alloc(None)

    def _alloc(self, solver):
        self.eig = EigBatch(self.device, self.tol, self.max_sweeps, solver)
        self._bound = None
        dev = self.device
        self.ws = []
        for li, L in enumerate(self.layers):
            O, I, KK, r0, r1 = L.O, L.I, L.KK, L.r0, L.r1
            w = dict(T=_Buf(L.numel, dev), P0=_Buf(O * KK * r1, dev), P1=_Buf(r0 * KK * I, dev),
                     core=_Buf(r0 * KK * r1, dev), tmp=_Buf(r0 * KK * I, dev), ZT=_Buf(L.numel, dev))
            w['e0'] = self.eig.buffers((li, 0), O, r0, max(KK * I, O), False)
            w['e1'] = self.eig.buffers((li, 1), I, r1, max(O * KK, I), False)
            self.ws.append(w)
        self.norms = torch.zeros(4 * max(len(self.layers), 1), dtype=torch.float64, device=dev)     # |X|^2, 3 sweeps of |core|^2

    # -- static per-layer task rows ---------------------------------------------------------------
    def bind(self, w_list, u_list, z_list):
        key = tuple((w.data_ptr(), (u.data_ptr() if u is not None else 0), z.data_ptr())
                    for w, u, z in zip(w_list, u_list, z_list))
        if key == self._bound:
            return
        for t in list(w_list) + [u for u in u_list if u is not None] + list(z_list):
            rt.require_device(t)
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise rt.TtaError('W/U/Z must be contiguous fp32 tensors')
        n = len(self.layers)
        self.row = dict(unfold=np.zeros(n, dtype=rt.FOLD_TASK), fold=np.zeros(n, dtype=rt.FOLD_TASK),
                        p0=np.zeros(n, dtype=rt.GEMM_TASK), p1=np.zeros(n, dtype=rt.GEMM_TASK),
                        core=np.zeros(n, dtype=rt.GEMM_TASK), rec1=np.zeros(n, dtype=rt.GEMM_TASK),
                        rec2=np.zeros(n, dtype=rt.GEMM_TASK), nx=np.zeros(n, dtype=rt.SQNORM_TASK),
                        nc=np.zeros(n, dtype=rt.SQNORM_TASK))
        self.gram_ops = []
        for li, (L, w) in enumerate(zip(self.layers, self.ws)):
            O, I, KK, r0, r1 = L.O, L.I, L.KK, L.r0, L.r1
            T, E0, E1 = w['T'].ptr, w['e0']['E'].ptr, w['e1']['E'].ptr
            up = u_list[li].data_ptr() if u_list[li] is not None else 0
            self.row['unfold'][li] = (w_list[li].data_ptr(), up, T, 0, O, I, KK, 0)
            self.row['fold'][li] = (0, 0, w['ZT'].ptr, z_list[li].data_ptr(), O, I, KK, 0)
            # P0 (O*KK x r1) = T (O*KK x I) . E1^T          [B(k=i, j) = E1[j, i]]
            self.row['p0'][li] = (T, E1, w['P0'].ptr, 0, I, 1, 1, I, r1, O * KK, r1, I, 0)
            # P1 (r0 x KK*I) = E0 (r0 x O) . T (O x KK*I)
            self.row['p1'][li] = (E0, T, w['P1'].ptr, 0, O, 1, KK * I, 1, KK * I, r0, KK * I, O, 0)
            # core (r0*KK x r1) = P1 (r0*KK x I) . E1^T
            self.row['core'][li] = (w['P1'].ptr, E1, w['core'].ptr, 0, I, 1, 1, I, r1, r0 * KK, r1, I, 0)
            # tmp (r0*KK x I) = core (r0*KK x r1) . E1 (r1 x I)
            self.row['rec1'][li] = (w['core'].ptr, E1, w['tmp'].ptr, 0, r1, 1, I, 1, I, r0 * KK, I, r1, 0)
            # ZT (O x KK*I) = E0^T (O x r0) . tmp (r0 x KK*I)       [A(i, k) = E0[k, i]]
            self.row['rec2'][li] = (E0, w['tmp'].ptr, w['ZT'].ptr, 0, 1, O, KK * I, 1, KK * I, O, KK * I, r0, 0)
            self.row['nx'][li] = (T, L.numel)
            self.row['nc'][li] = (w['core'].ptr, r0 * KK * r1)
            self.gram_ops.append(dict(
                # unfold_0(X) rows: read from W and U in place (the order of the reduction index is irrelevant)
                init0=(dict(a=w_list[li].data_ptr(), a2=up, si=KK * I, sb=0, sc=1, nb=1, nc=KK * I) if self.gram_in_place
                       else dict(a=T, si=KK * I, sb=0, sc=1, nb=1, nc=KK * I)),
                init1=dict(a=T, si=1, sb=0, sc=I, nb=1, nc=O * KK),                  # columns of (O*KK x I)
                sweep0=dict(a=w['P0'].ptr, si=KK * r1, sb=0, sc=1, nb=1, nc=KK * r1),
                sweep1=dict(a=w['P1'].ptr, si=1, sb=0, sc=I, nb=1, nc=r0 * KK)))
        self._tab_cache = {}
        self.eig.tables = {}
        self._bound = key

    def _tab(self, kind, active):
        key = (kind, active)
        if key not in self._tab_cache:
            self._tab_cache[key] = rt.TaskTable(self.row[kind][list(active)], self.device)
        return self._tab_cache[key]

    def _eig_tabs(self, which, active):
        mode = 0 if which in ('init0', 'sweep0') else 1
        probs = [(self.ws[li]['e%d' % mode], self.gram_ops[li][which]) for li in active]
        return self.eig.build((which, active), probs)

    def run(self, w_list, u_list, z_list):
        try:
            self._run(w_list, u_list, z_list)
        except EigClusterFlag:
            # exactly repeated singular values: redo the projection on the Jacobi solver, and stay there
            self._alloc('jacobi')
            self._run(w_list, u_list, z_list)

    def _run(self, w_list, u_list, z_list):
        self.bind(w_list, u_list, z_list)
        n = len(self.layers)
        everyone = tuple(range(n))
        ph = _Phases(self.profile)
        ph.mark('unfold')
        rt.unfold_add(self._tab('unfold', everyone))
        self.norms.zero_()
        rt.sqnorm(self._tab('nx', everyone), self.norms[:n])
        ph.mark('hosvd')
        # HOSVD initialisation: both factor problems of every layer in one eigensolver batch
        probs = [(self.ws[li]['e0'], self.gram_ops[li]['init0']) for li in everyone] + \
                [(self.ws[li]['e1'], self.gram_ops[li]['init1']) for li in everyone]
        sw = self.eig.run(self.eig.build(('init', everyone), probs))
        jac = {L.name: [int(sw[i]), int(sw[n + i])] for i, L in enumerate(self.layers)}
        norm_x2 = self.norms[:n].cpu().numpy().copy()
        errs = [[] for _ in range(n)]
        active = everyone
        it = 0
        ph.mark('hooi')
        # tensorly's rule (`iteration > 1 and |err[-2] - err[-1]| < tol`) cannot fire before the third sweep: the first three
        # sweeps are enqueued back to back (core norms into three device slots, eigensolver flags OR-ed on the device) and
        # read back ONCE; from then on one host synchronisation per sweep decides which layers stop.
        first = min(3, self.n_iter_max)
        while active and it < self.n_iter_max:
            batch = first if it == 0 else 1
            t0, t1 = self._eig_tabs('sweep0', active), self._eig_tabs('sweep1', active)
            slots = self.norms[n:n + batch * len(active)]
            slots.zero_()
            flags = []
            for b in range(batch):
                rt.gemm(self._tab('p0', active))
                self.eig.enqueue(t0, warm=True)          # previous solve of the same mode: HOSVD init or the last sweep
                rt.gemm(self._tab('p1', active))
                self.eig.enqueue(t1, warm=True)
                rt.gemm(self._tab('core', active))
                rt.sqnorm(self._tab('nc', active), slots[b * len(active):(b + 1) * len(active)])
                if b < batch - 1:                        # the flags of this sweep are overwritten by the next one
                    flags.append(self.eig.snapshot_flags(t0, t1))
            norm_c2 = slots.cpu().numpy().reshape(batch, len(active))      # the host synchronisation
            self.eig.check_flags(flags)
            s0, s1 = self.eig.results(t0), self.eig.results(t1)
            stop = []
            for q, li in enumerate(active):
                jac[self.layers[li].name] += [int(s0[q]), int(s1[q])]
                nx = math.sqrt(float(norm_x2[li]))
                for b in range(batch):
                    errs[li].append(math.sqrt(abs(float(norm_x2[li]) - float(norm_c2[b, q]))) / nx if nx > 0 else 0.0)
                last = it + batch - 1
                if (last > 1 and abs(errs[li][-2] - errs[li][-1]) < self.hooi_tol) or last == self.n_iter_max - 1:
                    stop.append(li)
            it += batch
            if stop:
                stop = tuple(stop)
                rt.gemm(self._tab('rec1', stop))
                rt.gemm(self._tab('rec2', stop))
                rt.fold_store(self._tab('fold', stop))
                for li in stop:
                    self.hooi_sweeps[self.layers[li].name] = it
                active = tuple(li for li in active if li not in stop)
        ph.finish()
        self.sweeps = jac
        self.errors = {L.name: errs[i] for i, L in enumerate(self.layers)}


def tucker2_decompose(w, ranks, device=None):
    """`partial_tucker(w, modes=[0, 1], rank=ranks, init='svd')` on the B200 kernels (the `dense_w`
    constructor paths of TKConv.py:79-83,192-196,294-298 and TKLinear.py:47-48,98-99).
    Returns (core, [last_factor (O x r_out), first_factor (I x r_in)]) as torch tensors on `device`."""
    if device is None:
        device = w.device if w.is_cuda or rt.backend_is_emulated() else torch.device('cuda', torch.cuda.current_device())
    wt = w.detach().to(device=device, dtype=torch.float32).contiguous()
    layer = TKLayer('tucker2', wt.shape, list(ranks))
    plan = TKProjectionPlan([layer], device)
    z = torch.empty_like(wt)
    plan.run([wt], [None], [z])
    ws = plan.ws[0]
    O, I, KK, r0, r1 = layer.O, layer.I, layer.KK, layer.r0, layer.r1
    core = ws['core'].t[:r0 * KK * r1].view(r0, KK, r1).permute(0, 2, 1).contiguous()
    e0 = ws['e0']['E'].t[:r0 * O].view(r0, O)
    e1 = ws['e1']['E'].t[:r1 * I].view(r1, I)
    core = core.view((r0, r1) + tuple(wt.shape[2:])) if wt.dim() == 4 else core.view(r0, r1)
    return core.clone(), [e0.t().contiguous(), e1.t().contiguous()]
