"""numpy restatement of the reference's projection / dual-update arithmetic (CPU oracle, test-only).

Every function cites the reference file:line it follows.  fp32 in, fp32 out, LAPACK `gesdd` through
`numpy.linalg.svd` exactly like the reference (ttd.py:17, admm.py:131,143).
"""
from __future__ import annotations

import math

import numpy as np


# ----------------------------------------------------------------------------------------------
# TT-SVD and reconstruction                                             (ttd.py:10-31, ttd.py:34-43)
# ----------------------------------------------------------------------------------------------
def tt_svd(x, shapes, ranks):
    """Sequential truncated-SVD sweep (ttd.py:10-31).

    `ranks` is clipped IN PLACE when an unfolding has fewer singular values than requested
    (ttd.py:18-19) -- callers rely on that side effect.
    Returns the list of cores, core i shaped (ranks[i], shapes[i], ranks[i+1]).
    """
    order = len(shapes)
    carry = x
    cores = []
    for i in range(order - 1):
        mat = np.reshape(carry, (ranks[i] * shapes[i], -1))
        left, sing, right = np.linalg.svd(mat, full_matrices=False)
        avail = sing.shape[0]
        if avail < ranks[i + 1]:
            ranks[i + 1] = avail
        keep = ranks[i + 1]
        cores.append(left[:, :keep].reshape(ranks[i], shapes[i], keep))
        # ttd.py:26 forms diag(s) @ v; row scaling is the same product
        carry = np.dot(np.diag(sing[:keep]), right[:keep, :])
    cores.append(np.reshape(carry, (ranks[order - 1], shapes[order - 1], ranks[order])))
    return cores


def tt_contract(cores, out_shape):
    """Left-to-right core chain (ttd.py:34-43)."""
    acc = cores[0]
    for core in cores[1:]:
        r = core.shape[0]
        acc = np.dot(np.reshape(acc, (-1, r)), np.reshape(core, (r, -1)))
    return np.reshape(acc, out_shape)


def project_conv_tt(w, shapes, ranks):
    """admm.py:91-101: (O,I,kh,kw) -> (O,kh*kw,I) -> TT-SVD -> back.  `ranks` passed by reference."""
    o, i, kh, kw = w.shape
    kk = kh * kw
    t = np.transpose(np.reshape(w, (o, i, kk)), (0, 2, 1))
    cores = tt_svd(t, shapes, ranks)
    z = tt_contract(cores, (o, kk, i))
    return np.reshape(np.transpose(z, (0, 2, 1)), w.shape)


def project_linear_tt(w, shapes, ranks):
    """admm.py:103-111: the rank tuple is copied (`list(...)`, :105), so no caller-visible clip."""
    ranks = list(ranks)
    cores = tt_svd(np.reshape(w, shapes), shapes, ranks)
    return tt_contract(cores, w.shape)


# ----------------------------------------------------------------------------------------------
# rank-r matrix projection                                                      (admm.py:129-149)
# ----------------------------------------------------------------------------------------------
def _rank_of(entry):
    return entry if isinstance(entry, int) else entry[0]


def project_matrix_svd(mat, rank):
    left, sing, right = np.linalg.svd(mat, full_matrices=False)
    return left[:, :rank] @ np.diag(sing[:rank]) @ right[:rank, :]


def project_conv_svd(w, rank_entry):
    """admm.py:129-139 (1x1 conv: squeeze, project, re-expand)."""
    z = project_matrix_svd(np.squeeze(w), _rank_of(rank_entry))
    return z[:, :, None, None]


def project_linear_svd(w, rank_entry):
    """admm.py:141-149."""
    return project_matrix_svd(w, _rank_of(rank_entry))


# ----------------------------------------------------------------------------------------------
# Tucker-2 HOOI: restatement of tensorly (<0.8) `partial_tucker(..., modes=[0,1], init='svd')`
# Call sites: admm.py:116,124; TKConv.py:79-80,192-193,294-295.   PARITY UNPINNED (see __init__).
# ----------------------------------------------------------------------------------------------
def _unfold(t, mode):
    return np.reshape(np.moveaxis(t, mode, 0), (t.shape[mode], -1))


def _mode_dot_t(t, factor, mode):
    """t x_mode factor^T  (factor: dim x r  ->  mode size becomes r)."""
    moved = np.moveaxis(t, mode, 0)
    shp = moved.shape
    res = factor.T @ moved.reshape(shp[0], -1)
    return np.moveaxis(res.reshape((factor.shape[1],) + shp[1:]), 0, mode)


def _mode_dot(t, factor, mode):
    moved = np.moveaxis(t, mode, 0)
    shp = moved.shape
    res = factor @ moved.reshape(shp[0], -1)
    return np.moveaxis(res.reshape((factor.shape[0],) + shp[1:]), 0, mode)


def _top_left_vectors(mat, r):
    left, _, _ = np.linalg.svd(mat, full_matrices=False)
    return left[:, :r]


def partial_tucker2(x, ranks, n_iter_max=100, tol=10e-5, return_info=False):
    """HOSVD init + HOOI sweeps on modes (0, 1); stop when iteration > 1 and |d err| < tol.

    Returns core, [A0, A1] (tensorly < 0.8 convention).  With `return_info`, also the sweep count and
    the error history.
    """
    modes = (0, 1)
    factors = [_top_left_vectors(_unfold(x, m), ranks[k]) for k, m in enumerate(modes)]
    norm_x = np.sqrt(np.sum(x * x))
    errs = []
    core = None
    sweeps = 0
    for it in range(n_iter_max):
        for k, m in enumerate(modes):
            other = 1 - k
            partial = _mode_dot_t(x, factors[other], modes[other])
            factors[k] = _top_left_vectors(_unfold(partial, m), ranks[k])
        core = _mode_dot_t(_mode_dot_t(x, factors[0], 0), factors[1], 1)
        norm_c = np.sqrt(np.sum(core * core))
        errs.append(math.sqrt(abs(float(norm_x) ** 2 - float(norm_c) ** 2)) / float(norm_x))
        sweeps = it + 1
        if it > 1 and abs(errs[-2] - errs[-1]) < tol:
            break
    if return_info:
        return core, factors, sweeps, errs
    return core, factors


def tucker2_to_tensor(core, factors):
    """tl.tucker_to_tensor for a mode-(0,1) partial Tucker (admm.py:117,125)."""
    return _mode_dot(_mode_dot(core, factors[0], 0), factors[1], 1)


def project_tk(w, ranks, return_sweeps=False):
    """admm.py:113-127 (conv and linear share the code)."""
    core, factors, sweeps, _ = partial_tucker2(w, list(ranks), return_info=True)
    z = tucker2_to_tensor(core, factors).astype(w.dtype, copy=False)
    return (z, sweeps) if return_sweeps else z


# ----------------------------------------------------------------------------------------------
# ADMM state machine                                                             (admm.py:15-89)
# ----------------------------------------------------------------------------------------------
def project_layer(v, fmt, rank_entry, shapes=None, info=None):
    """Dispatch of admm.py:47-69 on tensor rank / rank-list length."""
    multi = (not isinstance(rank_entry, int)) and len(rank_entry) > 1
    if v.ndim == 4:
        if fmt == 'tk' and multi:
            z, sw = project_tk(v, rank_entry, return_sweeps=True)
            if info is not None:
                info['sweeps'] = sw
            return z
        if fmt == 'tt' and multi:
            return project_conv_tt(v, shapes, rank_entry)
        return project_conv_svd(v, rank_entry)
    if v.ndim == 2:
        if fmt == 'tk':
            z, sw = project_tk(v, rank_entry, return_sweeps=True)
            if info is not None:
                info['sweeps'] = sw
            return z
        if fmt == 'tt':
            return project_linear_tt(v, shapes, rank_entry)
        return project_linear_svd(v, rank_entry)
    raise Exception('ERROR: unsupported layer in ADMM!')


class OracleADMM:
    """numpy mirror of admm.ADMM over a dict name -> fp32 ndarray (admm.py:15-89)."""

    def __init__(self, weights, rho, hp, fmt):
        if fmt == 'none':
            raise Exception('ERROR: Tensor format should be specified!')
        self.w = weights
        self.rho = self.init_rho = rho
        self.hp = hp
        self.fmt = fmt
        self.names = [n for n in weights if n in hp.ranks]
        self.u = {n: np.zeros_like(weights[n]) for n in self.names}
        self.z = {n: weights[n].copy() for n in self.names}
        self.sweeps = {}
        self.diff_norm = {}

    def update(self, update_u=True):
        for n in self.names:
            v = self.w[n] + self.u[n]
            info = {}
            shapes = getattr(self.hp, 'tt_shapes', {}).get(n) if self.fmt == 'tt' else None
            self.z[n] = np.ascontiguousarray(project_layer(v, self.fmt, self.hp.ranks[n], shapes, info),
                                             dtype=np.float32)
            if 'sweeps' in info:
                self.sweeps[n] = info['sweeps']
            if update_u:
                diff = self.w[n] - self.z[n]
                self.u[n] = self.u[n] + diff
                self.diff_norm[n] = float(np.sqrt(np.sum(diff.astype(np.float64) ** 2)))

    def penalty(self):
        """admm.py:80-85: sum_l 0.5*rho*||W - Z + U||^2 (fp32 accumulation like torch.norm)."""
        tot = 0.0
        for n in self.names:
            d = self.w[n] - self.z[n] + self.u[n]
            tot += 0.5 * self.rho * float(np.sum(d.astype(np.float64) ** 2))
        return tot

    def penalty_grad(self, n):
        return self.rho * (self.w[n] - self.z[n] + self.u[n])


# ----------------------------------------------------------------------------------------------
# decomposed-layer forwards (numpy/torch-free statement of the contraction chains)
# ----------------------------------------------------------------------------------------------
def split_tt_conv(shapes, ranks, out_channels):
    """TTConv.py:49-68: first prefix whose product equals out_channels is the 'out' part."""
    prod = 1
    for i, s in enumerate(shapes):
        prod *= s
        if prod == out_channels:
            out_order = i + 1
            break
    else:
        raise ValueError('tt_shapes do not factor out_channels')
    return out_order, len(shapes) - out_order - 1


def tt_linear_weight(cores, out_features, in_features):
    """TTLinear.py:151-157 (_recover_weight): chain of cores reshaped to (out, in)."""
    return tt_contract(cores, (out_features, in_features))
