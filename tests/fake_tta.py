"""numpy emulator of the libtta C ABI over HOST memory (test infrastructure only).

It lets `-m "not gpu"` tests drive the real host-side planning code (`projector.py`, `admm.py`, ...)
end to end on a box without a GPU: the plans bake raw addresses of torch CPU tensors into the same
task tables they would hand to the CUDA kernels, and this module interprets those tables exactly as
include/tta.h specifies.  It is never importable from the product package.
"""
from __future__ import annotations

import ctypes

import numpy as np

import tta_runtime as rt


def _view(ptr, count, dtype=np.float32):
    ptr = int(ptr)
    if count == 0:
        return np.zeros(0, dtype=dtype)
    nbytes = int(count) * np.dtype(dtype).itemsize
    return np.frombuffer((ctypes.c_char * nbytes).from_address(ptr), dtype=dtype)


def _val(p):
    if p is None:
        return 0
    if isinstance(p, ctypes.c_void_p):
        return p.value or 0
    return int(p)


def _table(ptr, n, dtype):
    return _view(_val(ptr), n * dtype.itemsize, np.uint8).view(dtype)


class FakeTTA:
    def __init__(self, scramble=True):
        self.err = b''
        self.calls = []
        self.scramble = scramble

    def tta_last_error(self):
        return self.err

    def tta_version(self):
        return 100

    def tta_check_device(self, dev):
        return 0

    # ---- elementwise ----
    def tta_dual_update_multi(self, tdev, thost, n, sq, stream):
        self.calls.append('dual_update')
        out = _view(_val(sq), n, np.float64) if _val(sq) else None
        for i, tk in enumerate(_table(thost, n, rt.EW_TASK)):
            m = int(tk['numel'])
            w, z, u = _view(tk['w'], m), _view(tk['z'], m), _view(tk['u'], m)
            d = w - z
            u += d
            if out is not None:
                out[i] += float(np.sum(d.astype(np.float64) ** 2))
        return 0

    def tta_penalty_fwd_multi(self, tdev, thost, n, rho, loss, stream):
        self.calls.append('penalty_fwd')
        out = _view(_val(loss), 1, np.float64)
        for tk in _table(thost, n, rt.EW_TASK):
            m = int(tk['numel'])
            d = _view(tk['w'], m) - _view(tk['z'], m) + _view(tk['u'], m)
            out[0] += 0.5 * float(np.float32(rho)) * float(np.sum(d.astype(np.float64) ** 2))
        return 0

    def tta_penalty_bwd_multi(self, tdev, thost, n, rho, gscale, accumulate, stream):
        self.calls.append('penalty_bwd')
        s = float(_view(_val(gscale), 1)[0]) if _val(gscale) else 1.0
        coef = np.float32(np.float32(rho) * np.float32(s))
        for tk in _table(thost, n, rt.EW_TASK):
            m = int(tk['numel'])
            d = coef * (_view(tk['w'], m) - _view(tk['z'], m) + _view(tk['u'], m))
            g = _view(tk['g'], m)
            if accumulate:
                g += d
            else:
                g[:] = d
        return 0

    # ---- fold / unfold ----
    def tta_unfold_add_batched(self, tdev, thost, n, stream):
        self.calls.append('unfold')
        for tk in _table(thost, n, rt.FOLD_TASK):
            O, I, KK = int(tk['O']), int(tk['I']), int(tk['KK'])
            v = _view(tk['w'], O * I * KK).reshape(O, I, KK).copy()
            if tk['u']:
                v += _view(tk['u'], O * I * KK).reshape(O, I, KK)
            _view(tk['t'], O * I * KK).reshape(O, KK, I)[:] = v.transpose(0, 2, 1)
        return 0

    def tta_fold_store_batched(self, tdev, thost, n, stream):
        self.calls.append('fold')
        for tk in _table(thost, n, rt.FOLD_TASK):
            O, I, KK = int(tk['O']), int(tk['I']), int(tk['KK'])
            t = _view(tk['t'], O * I * KK).reshape(O, KK, I)
            _view(tk['z'], O * I * KK).reshape(O, I, KK)[:] = t.transpose(0, 2, 1)
        return 0

    # ---- gram ----
    def tta_gram_batched(self, tdev, thost, n, stream):
        self.calls.append('gram')
        for tk in _table(thost, n, rt.GRAM_TASK):
            k, nb, nc = int(tk['k']), int(tk['nb']), int(tk['nc'])
            si, sb, sc = int(tk['si']), int(tk['sb']), int(tk['sc'])
            span = (k - 1) * si + (nb - 1) * sb + (nc - 1) * sc + 1
            base = _view(tk['a'], span)
            a = np.lib.stride_tricks.as_strided(base, shape=(k, nb, nc), strides=(4 * si, 4 * sb, 4 * sc))
            a = a.reshape(k, nb * nc).astype(np.float64)
            if tk['a2']:
                b2 = np.lib.stride_tricks.as_strided(_view(tk['a2'], span), shape=(k, nb, nc), strides=(4 * si, 4 * sb, 4 * sc))
                a = (a.astype(np.float32) + b2.reshape(k, nb * nc)).astype(np.float64)      # fp32 sum, as on the device
            g = a @ a.T
            ld, kpad = int(tk['ld']), int(tk['kpad'])
            if tk['x']:
                x = _view(tk['x'], ld * kpad).reshape(kpad, ld)
                x[:] = 0
                x[:k, :k] = g.T.astype(np.float32)
            if tk['g64']:
                _view(tk['g64'], k * k, np.float64).reshape(k, k)[:] = g
        return 0

    # ---- eigensolver: exact fp64 eigh, stored as x_j = lambda_j v_j in scrambled column order ----
    def tta_jacobi_scratch_bytes(self, thost, n):
        return 4096

    def tta_jacobi_eigh_batched(self, tdev, thost, n, tol, max_sweeps, scratch, scratch_bytes, sweeps_out, stream):
        self.calls.append('jacobi')
        sw = _view(_val(sweeps_out), n, np.int32) if _val(sweeps_out) else None
        rng = np.random.RandomState(1234)
        for i, tk in enumerate(_table(thost, n, rt.EIG_TASK)):
            k, ld, kpad = int(tk['k']), int(tk['ld']), int(tk['kpad'])
            x = _view(tk['x'], ld * kpad).reshape(kpad, ld)
            # one-sided Jacobi on the columns of M (= G, or G Q_prev with a warm start): M J = U S with
            # orthogonal columns, i.e. the left singular vectors scaled by the singular values
            m = x[:k, :k].T.astype(np.float64)
            v, lam, _ = np.linalg.svd(m)
            order = rng.permutation(k) if self.scramble else np.arange(k)
            x[:k, :k] = (v[:, order] * lam[order]).T.astype(np.float32)
            if sw is not None:
                sw[i] = 7
        # asynchronous mode (sweeps_out == NULL): results are left in the scratch buffer
        sc = _view(_val(scratch), 6 * n, np.int32)
        sc[2 * n:3 * n] = 7
        sc[5 * n:6 * n] = 1
        return 0

    def tta_jacobi_read_results(self, scratch_host, thost, n, max_sweeps, sweeps_out):
        sc = _view(_val(scratch_host), 6 * n, np.int32)
        if _val(sweeps_out):
            _view(_val(sweeps_out), n, np.int32)[:] = sc[2 * n:3 * n]
        return 0 if np.all(sc[5 * n:6 * n] != 0) else -4

    # ---- fp64 dominant-r eigensolver (tridiagonalisation route on the GPU): exact eigh here ----
    def tta_symeig_work_doubles(self, k, r):
        kp = (k + 31) // 32 * 32
        return 8 + 4 * kp + k * kp + 8 * ((k + 1) // 4 + 1)

    def tta_symeig_max_k(self):
        return 608

    def tta_symeig_top_batched(self, tdev, thost, n, stream):
        self.calls.append('symeig')
        for tk in _table(thost, n, rt.SYMEIG_TASK):
            k, r = int(tk['k']), int(tk['r'])
            g = _view(tk['g'], k * k, np.float64).reshape(k, k)
            lam, v = np.linalg.eigh(g)
            _view(tk['lam'], r, np.float64)[:] = lam[::-1][:r]
            _view(tk['e64'], r * k, np.float64).reshape(r, k)[:] = v[:, ::-1][:, :r].T * 1.7    # not normalised
            _view(tk['status'], 1, np.int32)[0] = 0
        return 0

    def tta_select_batched(self, tdev, thost, n, stream):
        self.calls.append('select')
        for tk in _table(thost, n, rt.SELECT_TASK):
            k, ld, r = int(tk['k']), int(tk['ld']), int(tk['r'])
            kpad_rows = k  # only the first k columns are read
            x = _view(tk['x'], ld * kpad_rows).reshape(kpad_rows, ld)[:, :k]
            lam = np.sqrt(np.sum(x.astype(np.float32) ** 2, axis=1, dtype=np.float32))
            order = np.argsort(-lam, kind='stable')[:r]
            lmax = lam.max() if k else 0.0
            e = np.zeros((r, k), dtype=np.float32)
            sg = np.zeros(r, dtype=np.float32)
            for p, j in enumerate(order):
                if lam[j] > lmax * 4e-7 and lam[j] > 0:
                    e[p] = x[j] / lam[j]
                    sg[p] = np.sqrt(lam[j])
            _view(tk['e'], r * k).reshape(r, k)[:] = e
            if tk['et']:
                _view(tk['et'], r * k).reshape(k, r)[:] = e.T
            if tk['se']:
                _view(tk['se'], r * k).reshape(r, k)[:] = e * sg[:, None]
            if tk['sigma']:
                _view(tk['sigma'], r)[:] = sg
            if tk['isigma']:
                _view(tk['isigma'], r)[:] = np.where(sg > 0, 1.0 / np.where(sg > 0, sg, 1), 0).astype(np.float32)
        return 0

    # ---- refinement (include/tta.h: one Ogita-Aishima step) ----
    def tta_refine_prepare_batched(self, tdev, thost, n, stream):
        self.calls.append('refine_prepare')
        for tk in _table(thost, n, rt.REFINE_TASK):
            k, ld = int(tk['k']), int(tk['ld'])
            x = _view(tk['x'], ld * k).reshape(k, ld)[:, :k].astype(np.float64)
            nrm2 = np.sum(x * x, axis=1)
            cut = nrm2.max() * (4e-7) ** 2 if k else 0.0
            live = (nrm2 > cut) & (nrm2 > 0)
            inv = np.where(live, 1.0 / np.sqrt(np.where(nrm2 > 0, nrm2, 1.0)), 0.0)
            order = np.argsort(-nrm2, kind='stable')
            _view(tk['qt'], k * k, np.float64).reshape(k, k)[:] = (x * inv[:, None])[order]
            _view(tk['lam0'], k, np.float64)[:] = np.where(live, np.sqrt(nrm2), 0.0)[order]
        return 0

    def tta_refine_coeff_batched(self, tdev, thost, n, stream):
        self.calls.append('refine_coeff')
        for tk in _table(thost, n, rt.REFINE_TASK):
            k, r, wnd = int(tk['k']), int(tk['r']), int(tk['wnd'])
            S = _view(tk['s'], wnd * k, np.float64).reshape(wnd, k)
            T = _view(tk['t'], wnd * k, np.float64).reshape(wnd, k)
            lam0 = _view(tk['lam0'], k, np.float64)
            tdiag = np.ones(k)
            tdiag[:wnd] = T[np.arange(wnd), np.arange(wnd)]
            sdiag = S[np.arange(wnd), np.arange(wnd)]
            lam = lam0.copy()
            lam[:wnd] = np.where(tdiag[:wnd] > 0.5, sdiag / np.where(tdiag[:wnd] > 0.5, tdiag[:wnd], 1.0), 0.0)
            live = np.where(np.arange(k) < wnd, tdiag > 0.5, lam > 0.0)
            order = np.argsort(-lam[:wnd], kind='stable')[:r]
            lmax = lam.max()
            C = np.zeros((r, k))
            for p, j in enumerate(order):
                rrow = -T[j]
                gap = lam[j] - lam
                c = 0.5 * rrow
                with np.errstate(divide='ignore', invalid='ignore'):
                    e = (S[j] + lam[j] * rrow) / gap
                use = (np.abs(gap) > 1e-6 * lmax) & (np.abs(e) <= 0.05)
                c = np.where(use, e, c)
                c = np.where(live, c, 0.0)
                c[j] = 1.0 + 0.5 * (1.0 - T[j, j])
                C[p] = c
            _view(tk['c'], r * k, np.float64).reshape(r, k)[:] = C
            _view(tk['lam'], r, np.float64)[:] = lam[order]
        return 0

    def tta_refine_finalize_batched(self, tdev, thost, n, stream):
        self.calls.append('refine_finalize')
        for tk in _table(thost, n, rt.REFINE_TASK):
            k, r = int(tk['k']), int(tk['r'])
            lam = _view(tk['lam'], r, np.float64)
            live = (lam > 4e-7 * lam[0]) & (lam > 0)
            e64 = _view(tk['e64'], r * k, np.float64).reshape(r, k)
            nrm = np.linalg.norm(e64, axis=1)
            e64 = e64 / np.where(nrm > 0, nrm, 1.0)[:, None]
            e = np.where(live[:, None], e64, 0.0).astype(np.float32)
            sg = np.where(live, np.sqrt(np.where(live, lam, 0.0)), 0.0).astype(np.float32)
            _view(tk['e'], r * k).reshape(r, k)[:] = e
            if tk['et']:
                _view(tk['et'], r * k).reshape(k, r)[:] = e.T
            if tk['se']:
                _view(tk['se'], r * k).reshape(r, k)[:] = e * sg[:, None]
            if tk['sigma']:
                _view(tk['sigma'], r)[:] = sg
            if tk['isigma']:
                _view(tk['isigma'], r)[:] = np.where(live, 1.0 / np.sqrt(np.where(live, lam, 1.0)), 0.0).astype(np.float32)
        return 0

    # ---- gemm ----
    def tta_gemm_f64_batched(self, tdev, thost, n, stream):
        self.calls.append('gemm_f64')
        return self._gemm(thost, n, np.float64)

    def tta_gemm_batched(self, tdev, thost, n, stream):
        self.calls.append('gemm')
        return self._gemm(thost, n, np.float32)

    def _gemm(self, thost, n, dt):
        isz = np.dtype(dt).itemsize
        for tk in _table(thost, n, rt.GEMM_TASK):
            M, N, K = int(tk['M']), int(tk['N']), int(tk['K'])
            guarded = dt == np.float64 and int(tk['flags']) & rt.GEMM_GUARD
            if guarded and float(_view(tk['colscale'], 1, np.float64)[0]) == 0.0:
                continue
            sai, sak, sbk, sbj, ldc = (int(tk[f]) for f in ('sai', 'sak', 'sbk', 'sbj', 'ldc'))
            abase = _view(tk['a'], (M - 1) * sai + (K - 1) * sak + 1, dt)
            bbase = _view(tk['b'], (K - 1) * sbk + (N - 1) * sbj + 1, dt)
            a = np.lib.stride_tricks.as_strided(abase, shape=(M, K), strides=(isz * sai, isz * sak))
            b = np.lib.stride_tricks.as_strided(bbase, shape=(K, N), strides=(isz * sbk, isz * sbj))
            c = (a @ b).astype(dt)
            if tk['colscale'] and not guarded:
                c = c * _view(tk['colscale'], N, dt)[None, :]
            odt = np.float32 if (dt == np.float64 and int(tk['flags']) & rt.GEMM_STORE_F32) else dt
            osz = np.dtype(odt).itemsize
            cbase = _view(tk['c'], (M - 1) * ldc + N, odt)
            np.lib.stride_tricks.as_strided(cbase, shape=(M, N), strides=(osz * ldc, osz))[:] = c.astype(odt)
        return 0

    # ---- orthogonality regulariser ----
    def _orth_x(self, tk):
        n, ln, si, st = int(tk['n']), int(tk['len']), int(tk['si']), int(tk['st'])
        base = _view(tk['p'], (n - 1) * si + (ln - 1) * st + 1)
        return np.lib.stride_tricks.as_strided(base, shape=(n, ln), strides=(4 * si, 4 * st))

    def tta_orth_penalty_fwd_batched(self, tdev, thost, n, rho, loss, stream):
        self.calls.append('orth_fwd')
        out = _view(_val(loss), 1, np.float64)
        for tk in _table(thost, n, rt.ORTH_TASK):
            x = self._orth_x(tk).astype(np.float32)
            r = x @ x.T - np.eye(x.shape[0], dtype=np.float32)
            _view(tk['r'], r.size)[:] = r.reshape(-1)
            out[0] += 0.5 * float(np.float32(rho)) * float(np.sum(r.astype(np.float64) ** 2))
        return 0

    def tta_orth_penalty_bwd_batched(self, tdev, thost, n, rho, gscale, accumulate, stream):
        self.calls.append('orth_bwd')
        s = float(_view(_val(gscale), 1)[0]) if _val(gscale) else 1.0
        for tk in _table(thost, n, rt.ORTH_TASK):
            x = self._orth_x(tk).astype(np.float32)
            nn, ln, si, st = int(tk['n']), int(tk['len']), int(tk['si']), int(tk['st'])
            r = _view(tk['r'], nn * nn).reshape(nn, nn)
            d = (np.float32(2.0 * rho * s) * (r @ x)).astype(np.float32)
            gbase = _view(tk['g'], (nn - 1) * si + (ln - 1) * st + 1)
            g = np.lib.stride_tricks.as_strided(gbase, shape=(nn, ln), strides=(4 * si, 4 * st))
            if accumulate:
                g += d
            else:
                g[:] = d
        return 0

    def tta_sqnorm_batched(self, tdev, thost, n, out, stream):
        self.calls.append('sqnorm')
        o = _view(_val(out), n, np.float64)
        for i, tk in enumerate(_table(thost, n, rt.SQNORM_TASK)):
            o[i] += float(np.sum(_view(tk['x'], int(tk['n'])).astype(np.float64) ** 2))
        return 0
