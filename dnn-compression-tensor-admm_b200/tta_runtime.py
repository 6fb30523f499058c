"""ctypes binding of libtta.so (include/tta.h) -- the only bridge between the Python host code and
the sm_100a kernels.  No torch types cross the boundary: raw device pointers, sizes and the current
CUDA stream handle.

There is NO CPU fallback: if libtta.so is missing or the device is not sm_100, every entry point
raises.  (tests/ may inject an emulator of the C ABI through `set_backend_for_tests` to exercise the
host-side index logic on a box without a GPU; product code never does.)
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

# The layer groups of ADMM.update() and the eigensolver's internal streams need more hardware work queues
# than the default 8 (streams that share a queue serialise); read by the driver when the context is created.
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('TTA_LIB') or os.path.join(_HERE, 'libtta.so')      # TTA_LIB: another build of the same library (A/B measurements)

# ---- struct layouts (must mirror include/tta.h; checked by tests/test_abi.py against the header) ----
P = np.uint64  # pointers
EW_TASK = np.dtype([('w', P), ('z', P), ('u', P), ('g', P), ('numel', np.int64)], align=True)
FOLD_TASK = np.dtype([('w', P), ('u', P), ('t', P), ('z', P), ('O', np.int32), ('I', np.int32),
                      ('KK', np.int32), ('pad_', np.int32)], align=True)
GRAM_TASK = np.dtype([('a', P), ('part', P), ('x', P), ('g64', P), ('si', np.int64), ('sb', np.int64), ('sc', np.int64),
                      ('k', np.int32), ('nb', np.int32), ('nc', np.int32), ('nsplit', np.int32),
                      ('ld', np.int32), ('kpad', np.int32), ('a2', P)], align=True)
EIG_TASK = np.dtype([('x', P), ('k', np.int32), ('ld', np.int32), ('kpad', np.int32), ('bw', np.int32)],
                    align=True)
SELECT_TASK = np.dtype([('x', P), ('e', P), ('et', P), ('se', P), ('sigma', P), ('isigma', P),
                        ('k', np.int32), ('ld', np.int32), ('r', np.int32), ('pad_', np.int32)], align=True)
GEMM_TASK = np.dtype([('a', P), ('b', P), ('c', P), ('colscale', P), ('sai', np.int64), ('sak', np.int64),
                      ('sbk', np.int64), ('sbj', np.int64), ('ldc', np.int64), ('M', np.int32),
                      ('N', np.int32), ('K', np.int32), ('flags', np.int32)], align=True)
GEMM_STORE_F32 = 1     # tta_gemm_f64_batched: c is a float array
GEMM_GUARD = 2         # tta_gemm_f64_batched: colscale points to one double; the task is skipped when it is 0
SQNORM_TASK = np.dtype([('x', P), ('n', np.int64)], align=True)
REFINE_TASK = np.dtype([('x', P), ('qt', P), ('s', P), ('t', P), ('c', P), ('lam', P), ('lam0', P), ('e64', P),
                        ('e', P), ('et', P), ('se', P), ('sigma', P), ('isigma', P), ('k', np.int32),
                        ('ld', np.int32), ('r', np.int32), ('wnd', np.int32)], align=True)

SYMEIG_TASK = np.dtype([('g', P), ('work', P), ('lam', P), ('e64', P), ('status', P), ('k', np.int32), ('r', np.int32)],
                       align=True)

ORTH_TASK = np.dtype([('p', P), ('r', P), ('g', P), ('si', np.int64), ('st', np.int64), ('n', np.int32),
                      ('len', np.int32)], align=True)

STRUCT_SIZES = {'tta_ew_task': EW_TASK.itemsize, 'tta_fold_task': FOLD_TASK.itemsize,
                'tta_gram_task': GRAM_TASK.itemsize, 'tta_eig_task': EIG_TASK.itemsize,
                'tta_select_task': SELECT_TASK.itemsize, 'tta_gemm_task': GEMM_TASK.itemsize,
                'tta_sqnorm_task': SQNORM_TASK.itemsize, 'tta_refine_task': REFINE_TASK.itemsize,
                'tta_symeig_task': SYMEIG_TASK.itemsize, 'tta_orth_task': ORTH_TASK.itemsize}

EXPORTS = ['tta_last_error', 'tta_version', 'tta_launch_count', 'tta_check_device', 'tta_jacobi_profile_enable',
           'tta_jacobi_profile_read', 'tta_jacobi_force_multilaunch', 'tta_jacobi_enable_gra',
           'tta_jacobi_set_stop_rel', 'tta_gemm_enable_tc', 'tta_gram_enable_tc', 'tta_dual_update_multi',
           'tta_penalty_fwd_multi', 'tta_penalty_bwd_multi', 'tta_unfold_add_batched',
           'tta_fold_store_batched', 'tta_gram_batched', 'tta_jacobi_eigh_batched',
           'tta_jacobi_scratch_bytes', 'tta_jacobi_read_results', 'tta_select_batched', 'tta_gemm_batched', 'tta_sqnorm_batched',
           'tta_gemm_f64_batched', 'tta_refine_prepare_batched', 'tta_refine_coeff_batched',
           'tta_refine_finalize_batched', 'tta_gemm_bf16_tc', 'tta_small_gemm', 'tta_cast_bf16',
           'tta_nchw_to_nhwc_bf16', 'tta_nhwc_to_nchw_f32', 'tta_im2col_bf16', 'tta_ttconv_fused_fwd', 'tta_ttconv_tc_fwd', 'tta_ttconv_tc_supported', 'tta_ttconv_tc_pack', 'tta_ttconv_tc_blob_bytes', 'tta_gemm_bf16_tn', 'tta_gemm_bf16_tn_workspace_bytes',
           'tta_lowrank2_fwd', 'tta_symeig_top_batched', 'tta_symeig_work_doubles', 'tta_symeig_max_k',
           'tta_symeig_profile_enable', 'tta_symeig_profile_read', 'tta_symeig_stage_profile_read', 'tta_orth_penalty_fwd_batched',
           'tta_orth_penalty_bwd_batched']


class TtaError(RuntimeError):
    pass


_LIB = None
_FAKE = None
_CHECKED_DEVICES = set()


def set_backend_for_tests(fake):
    """tests/ only: route the C-ABI calls to an emulator operating on host memory."""
    global _FAKE
    _FAKE = fake


def backend_is_emulated():
    return _FAKE is not None


def _load():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.isfile(LIB_PATH):
        raise TtaError('libtta.so not found at {} -- build it with `python -c "import __graft_entry__ as g; '
                       'g.build()"` (nvcc, sm_100a). There is no CPU fallback.'.format(LIB_PATH))
    lib = ctypes.CDLL(LIB_PATH)
    vp, ci, cf, cs = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t
    lib.tta_last_error.restype = ctypes.c_char_p
    lib.tta_last_error.argtypes = []
    lib.tta_version.restype = ci
    lib.tta_launch_count.restype = ctypes.c_ulonglong
    lib.tta_launch_count.argtypes = []
    lib.tta_jacobi_profile_enable.argtypes = [ci]
    lib.tta_jacobi_profile_enable.restype = None
    lib.tta_jacobi_force_multilaunch.argtypes = [ci]
    lib.tta_jacobi_force_multilaunch.restype = None
    lib.tta_jacobi_enable_gra.argtypes = [ci]
    lib.tta_jacobi_enable_gra.restype = None
    lib.tta_jacobi_set_stop_rel.argtypes = [cf]
    lib.tta_jacobi_set_stop_rel.restype = None
    lib.tta_gemm_enable_tc.argtypes = [ci]
    lib.tta_gemm_enable_tc.restype = None
    lib.tta_gram_enable_tc.argtypes = [ci]
    lib.tta_gram_enable_tc.restype = None
    lib.tta_jacobi_profile_read.argtypes = [vp, vp]
    lib.tta_jacobi_profile_read.restype = None
    lib.tta_check_device.argtypes = [ci]
    lib.tta_dual_update_multi.argtypes = [vp, vp, ci, vp, vp]
    lib.tta_penalty_fwd_multi.argtypes = [vp, vp, ci, cf, vp, vp]
    lib.tta_penalty_bwd_multi.argtypes = [vp, vp, ci, cf, vp, ci, vp]
    lib.tta_unfold_add_batched.argtypes = [vp, vp, ci, vp]
    lib.tta_fold_store_batched.argtypes = [vp, vp, ci, vp]
    lib.tta_gram_batched.argtypes = [vp, vp, ci, vp]
    lib.tta_jacobi_eigh_batched.argtypes = [vp, vp, ci, cf, ci, vp, cs, vp, vp]
    lib.tta_jacobi_scratch_bytes.argtypes = [vp, ci]
    lib.tta_jacobi_scratch_bytes.restype = cs
    lib.tta_jacobi_read_results.argtypes = [vp, vp, ci, ci, vp]
    lib.tta_select_batched.argtypes = [vp, vp, ci, vp]
    lib.tta_gemm_batched.argtypes = [vp, vp, ci, vp]
    lib.tta_sqnorm_batched.argtypes = [vp, vp, ci, vp, vp]
    lib.tta_orth_penalty_fwd_batched.argtypes = [vp, vp, ci, cf, vp, vp]
    lib.tta_orth_penalty_bwd_batched.argtypes = [vp, vp, ci, cf, vp, ci, vp]
    lib.tta_symeig_work_doubles.argtypes = [ci, ci]
    lib.tta_symeig_work_doubles.restype = cs
    lib.tta_symeig_max_k.argtypes = []
    lib.tta_symeig_profile_enable.argtypes = [ci]
    lib.tta_symeig_profile_enable.restype = None
    lib.tta_symeig_profile_read.argtypes = [vp, vp]
    lib.tta_symeig_profile_read.restype = None
    lib.tta_symeig_stage_profile_read.argtypes = [vp, vp, ci]
    for nm in ('tta_gemm_f64_batched', 'tta_refine_prepare_batched', 'tta_refine_coeff_batched',
               'tta_refine_finalize_batched', 'tta_symeig_top_batched'):
        getattr(lib, nm).argtypes = [vp, vp, ci, vp]
    i64 = ctypes.c_int64
    lib.tta_gemm_bf16_tc.argtypes = [vp, i64, vp, i64, vp, i64, ci, ci, ci, vp, ci, vp]
    lib.tta_small_gemm.argtypes = [vp, ci, vp, vp, ci, vp, i64, ci, ci, i64, i64, i64, i64, i64, i64, i64, i64, vp]
    lib.tta_cast_bf16.argtypes = [vp, vp, i64, vp]
    lib.tta_nchw_to_nhwc_bf16.argtypes = [vp, vp, ci, ci, ci, ci, vp]
    lib.tta_nhwc_to_nchw_f32.argtypes = [vp, ci, vp, vp, ci, ci, ci, ci, vp]
    lib.tta_im2col_bf16.argtypes = [vp, vp] + [ci] * 16 + [vp]
    lib.tta_ttconv_fused_fwd.argtypes = [vp] * 6 + [ci] * 10 + [vp]
    lib.tta_ttconv_tc_fwd.argtypes = [vp] * 3 + [ci] * 10 + [vp]
    lib.tta_ttconv_tc_supported.argtypes = [ci] * 7
    lib.tta_ttconv_tc_pack.argtypes = [vp] * 5 + [ci] * 4 + [vp]
    lib.tta_ttconv_tc_blob_bytes.argtypes = [ci] * 4
    lib.tta_ttconv_tc_blob_bytes.restype = ctypes.c_int64
    lib.tta_gemm_bf16_tn_workspace_bytes.argtypes = [ci] * 3
    lib.tta_gemm_bf16_tn_workspace_bytes.restype = ctypes.c_int64
    lib.tta_gemm_bf16_tn.argtypes = [vp, i64, vp, i64, vp, i64, ci, ci, ci, vp, i64, vp]
    lib.tta_lowrank2_fwd.argtypes = [vp, i64, vp, i64, vp, i64, vp, vp, i64, ci, i64, ci, ci, ci, vp]
    for name in EXPORTS:
        if name not in ('tta_last_error', 'tta_jacobi_scratch_bytes', 'tta_launch_count', 'tta_symeig_work_doubles',
                        'tta_symeig_profile_enable', 'tta_symeig_profile_read',
                        'tta_jacobi_profile_enable', 'tta_jacobi_profile_read', 'tta_jacobi_force_multilaunch',
                        'tta_jacobi_enable_gra', 'tta_jacobi_set_stop_rel', 'tta_gemm_enable_tc', 'tta_gram_enable_tc',
                        'tta_ttconv_tc_blob_bytes', 'tta_gemm_bf16_tn_workspace_bytes'):
            getattr(lib, name).restype = ci
    # measurement switches: the production defaults are the tensor-core paths
    if os.environ.get('TTA_GRAM_TC') is not None:
        lib.tta_gram_enable_tc(int(os.environ['TTA_GRAM_TC']))
    if os.environ.get('TTA_GEMM_TC') is not None:
        lib.tta_gemm_enable_tc(int(os.environ['TTA_GEMM_TC']))
    _LIB = lib
    return lib


def lib():
    return _FAKE if _FAKE is not None else _load()


def _check(rc, what):
    if rc != 0:
        msg = lib().tta_last_error()
        if isinstance(msg, bytes):
            msg = msg.decode()
        raise TtaError('{} failed (rc={}): {}'.format(what, rc, msg))


def require_device(t):
    """Fail loudly unless `t` lives on an sm_100 GPU (or the tests' emulator is active)."""
    if _FAKE is not None:
        return
    if not t.is_cuda:
        raise TtaError('tensor on {}: this implementation runs on B200 (sm_100a) only; there is no CPU path'
                       .format(t.device))
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if idx not in _CHECKED_DEVICES:
        _check(lib().tta_check_device(idx), 'tta_check_device')
        _CHECKED_DEVICES.add(idx)


def stream_handle():
    if _FAKE is not None:
        return None
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class TaskTable:
    """A task table: numpy structured array (host copy) + its device mirror."""

    def __init__(self, host, device):
        self.host = np.ascontiguousarray(host)
        self.n = int(self.host.shape[0])
        if self.n == 0:
            self.dev = None
        elif _FAKE is not None:
            self.dev = None  # emulator reads the host copy
        else:
            self.dev = torch.from_numpy(self.host.view(np.uint8).reshape(-1).copy()).to(device)

    @property
    def host_ptr(self):
        return ctypes.c_void_p(self.host.ctypes.data) if self.n else None

    @property
    def dev_ptr(self):
        if self.n == 0:
            return None
        if _FAKE is not None:
            return ctypes.c_void_p(self.host.ctypes.data)
        return ctypes.c_void_p(self.dev.data_ptr())


# ---- thin call wrappers -------------------------------------------------------------------------
def dual_update(tab, sqnorm_out=None):
    p = ctypes.c_void_p(sqnorm_out.data_ptr()) if sqnorm_out is not None else None
    _check(lib().tta_dual_update_multi(tab.dev_ptr, tab.host_ptr, tab.n, p, stream_handle()), 'tta_dual_update_multi')


def penalty_fwd(tab, rho, loss_out):
    _check(lib().tta_penalty_fwd_multi(tab.dev_ptr, tab.host_ptr, tab.n, float(rho),
                                       ctypes.c_void_p(loss_out.data_ptr()), stream_handle()),
           'tta_penalty_fwd_multi')


def penalty_bwd(tab, rho, grad_scale, accumulate):
    p = ctypes.c_void_p(grad_scale.data_ptr()) if grad_scale is not None else None
    _check(lib().tta_penalty_bwd_multi(tab.dev_ptr, tab.host_ptr, tab.n, float(rho), p, int(bool(accumulate)),
                                       stream_handle()), 'tta_penalty_bwd_multi')


def unfold_add(tab):
    _check(lib().tta_unfold_add_batched(tab.dev_ptr, tab.host_ptr, tab.n, stream_handle()), 'tta_unfold_add_batched')


def fold_store(tab):
    _check(lib().tta_fold_store_batched(tab.dev_ptr, tab.host_ptr, tab.n, stream_handle()), 'tta_fold_store_batched')


def gram(tab):
    _check(lib().tta_gram_batched(tab.dev_ptr, tab.host_ptr, tab.n, stream_handle()), 'tta_gram_batched')


def jacobi_scratch_bytes(tab):
    return int(lib().tta_jacobi_scratch_bytes(tab.host_ptr, tab.n))


def jacobi_eigh_async(tab, scratch, tol=5e-7, max_sweeps=40):
    """Enqueue only (no host synchronisation); collect with `jacobi_results` after a stream sync.
    Problems that need the multi-launch solver (k > 512) still synchronise inside the call."""
    _check(lib().tta_jacobi_eigh_batched(tab.dev_ptr, tab.host_ptr, tab.n, float(tol), int(max_sweeps),
                                         ctypes.c_void_p(scratch.data_ptr()),
                                         scratch.numel() * scratch.element_size(), None, stream_handle()),
           'tta_jacobi_eigh_batched')


def jacobi_results(tab, scratch, max_sweeps=40):
    """Sweep counts of an earlier `jacobi_eigh_async` (raises TtaError if a problem did not converge)."""
    sweeps = np.zeros(max(tab.n, 1), dtype=np.int32)
    if tab.n == 0:
        return sweeps[:0]
    host = scratch[:6 * tab.n].cpu().numpy() if _FAKE is None else scratch[:6 * tab.n].numpy()
    return jacobi_results_from_host(host, tab, max_sweeps)


def jacobi_results_from_host(host, tab, max_sweeps=40):
    """Same, from a host copy of the first 6 * n int32 of the scratch buffer (one D2H copy can serve many waves)."""
    sweeps = np.zeros(max(tab.n, 1), dtype=np.int32)
    host = np.ascontiguousarray(host, dtype=np.int32)
    _check(lib().tta_jacobi_read_results(ctypes.c_void_p(host.ctypes.data), tab.host_ptr, tab.n, int(max_sweeps),
                                         ctypes.c_void_p(sweeps.ctypes.data)), 'tta_jacobi_read_results')
    return sweeps[:tab.n]


def jacobi_eigh(tab, scratch, tol=5e-7, max_sweeps=40):
    sweeps = np.zeros(max(tab.n, 1), dtype=np.int32)
    _check(lib().tta_jacobi_eigh_batched(tab.dev_ptr, tab.host_ptr, tab.n, float(tol), int(max_sweeps),
                                         ctypes.c_void_p(scratch.data_ptr()),
                                         scratch.numel() * scratch.element_size(),
                                         ctypes.c_void_p(sweeps.ctypes.data), stream_handle()),
           'tta_jacobi_eigh_batched')
    return sweeps[:tab.n]


def select(tab):
    _check(lib().tta_select_batched(tab.dev_ptr, tab.host_ptr, tab.n, stream_handle()), 'tta_select_batched')


def gemm(tab):
    _check(lib().tta_gemm_batched(tab.dev_ptr, tab.host_ptr, tab.n, stream_handle()), 'tta_gemm_batched')


def gemm_f64(tab):
    _check(lib().tta_gemm_f64_batched(tab.dev_ptr, tab.host_ptr, tab.n, stream_handle()), 'tta_gemm_f64_batched')


def refine_prepare(tab):
    _check(lib().tta_refine_prepare_batched(tab.dev_ptr, tab.host_ptr, tab.n, stream_handle()),
           'tta_refine_prepare_batched')


def refine_coeff(tab):
    _check(lib().tta_refine_coeff_batched(tab.dev_ptr, tab.host_ptr, tab.n, stream_handle()),
           'tta_refine_coeff_batched')


def refine_finalize(tab):
    _check(lib().tta_refine_finalize_batched(tab.dev_ptr, tab.host_ptr, tab.n, stream_handle()),
           'tta_refine_finalize_batched')


def orth_penalty_fwd(tab, rho, loss_out):
    _check(lib().tta_orth_penalty_fwd_batched(tab.dev_ptr, tab.host_ptr, tab.n, float(rho),
                                              ctypes.c_void_p(loss_out.data_ptr()), stream_handle()),
           'tta_orth_penalty_fwd_batched')


def orth_penalty_bwd(tab, rho, grad_scale, accumulate):
    p = ctypes.c_void_p(grad_scale.data_ptr()) if grad_scale is not None else None
    _check(lib().tta_orth_penalty_bwd_batched(tab.dev_ptr, tab.host_ptr, tab.n, float(rho), p, int(bool(accumulate)),
                                              stream_handle()), 'tta_orth_penalty_bwd_batched')


def symeig_top(tab):
    """fp64 dominant-r eigensolver (tridiagonalisation route) for every task of the table; enqueue only."""
    _check(lib().tta_symeig_top_batched(tab.dev_ptr, tab.host_ptr, tab.n, stream_handle()), 'tta_symeig_top_batched')


def symeig_work_doubles(k, r):
    return int(lib().tta_symeig_work_doubles(int(k), int(r)))


def symeig_max_k():
    return int(lib().tta_symeig_max_k())


def symeig_profile(enable):
    if _FAKE is None:
        lib().tta_symeig_profile_enable(int(bool(enable)))


def symeig_profile_read():
    """(summed device ms of the tridiagonalisation launches, number of launches) since the last read."""
    ms = ctypes.c_double(0.0)
    n = ctypes.c_ulonglong(0)
    if _FAKE is None:
        lib().tta_symeig_profile_read(ctypes.byref(ms), ctypes.byref(n))
    return ms.value, int(n.value)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def gemm_bf16_tc(a, b, c, M, N, K, lda=None, ldb=None, ldc=None, bias=None):
    """c[M,N] = a[M,K] @ b[N,K]^T (+bias) on tcgen05; a, b bf16 tensors, c bf16 or fp32."""
    _check(lib().tta_gemm_bf16_tc(_p(a), lda if lda is not None else K, _p(b), ldb if ldb is not None else K, _p(c),
                                  ldc if ldc is not None else N, int(M), int(N), int(K), _p(bias),
                                  int(c.dtype == torch.float32), stream_handle()), 'tta_gemm_bf16_tc')


def small_gemm(a, b, c, M, N, K, m_inner=None, s_outer=None, s_inner=0, s_col=1, bias=None, bias_inner=0, bias_col=1,
               a_inner=1, a_outer=None):
    if m_inner is None:
        m_inner, s_outer, s_inner = 1, N, 0
    if a_outer is None:
        a_outer = K * a_inner
    _check(lib().tta_small_gemm(_p(a), int(a.dtype == torch.float32), _p(b), _p(c), int(c.dtype == torch.float32),
                                _p(bias), int(M), int(N), int(K), int(a_inner), int(a_outer), int(m_inner),
                                int(s_outer), int(s_inner), int(s_col),
                                int(bias_inner), int(bias_col), stream_handle()), 'tta_small_gemm')


def cast_bf16(x, y):
    _check(lib().tta_cast_bf16(_p(x), _p(y), x.numel(), stream_handle()), 'tta_cast_bf16')


def nchw_to_nhwc_bf16(x, y, B, C, HW, ldc):
    _check(lib().tta_nchw_to_nhwc_bf16(_p(x), _p(y), B, C, HW, ldc, stream_handle()), 'tta_nchw_to_nhwc_bf16')


def nhwc_to_nchw_f32(x, y, bias, B, C, HW, ldc):
    _check(lib().tta_nhwc_to_nchw_f32(_p(x), int(x.dtype == torch.float32), _p(y), _p(bias), B, C, HW, ldc,
                                      stream_handle()), 'tta_nhwc_to_nchw_f32')


def im2col_bf16(x, out, B, H, W, C, ldx, KH, KW, sh, sw, ph, pw, dh, dw, Ho, Wo, ldo):
    _check(lib().tta_im2col_bf16(_p(x), _p(out), B, H, W, C, ldx, KH, KW, sh, sw, ph, pw, dh, dw, Ho, Wo, ldo,
                                 stream_handle()), 'tta_im2col_bf16')


def lowrank2_fwd(x, w1, w2, bias, y, M, K1, N1, N2, ldx=None, ld1=None, ld2=None, ldy=None):
    """y[M,N2] = bf16(x[M,K1] @ w1[N1,K1]^T) @ w2[N2,N1]^T + bias: one fused tcgen05 kernel (x, w1, w2 bf16)."""
    _check(lib().tta_lowrank2_fwd(_p(x), ldx if ldx is not None else K1, _p(w1), ld1 if ld1 is not None else K1,
                                  _p(w2), ld2 if ld2 is not None else N1, _p(bias), _p(y),
                                  ldy if ldy is not None else N2, int(y.dtype == torch.float32), int(M), int(K1),
                                  int(N1), int(N2), stream_handle()), 'tta_lowrank2_fwd')


def ttconv_fused_fwd(x, a_in, kern, a_out, bias, y, B, Cin, H, W, Ra, Rb, Cout, KS, stride, pad):
    _check(lib().tta_ttconv_fused_fwd(_p(x), _p(a_in), _p(kern), _p(a_out), _p(bias), _p(y), int(B), int(Cin), int(H),
                                      int(W), int(Ra), int(Rb), int(Cout), int(KS), int(stride), int(pad),
                                      stream_handle()), 'tta_ttconv_fused_fwd')


def ttconv_tc_supported(Cin, Ra, Rb, Cout, KS, stride, pad):
    """True when the bf16 tcgen05 kernel (csrc/ttconv_tc.cu) serves this geometry."""
    if _FAKE is not None:
        return False
    return bool(lib().tta_ttconv_tc_supported(int(Cin), int(Ra), int(Rb), int(Cout), int(KS), int(stride), int(pad)))


_TN_WS = {}


def gemm_bf16_tn(a, b, out, M, N, K, lda=None, ldb=None, ldc=None):
    """out[M, N] (fp32) = a^T b with a (K x M), b (K x N) bf16 row-major: the weight-gradient products of the fused
    training path on the tcgen05 kernel of csrc/gemm_tn.cu (TMA boxes as MN-major operands, split-K)."""
    need = int(lib().tta_gemm_bf16_tn_workspace_bytes(int(M), int(N), int(K)))
    ws = _TN_WS.get(a.device)
    if ws is None or ws.numel() < need:
        ws = _TN_WS[a.device] = torch.empty(need, dtype=torch.uint8, device=a.device)
    _check(lib().tta_gemm_bf16_tn(_p(a), lda if lda is not None else M, _p(b), ldb if ldb is not None else N, _p(out),
                                  ldc if ldc is not None else N, int(M), int(N), int(K), _p(ws), ws.numel(), stream_handle()),
           'tta_gemm_bf16_tn')
    return out


def ttconv_tc_pack(a_in, kern, a_out, bias):
    """Weight image of the tensor-core fused convolution (bf16 planes + fp32 bias) as a uint8 device tensor."""
    ra, cin = a_in.shape
    rb, cout = kern.shape[0], a_out.shape[0]
    blob = torch.empty(int(lib().tta_ttconv_tc_blob_bytes(int(cin), int(ra), int(rb), int(cout))), dtype=torch.uint8,
                       device=a_in.device)
    _check(lib().tta_ttconv_tc_pack(_p(a_in), _p(kern), _p(a_out), _p(bias), _p(blob), int(cin), int(ra), int(rb), int(cout),
                                    stream_handle()), 'tta_ttconv_tc_pack')
    return blob


def ttconv_tc_fwd(x, blob, y, B, Cin, H, W, Ra, Rb, Cout, KS, stride, pad):
    _check(lib().tta_ttconv_tc_fwd(_p(x), _p(blob), _p(y), int(B), int(Cin), int(H), int(W), int(Ra), int(Rb), int(Cout),
                                   int(KS), int(stride), int(pad), stream_handle()), 'tta_ttconv_tc_fwd')


def ttconv_tc_fwd_raw(x, blob, y, B, Cin, H, W, Ra, Rb, Cout, KS, stride, pad):
    rc = lib().tta_ttconv_tc_fwd(x, blob, y, B, Cin, H, W, Ra, Rb, Cout, KS, stride, pad,
                                 torch.cuda.current_stream().cuda_stream)
    if rc:
        _check(rc, 'tta_ttconv_tc_fwd')


def ttconv_fused_fwd_raw(x, a_in, kern, a_out, bias, y, B, Cin, H, W, Ra, Rb, Cout, KS, stride, pad):
    """Same call with raw device addresses (ints) -- the per-layer hot path of a network forward."""
    rc = lib().tta_ttconv_fused_fwd(x, a_in, kern, a_out, bias, y, B, Cin, H, W, Ra, Rb, Cout, KS, stride, pad,
                                    torch.cuda.current_stream().cuda_stream if _FAKE is None else None)
    if rc:
        _check(rc, 'tta_ttconv_fused_fwd')


def launch_count():
    return int(lib().tta_launch_count()) if _FAKE is None else 0


def jacobi_force_multilaunch(on):
    if _FAKE is None:
        lib().tta_jacobi_force_multilaunch(int(bool(on)))


def gemm_enable_tc(mode):
    """0 / False: every fp32 GEMM task on the CUDA-core kernel; 1 / True (library default): tasks of >= 0.6 GFLOP on the
    tcgen05 3xTF32 kernel (csrc/gemm_tf32.cu: fresh TMEM accumulator every 32 reduction indices); 2: every task (tests)."""
    if _FAKE is None:
        lib().tta_gemm_enable_tc(int(mode))


def gram_enable_tc(on):
    """False keeps every Gram task on the fp64 CUDA-core kernels (the library default routes TMA-addressable operands
    to the tcgen05 3xTF32 kernel of csrc/gram_tc.cu)."""
    if _FAKE is None:
        lib().tta_gram_enable_tc(int(bool(on)))


def jacobi_enable_gra(on):
    """Test hook: False keeps 32 < k <= 512 problems on the column-rotation cluster kernel."""
    if _FAKE is None:
        lib().tta_jacobi_enable_gra(int(bool(on)))


def jacobi_set_stop_rel(stop_rel):
    if _FAKE is None:
        lib().tta_jacobi_set_stop_rel(float(stop_rel))


def jacobi_profile(enable):
    if _FAKE is None:
        lib().tta_jacobi_profile_enable(int(bool(enable)))


def jacobi_profile_read():
    """(device ms spent in jacobi_step launch sequences, number of jacobi_step launches) since last read."""
    ms = ctypes.c_double(0.0)
    n = ctypes.c_ulonglong(0)
    if _FAKE is None:
        lib().tta_jacobi_profile_read(ctypes.byref(ms), ctypes.byref(n))
    return ms.value, int(n.value)


def sqnorm(tab, out):
    _check(lib().tta_sqnorm_batched(tab.dev_ptr, tab.host_ptr, tab.n, ctypes.c_void_p(out.data_ptr()),
                                    stream_handle()), 'tta_sqnorm_batched')
