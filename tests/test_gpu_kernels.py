"""Kernel-level parity through the C ABI on a B200: each batched operator of include/tta.h against
numpy on the same seeded inputs (bit-exact where the arithmetic is order-free, tolerance stated)."""
import numpy as np
import pytest
import torch

import tta_runtime as rt

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _ew_table(ws, zs, us, gs=None):
    arr = np.zeros(len(ws), dtype=rt.EW_TASK)
    for i in range(len(ws)):
        arr[i] = (ws[i].data_ptr(), zs[i].data_ptr(), us[i].data_ptr(), gs[i].data_ptr() if gs else 0, ws[i].numel())
    return rt.TaskTable(arr, DEV)


@pytest.mark.parametrize('offset', [0, 1])
def test_dual_update_bit_exact_ragged(offset):
    rng = np.random.RandomState(0)
    sizes = [1, 7, 4095, 4096, 4097, 36864, 1000003]
    W = [rng.randn(n + offset).astype(np.float32) for n in sizes]
    Z = [rng.randn(n + offset).astype(np.float32) for n in sizes]
    U = [rng.randn(n + offset).astype(np.float32) for n in sizes]
    w = [_t(a)[offset:] for a in W]
    z = [_t(a)[offset:] for a in Z]
    u = [_t(a)[offset:] for a in U]
    sq = torch.zeros(len(sizes), dtype=torch.float64, device=DEV)
    rt.dual_update(_ew_table(w, z, u), sq)
    torch.cuda.synchronize()
    for i in range(len(sizes)):
        d = W[i][offset:] - Z[i][offset:]
        assert np.array_equal(u[i].cpu().numpy(), U[i][offset:] + d)          # bit-exact (integer/byte bar)
        ref = float(np.sum(d.astype(np.float64) ** 2))
        assert abs(float(sq[i]) - ref) <= 1e-6 * ref + 1e-30                   # fp32 partials, fp64 tree


def test_penalty_forward_backward():
    rng = np.random.RandomState(1)
    sizes = [5, 4096 * 3, 123457]
    W, Z, U = ([rng.randn(n).astype(np.float32) for n in sizes] for _ in range(3))
    w, z, u = [_t(a) for a in W], [_t(a) for a in Z], [_t(a) for a in U]
    rho = 1e-3
    loss = torch.zeros(1, dtype=torch.float64, device=DEV)
    rt.penalty_fwd(_ew_table(w, z, u), rho, loss)
    ref = sum(0.5 * float(np.float32(rho)) * float(np.sum((a - b + c).astype(np.float64) ** 2)) for a, b, c in zip(W, Z, U))
    assert abs(float(loss) - ref) <= 1e-6 * ref                                 # tolerance: 1e-6 relative
    g = [torch.full_like(a, 2.0) for a in w]
    scale = torch.tensor([3.0], device=DEV)
    rt.penalty_bwd(_ew_table(w, z, u, g), rho, scale, accumulate=False)
    coef = np.float32(np.float32(rho) * np.float32(3.0))
    for i in range(len(sizes)):
        assert np.array_equal(g[i].cpu().numpy(), coef * (W[i] - Z[i] + U[i]))  # overwrite: bit-exact
    rt.penalty_bwd(_ew_table(w, z, u, g), rho, scale, accumulate=True)
    for i in range(len(sizes)):
        ref_g = 2.0 * coef * (W[i] - Z[i] + U[i])
        assert np.allclose(g[i].cpu().numpy(), ref_g, rtol=1e-6, atol=1e-9)    # FMA contraction allowed


@pytest.mark.parametrize('O,I,KK', [(16, 16, 9), (64, 37, 9), (8, 600, 9), (12, 20, 1), (5, 33, 4), (3, 7, 25)])
def test_unfold_fold_exact(O, I, KK):
    rng = np.random.RandomState(2)
    W = rng.randn(O, I, KK).astype(np.float32)
    U = rng.randn(O, I, KK).astype(np.float32)
    w, u = _t(W), _t(U)
    t = torch.empty(O * KK * I, device=DEV)
    z = torch.empty(O * I * KK, device=DEV)
    tab = np.zeros(1, dtype=rt.FOLD_TASK)
    tab[0] = (w.data_ptr(), u.data_ptr(), t.data_ptr(), z.data_ptr(), O, I, KK, 0)
    table = rt.TaskTable(tab, DEV)
    rt.unfold_add(table)
    assert np.array_equal(t.cpu().numpy().reshape(O, KK, I), (W + U).transpose(0, 2, 1))
    rt.fold_store(table)
    assert np.array_equal(z.cpu().numpy().reshape(O, I, KK), W + U)


def _gram_task(a, x, part, k, si, sb, sc, nb, nc, nsplit, ld, kpad, a2=None, g64=None):
    tab = np.zeros(1, dtype=rt.GRAM_TASK)
    tab[0] = (a.data_ptr(), part.data_ptr(), x.data_ptr() if x is not None else 0, g64.data_ptr() if g64 is not None else 0,
              si, sb, sc, k, nb, nc, nsplit, ld, kpad, a2.data_ptr() if a2 is not None else 0)
    return rt.TaskTable(tab, DEV)


@pytest.mark.parametrize('m,n,nsplit', [(8, 4608, 9), (64, 576, 2), (75, 512, 1), (130, 512, 3), (480, 1000, 2)])
def test_gram_row_and_col(m, n, nsplit):
    rng = np.random.RandomState(3)
    A = rng.randn(m, n).astype(np.float32)
    a = _t(A)
    for mode in ('row', 'col'):
        k = m if mode == 'row' else n
        if k > 600:
            continue
        ld, kpad = (k + 3) // 4 * 4, (k + 15) // 16 * 16
        x = torch.full((kpad * ld,), 7.0, device=DEV)
        part = torch.empty(nsplit * k * k, dtype=torch.float64, device=DEV)
        if mode == 'row':
            tab = _gram_task(a, x, part, k, n, 0, 1, 1, n, nsplit, ld, kpad)
            G = A.astype(np.float64) @ A.astype(np.float64).T
        else:
            tab = _gram_task(a, x, part, k, 1, 0, n, 1, m, nsplit, ld, kpad)
            G = A.astype(np.float64).T @ A.astype(np.float64)
        rt.gram_enable_tc(False)             # the fp64 CUDA-core kernels (operands TMA cannot address use them)
        try:
            rt.gram(tab)
        finally:
            rt.gram_enable_tc(True)
        X = x.cpu().numpy().reshape(kpad, ld)
        assert np.allclose(X[:k, :k], G.astype(np.float32).T, rtol=2e-7, atol=1e-12 * np.abs(G).max())  # fp64 accumulation
        assert not X[k:].any() and not X[:, k:].any()                 # zero padding


@pytest.mark.parametrize('m,n,nsplit', [(8, 4608, 9), (32, 73728, 48), (64, 576, 2), (75, 512, 1), (130, 512, 3),
                                        (480, 4608, 18), (512, 948, 4), (300, 100, 1), (1680, 32, 3), (96, 36, 1),
                                        (64, 200, 4)])     # last: more slices than 128-index chunks (empty slices)
@pytest.mark.parametrize('with_u', [False, True])
def test_gram_tensor_core(m, n, nsplit, with_u):
    """tcgen05 3xTF32 Gram (csrc/gram_tc.cu), TMA-fed, the optional second addend summed in shared memory: row and
    column Grams against fp64, error relative to sqrt(g_ii g_jj) (the accumulator is flushed every 32 indices, so the
    error does not grow with the reduction length: 1e-7 grade, tolerance 5e-7)."""
    rng = np.random.RandomState(m + n)
    A = rng.randn(m, n).astype(np.float32)
    U = (0.3 * rng.randn(m, n)).astype(np.float32) if with_u else None
    a = _t(A)
    u = _t(U) if with_u else None
    V = (A + U) if with_u else A
    V64 = V.astype(np.float64)
    for mode in ('row', 'col'):
        k = m if mode == 'row' else n
        if k > 640 or (k if mode == 'col' else n) % 4:
            continue
        ld, kpad = (k + 3) // 4 * 4, (k + 15) // 16 * 16
        x = torch.full((kpad * ld,), 7.0, device=DEV)
        g64 = torch.full((k * k,), 7.0, dtype=torch.float64, device=DEV)
        part = torch.empty(nsplit * k * k, dtype=torch.float64, device=DEV)
        if mode == 'row':
            tab = _gram_task(a, x, part, k, n, 0, 1, 1, n, nsplit, ld, kpad, a2=u, g64=g64)
            G = V64 @ V64.T
        else:
            tab = _gram_task(a, x, part, k, 1, 0, n, 1, m, nsplit, ld, kpad, a2=u, g64=g64)
            G = V64.T @ V64
        n0 = rt.launch_count()
        rt.gram(tab)
        torch.cuda.synchronize()
        assert rt.launch_count() - n0 == 2            # gram_tc_kernel + gram_finish: no CUDA-core partial pass
        assert bool((a == _t(A)).all())               # operands untouched (TMA reads only)
        got = g64.cpu().numpy().reshape(k, k)
        scale = np.sqrt(np.outer(np.diag(G), np.diag(G)))
        err = np.abs(got - G) / scale
        assert err.max() <= 5e-7, (mode, err.max())
        assert np.array_equal(got, got.T)
        X = x.cpu().numpy().reshape(kpad, ld)
        assert np.array_equal(X[:k, :k], got.astype(np.float32).T)
        assert not X[k:].any() and not X[:, k:].any()


def test_gram_second_addend_on_cuda_cores():
    """a2 on the fp64 CUDA-core kernels (operand not TMA-addressable: odd pitch)."""
    rng = np.random.RandomState(12)
    m, n = 40, 1001
    A = rng.randn(m, n).astype(np.float32)
    U = rng.randn(m, n).astype(np.float32)
    a, u = _t(A), _t(U)
    k, ld, kpad = m, 40, 48
    x = torch.empty(kpad * ld, device=DEV)
    part = torch.empty(2 * k * k, dtype=torch.float64, device=DEV)
    rt.gram(_gram_task(a, x, part, k, n, 0, 1, 1, n, 2, ld, kpad, a2=u))
    V = (A + U).astype(np.float64)
    G = V @ V.T
    assert np.allclose(x.cpu().numpy().reshape(kpad, ld)[:k, :k], G.astype(np.float32).T, rtol=2e-7, atol=1e-12 * np.abs(G).max())


def test_gram_mode_layout():
    """Tucker mode Gram: tensor (B, k, c), G = sum_b M_b M_b^T."""
    rng = np.random.RandomState(4)
    B, k, c = 11, 40, 9
    Y = rng.randn(B, k, c).astype(np.float32)
    y = _t(Y)
    ld, kpad = 40, 48
    x = torch.empty(kpad * ld, device=DEV)
    part = torch.empty(2 * k * k, dtype=torch.float64, device=DEV)
    rt.gram(_gram_task(y, x, part, k, c, k * c, 1, B, c, 2, ld, kpad))
    G = np.einsum('bic,bjc->ij', Y.astype(np.float64), Y.astype(np.float64))
    assert np.allclose(x.cpu().numpy().reshape(kpad, ld)[:k, :k], G.astype(np.float32).T, rtol=2e-7, atol=1e-12 * np.abs(G).max())


def _eig_problem(k, rng, spectrum='flat'):
    n = 3 * k
    A = rng.randn(k, n).astype(np.float64)
    if spectrum == 'decay':
        A = A * np.logspace(0, -2, k)[:, None]
    G = (A @ A.T).astype(np.float32)
    return G


@pytest.mark.parametrize('ks', [[8], [64, 75, 16], [130, 240], [512], [256, 32, 480]])
def test_jacobi_eigh_and_select(ks):
    import projector
    rng = np.random.RandomState(5)
    Gs = [_eig_problem(k, rng, 'decay' if i % 2 else 'flat') for i, k in enumerate(ks)]
    etab = np.zeros(len(ks), dtype=rt.EIG_TASK)
    stab = np.zeros(len(ks), dtype=rt.SELECT_TASK)
    bufs = []
    for i, (k, G) in enumerate(zip(ks, Gs)):
        ld, kpad, bw = projector.eig_geometry(k)
        X = np.zeros((kpad, ld), dtype=np.float32)
        X[:k, :k] = G.T
        x = _t(X.reshape(-1))
        r = max(1, k // 3)
        e = torch.empty(r * k, device=DEV)
        et = torch.empty(k * r, device=DEV)
        se = torch.empty(r * k, device=DEV)
        sg = torch.empty(r, device=DEV)
        isg = torch.empty(r, device=DEV)
        etab[i] = (x.data_ptr(), k, ld, kpad, bw)
        stab[i] = (x.data_ptr(), e.data_ptr(), et.data_ptr(), se.data_ptr(), sg.data_ptr(), isg.data_ptr(), k, ld, r, 0)
        bufs.append((x, e, et, se, sg, isg, r, ld, kpad))
    et_tab = rt.TaskTable(etab, DEV)
    scratch = torch.empty(rt.jacobi_scratch_bytes(et_tab) // 4 + 16, dtype=torch.int32, device=DEV)
    sweeps = rt.jacobi_eigh(et_tab, scratch, tol=5e-7, max_sweeps=40)
    rt.select(rt.TaskTable(stab, DEV))
    torch.cuda.synchronize()
    for i, (k, G) in enumerate(zip(ks, Gs)):
        x, e, et, se, sg, isg, r, ld, kpad = bufs[i]
        assert 1 <= sweeps[i] <= 20, sweeps
        X = x.cpu().numpy().reshape(kpad, ld)[:k, :k].astype(np.float64)   # rows = columns x_j
        lam = np.linalg.norm(X, axis=1)
        ref = np.linalg.eigvalsh(G.astype(np.float64))[::-1]
        # column norms are the pre-refinement eigenvalue estimates: the rotations are applied as fp32
        # 32 x 32 matrices, whose rounding is coherent over the rows of a column (5e-5 * lambda_max)
        assert np.max(np.abs(np.sort(lam)[::-1] - ref)) <= 5e-5 * ref[0]
        E = e.cpu().numpy().reshape(r, k).astype(np.float64)
        assert np.max(np.abs(E @ E.T - np.eye(r))) <= 2e-5                   # orthonormal rows
        # invariant-subspace residual ||G E^T - E^T (E G E^T)|| relative to lambda_max
        Gd = G.astype(np.float64)
        res = Gd @ E.T - E.T @ (E @ Gd @ E.T)
        assert np.linalg.norm(res, 2) <= 5e-5 * ref[0]
        assert np.array_equal(et.cpu().numpy().reshape(k, r), e.cpu().numpy().reshape(r, k).T)
        s = sg.cpu().numpy()
        assert np.allclose(s ** 2, ref[:r], rtol=1e-4, atol=1e-5 * ref[0])
        assert np.allclose(se.cpu().numpy().reshape(r, k), e.cpu().numpy().reshape(r, k) * s[:, None], rtol=1e-6, atol=0)
        assert np.allclose(isg.cpu().numpy() * s, 1.0, rtol=1e-5)


def test_jacobi_rank_deficient_and_zero():
    rng = np.random.RandomState(6)
    k, rank = 96, 20
    B = rng.randn(k, rank)
    G = (B @ B.T).astype(np.float32)
    Z = np.zeros((32, 32), dtype=np.float32)
    etab = np.zeros(2, dtype=rt.EIG_TASK)
    xs = []
    for i, M in enumerate((G, Z)):
        kk = M.shape[0]
        X = np.zeros((kk, kk), dtype=np.float32)
        X[:] = M.T
        x = _t(X.reshape(-1))
        xs.append(x)
        etab[i] = (x.data_ptr(), kk, kk, kk, 16)
    tab = rt.TaskTable(etab, DEV)
    scratch = torch.empty(rt.jacobi_scratch_bytes(tab) // 4 + 16, dtype=torch.int32, device=DEV)
    sweeps = rt.jacobi_eigh(tab, scratch, tol=5e-7, max_sweeps=40)
    assert sweeps[0] <= 25 and sweeps[1] == 1
    r = 30
    e = torch.empty(r * k, device=DEV)
    stab = np.zeros(1, dtype=rt.SELECT_TASK)
    stab[0] = (xs[0].data_ptr(), e.data_ptr(), 0, 0, 0, 0, k, k, r, 0)
    rt.select(rt.TaskTable(stab, DEV))
    E = e.cpu().numpy().reshape(r, k).astype(np.float64)
    P = E.T @ E
    Gd = G.astype(np.float64)
    assert np.linalg.norm(P @ Gd - Gd) <= 1e-4 * np.linalg.norm(Gd)      # spans the range; extra rows are zero
    assert np.count_nonzero(np.linalg.norm(E, axis=1) > 0.5) == rank


@pytest.mark.parametrize('M,N,K', [(8, 4608, 8), (105, 513, 480), (300, 70, 15), (1, 1, 1), (64, 64, 64)])
def test_gemm_strides(M, N, K):
    rng = np.random.RandomState(7)
    A = rng.randn(M, K).astype(np.float32)
    B = rng.randn(K, N).astype(np.float32)
    cs = rng.rand(N).astype(np.float32) + 0.5
    ref = A.astype(np.float64) @ B.astype(np.float64)
    tol = 2e-6 * np.sqrt(K) * np.abs(A).max() * np.abs(B).max() * 4
    for ta in (False, True):
        for tb in (False, True):
            a = _t(A.T if ta else A)
            b = _t(B.T if tb else B)
            c = torch.zeros(M, N + 3, device=DEV)
            s = _t(cs)
            tab = np.zeros(2, dtype=rt.GEMM_TASK)
            sai, sak = (1, M) if ta else (K, 1)
            sbk, sbj = (1, K) if tb else (N, 1)
            tab[0] = (a.data_ptr(), b.data_ptr(), c.data_ptr(), 0, sai, sak, sbk, sbj, N + 3, M, N, K, 0)
            c2 = torch.zeros(M, N, device=DEV)
            tab[1] = (a.data_ptr(), b.data_ptr(), c2.data_ptr(), s.data_ptr(), sai, sak, sbk, sbj, N, M, N, K, 0)
            rt.gemm(rt.TaskTable(tab, DEV))
            out = c.cpu().numpy()
            assert np.max(np.abs(out[:, :N] - ref)) <= tol
            assert not out[:, N:].any()
            assert np.max(np.abs(c2.cpu().numpy() - ref * cs[None, :])) <= tol * 1.5


@pytest.mark.parametrize('M,N,K', [(128, 128, 32), (105, 4608, 480), (945, 105, 512), (2048, 512, 130), (300, 70, 9),
                                   (130, 1024, 130), (4608, 80, 15), (49, 25, 8), (30, 73728, 32), (64, 9, 40),
                                   (257, 129, 33), (1100, 700, 520), (300, 4100, 260), (2305, 131, 1001)])
def test_gemm_tensor_core_3xtf32_is_fp32_grade(M, N, K):
    """Tasks big enough for the tcgen05 kernel: 3xTF32 keeps fp32-grade accuracy (a single TF32 pass would
    be ~3e-4 relative) and agrees with the CUDA-core kernel; every operand stride combination."""
    rng = np.random.RandomState(M + N + K)
    A = rng.randn(M, K).astype(np.float32)
    B = rng.randn(K, N).astype(np.float32)
    cs = (rng.rand(N).astype(np.float32) + 0.5)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    for ta in (False, True):
        for tb in (False, True):
            a = _t(A.T if ta else A)
            b = _t(B.T if tb else B)
            s = _t(cs)
            sai, sak = (1, M) if ta else (K, 1)
            sbk, sbj = (1, K) if tb else (N, 1)
            outs = []
            for tc in (2, 0):          # 2: the tensor-core kernel whatever the size of the task; 0: CUDA cores
                c = torch.full((M, N + 5), 3.0, device=DEV)
                tab = np.zeros(1, dtype=rt.GEMM_TASK)
                tab[0] = (a.data_ptr(), b.data_ptr(), c.data_ptr(), s.data_ptr(), sai, sak, sbk, sbj, N + 5, M, N, K, 0)
                rt.gemm_enable_tc(tc)
                try:
                    rt.gemm(rt.TaskTable(tab, DEV))
                    torch.cuda.synchronize()
                finally:
                    rt.gemm_enable_tc(1)          # library default: by size
                out = c.cpu().numpy()
                assert np.all(out[:, N:] == 3.0)                      # nothing written past column N
                outs.append(out[:, :N].astype(np.float64))
            want = ref * cs[None, :]
            err_tc = np.linalg.norm(outs[0] - want) / np.linalg.norm(want)
            err_cc = np.linalg.norm(outs[1] - want) / np.linalg.norm(want)
            # the TMEM accumulator is drained every 32 reduction indices, so the truncating accumulation of the tensor
            # core does not build up with K (it would reach 3.3e-6 at K = 480); a single TF32 pass would be ~3e-4
            assert err_tc <= 1e-6, (err_tc, ta, tb)
            assert err_cc <= 2e-6, (err_cc, ta, tb)


def test_sqnorm():
    rng = np.random.RandomState(8)
    xs = [rng.randn(n).astype(np.float32) for n in (1, 1000, 300001)]
    ts = [_t(x) for x in xs]
    tab = np.zeros(len(xs), dtype=rt.SQNORM_TASK)
    for i, t in enumerate(ts):
        tab[i] = (t.data_ptr(), t.numel())
    out = torch.zeros(len(xs), dtype=torch.float64, device=DEV)
    rt.sqnorm(rt.TaskTable(tab, DEV), out)
    for i, x in enumerate(xs):
        ref = float(np.sum(x.astype(np.float64) ** 2))
        assert abs(float(out[i]) - ref) <= 1e-12 * ref


@pytest.mark.parametrize('k,r', [(64, 20), (240, 82), (512, 105)])
def test_refinement_reaches_fp64_subspace(k, r):
    """fp32 Jacobi + one fp64 Ogita-Aishima step: dominant-r projector vs numpy fp64 eigh."""
    rng = np.random.RandomState(9)
    A = rng.randn(k, 3 * k)
    G64 = A @ A.T
    import projector
    ld, kpad, bw = projector.eig_geometry(k)
    wnd = projector.refine_window(k, r)
    X = np.zeros((kpad, ld), dtype=np.float32)
    X[:k, :k] = G64.astype(np.float32).T
    x = _t(X.reshape(-1))
    g64 = _t(G64.reshape(-1))
    f64 = dict(dtype=torch.float64, device=DEV)
    qt, y, s, t = (torch.empty(k * k, **f64) for _ in range(4))
    c, e64 = torch.empty(r * k, **f64), torch.empty(r * k, **f64)
    lam = torch.empty(r, **f64)
    lam0 = torch.empty(k, **f64)
    e = torch.empty(r * k, device=DEV)
    et = torch.empty(r * k, device=DEV)
    sg = torch.empty(r, device=DEV)
    etab = np.zeros(1, dtype=rt.EIG_TASK)
    etab[0] = (x.data_ptr(), k, ld, kpad, bw)
    tab = rt.TaskTable(etab, DEV)
    scratch = torch.empty(rt.jacobi_scratch_bytes(tab) // 4 + 16, dtype=torch.int32, device=DEV)
    rt.jacobi_eigh(tab, scratch, tol=2e-6, max_sweeps=40)
    rf = np.zeros(1, dtype=rt.REFINE_TASK)
    rf[0] = (x.data_ptr(), qt.data_ptr(), s.data_ptr(), t.data_ptr(), c.data_ptr(), lam.data_ptr(), lam0.data_ptr(),
             e64.data_ptr(), e.data_ptr(), et.data_ptr(), 0, sg.data_ptr(), 0, k, ld, r, wnd)
    rtab = rt.TaskTable(rf, DEV)

    def dg(a, b, cc, sai, sak, sbk, sbj, ldc, M, N, K):
        tb = np.zeros(1, dtype=rt.GEMM_TASK)
        tb[0] = (a.data_ptr(), b.data_ptr(), cc.data_ptr(), 0, sai, sak, sbk, sbj, ldc, M, N, K, 0)
        rt.gemm_f64(rt.TaskTable(tb, DEV))

    rt.refine_prepare(rtab)
    dg(qt, g64, y, k, 1, k, 1, k, wnd, k, k)      # only the rows that can be selected (first wnd after sorting)
    dg(y, qt, s, k, 1, 1, k, k, wnd, k, k)
    dg(qt, qt, t, k, 1, 1, k, k, wnd, k, k)
    rt.refine_coeff(rtab)
    dg(c, qt, e64, k, 1, k, 1, k, r, k, k)
    rt.refine_finalize(rtab)
    torch.cuda.synchronize()
    lam_ref, v_ref = np.linalg.eigh(G64)
    lam_ref, v_ref = lam_ref[::-1], v_ref[:, ::-1]
    E = e.cpu().numpy().reshape(r, k).astype(np.float64)
    P = E.T @ E
    P_ref = v_ref[:, :r] @ v_ref[:, :r].T
    gap = (lam_ref[r - 1] - lam_ref[r]) / lam_ref[0]
    # fp32 storage of E bounds the projector error at ~1e-7 * sqrt(r); gap amplification is second order
    assert np.linalg.norm(P - P_ref) <= 2e-6 * np.sqrt(r) / max(min(gap * 1e3, 1.0), 1e-3), (np.linalg.norm(P - P_ref), gap)
    assert np.allclose(lam.cpu().numpy(), lam_ref[:r], rtol=1e-7, atol=1e-9 * lam_ref[0])   # Rayleigh quotients
    assert np.allclose(sg.cpu().numpy() ** 2, lam_ref[:r], rtol=1e-6)
    assert np.array_equal(et.cpu().numpy().reshape(k, r), e.cpu().numpy().reshape(r, k).T)


@pytest.mark.parametrize('multilaunch', [False, True])
def test_jacobi_cluster_and_multilaunch_paths_agree(multilaunch):
    """Plan geometry (persistent cluster solver) vs the multi-launch solver on the same matrices."""
    import projector
    rng = np.random.RandomState(11)
    ks = [8, 16, 30, 64, 75, 130, 240, 384, 480, 512]
    Gs = [_eig_problem(k, rng) for k in ks]
    etab = np.zeros(len(ks), dtype=rt.EIG_TASK)
    xs = []
    for i, (k, G) in enumerate(zip(ks, Gs)):
        ld, kpad, bw = projector.eig_geometry(k)
        if multilaunch and bw > 16:
            pytest.skip('multi-launch solver supports bw <= 16')
        X = np.zeros((kpad, ld), dtype=np.float32)
        X[:k, :k] = G.T
        x = _t(X.reshape(-1))
        xs.append((x, ld, kpad))
        etab[i] = (x.data_ptr(), k, ld, kpad, bw)
    tab = rt.TaskTable(etab, DEV)
    rt.jacobi_force_multilaunch(multilaunch)
    try:
        scratch = torch.empty(rt.jacobi_scratch_bytes(tab) // 4 + 16, dtype=torch.int32, device=DEV)
        sweeps = rt.jacobi_eigh(tab, scratch, tol=5e-7, max_sweeps=40)
    finally:
        rt.jacobi_force_multilaunch(False)
    torch.cuda.synchronize()
    for i, (k, G) in enumerate(zip(ks, Gs)):
        x, ld, kpad = xs[i]
        assert 1 <= sweeps[i] <= 24, sweeps
        X = x.cpu().numpy().reshape(kpad, ld).astype(np.float64)
        lam = np.sort(np.linalg.norm(X, axis=1))[::-1]
        ref = np.linalg.eigvalsh(G.astype(np.float64))[::-1]
        assert np.max(np.abs(lam[:k] - ref)) <= 5e-5 * ref[0], (k, multilaunch)
        assert np.all(lam[k:] <= 1e-6 * ref[0])                       # padded columns stay (numerically) zero
        Xn = X[np.argsort(-np.linalg.norm(X, axis=1))[:k]]
        Q = Xn / np.linalg.norm(Xn, axis=1, keepdims=True)
        assert np.max(np.abs(Q @ Q.T - np.eye(k))) <= 5e-6            # columns mutually orthogonal (tol 5e-7)


# ---- fp64 dominant-r eigensolver (csrc/trd.cu) against numpy.linalg.eigh -------------------------------------
def _symeig_run(gs, rs):
    """gs: list of k x k fp64 SPD matrices; returns (lam, E) per problem through tta_symeig_top_batched."""
    order = sorted(range(len(gs)), key=lambda i: -gs[i].shape[0])
    bufs, tab = [], np.zeros(len(gs), dtype=rt.SYMEIG_TASK)
    status = torch.full((len(gs),), 7, dtype=torch.int32, device=DEV)
    for slot, i in enumerate(order):
        k, r = gs[i].shape[0], rs[i]
        g = _t(gs[i])
        work = torch.full((rt.symeig_work_doubles(k, r),), float("nan"), dtype=torch.float64, device=DEV)   # scratch: contents must not matter
        lam = torch.empty(r, dtype=torch.float64, device=DEV)
        e64 = torch.empty(r * k, dtype=torch.float64, device=DEV)
        bufs.append((g, work, lam, e64))
        tab[slot] = (g.data_ptr(), work.data_ptr(), lam.data_ptr(), e64.data_ptr(), status.data_ptr() + 4 * slot, k, r)
    rt.symeig_top(rt.TaskTable(tab, DEV))
    torch.cuda.synchronize()
    out = [None] * len(gs)
    st = status.cpu().numpy()
    for slot, i in enumerate(order):
        k, r = gs[i].shape[0], rs[i]
        out[i] = (bufs[slot][2].cpu().numpy(), bufs[slot][3].cpu().numpy().reshape(r, k), int(st[slot]))
    return out


@pytest.mark.parametrize('ks', [(33, 64, 65), (100, 130, 256), (300, 384), (480, 512), (513, 608), (3, 5, 40)])
def test_symeig_top_matches_eigh(ks):
    """Dominant-r invariant subspace and eigenvalues of flat-spectrum Gram matrices: eigenvalues to 1e-12 |G|,
    orthonormality and the spectral projector to 1e-9 (fp64 end to end)."""
    rng = np.random.RandomState(11)
    gs, rs = [], []
    for k in ks:
        a = rng.randn(k, 3 * k + 7).astype(np.float32).astype(np.float64)
        gs.append(a @ a.T)
        rs.append(max(1, k // 4 + 1))
    for (lam, e, status), g, r in zip(_symeig_run(gs, rs), gs, rs):
        w, v = np.linalg.eigh(g)
        w, v = w[::-1], v[:, ::-1]
        assert status == 0
        assert np.max(np.abs(lam - w[:r])) <= 1e-12 * w[0]
        e = e / np.linalg.norm(e, axis=1, keepdims=True)
        assert np.max(np.abs(e @ e.T - np.eye(r))) <= 1e-9
        assert np.linalg.norm(e.T @ e - v[:, :r] @ v[:, :r].T) <= 1e-9 * max(1.0, w[0] / (w[r - 1] - w[r]) * 1e-3)


def test_symeig_rank_deficient_and_cluster_flag():
    """A rank-deficient Gram matrix keeps its live eigenpairs exact (the null vectors are the caller's to drop);
    exactly repeated eigenvalues above the zero cut raise the status flag."""
    rng = np.random.RandomState(12)
    k = 96
    a = rng.randn(k, 20).astype(np.float64)
    g = a @ a.T                                          # rank 20
    (lam, e, status), = _symeig_run([g], [40])
    w, v = np.linalg.eigh(g)
    w, v = w[::-1], v[:, ::-1]
    assert np.max(np.abs(lam[:20] - w[:20])) <= 1e-12 * w[0]
    assert np.max(np.abs(lam[20:])) <= 1e-12 * w[0]
    e = e[:20] / np.linalg.norm(e[:20], axis=1, keepdims=True)
    assert np.linalg.norm(e.T @ e - v[:, :20] @ v[:, :20].T) <= 1e-8
    q, _ = np.linalg.qr(rng.randn(k, k))
    d = np.linspace(1.0, 2.0, k)
    d[-3:] = 2.0                                          # a triple dominant eigenvalue
    (lam, e, status), = _symeig_run([(q * d) @ q.T], [10])
    assert status == 1
