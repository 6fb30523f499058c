"""CPU model (numpy, fp64) of the tridiagonalisation eigensolver of csrc/trd*.cu -- the arithmetic the GPU
kernels execute, step for step, used to validate the algorithm against numpy.linalg.eigh / the oracle.

  A. Householder tridiagonalisation  G = Q T Q^T   (unblocked, LAPACK dsytd2 'L' convention)
  B. the r largest eigenvalues of T by multisection on the Sturm count (quotient form, pivmin guard)
  C. their eigenvectors by the twisted factorisation of T - lambda I (forward + backward quotient sweeps,
     twist at the smallest |gamma|)
  D. back-transformation x = H_0 H_1 ... H_{k-3} z
"""
import sys

import numpy as np


def sytd2(g):
    a = np.array(g, dtype=np.float64)
    k = a.shape[0]
    d = np.zeros(k)
    e = np.zeros(max(k - 1, 0))
    tau = np.zeros(max(k - 2, 0))
    vs = np.zeros((max(k - 2, 0), k))
    for j in range(k - 2):
        x = a[j + 1:, j].copy()
        alpha = x[0]
        sig = float(np.dot(x[1:], x[1:]))
        d[j] = a[j, j]
        if sig == 0.0:
            e[j] = alpha
            tau[j] = 0.0
            continue
        beta = -np.copysign(np.sqrt(alpha * alpha + sig), alpha)
        t = (beta - alpha) / beta
        v = x / (alpha - beta)
        v[0] = 1.0
        e[j] = beta
        tau[j] = t
        vs[j, j + 1:] = v
        sub = a[j + 1:, j + 1:]
        p = t * (sub @ v)
        kk = 0.5 * t * float(p @ v)
        w = p - kk * v
        sub -= np.outer(v, w) + np.outer(w, v)
    if k >= 2:
        d[k - 2] = a[k - 2, k - 2]
        e[k - 2] = a[k - 1, k - 2]
    d[k - 1] = a[k - 1, k - 1]
    return d, e, tau, vs


def sturm_count(d, e2, x, pivmin):
    """number of eigenvalues < x (vectorised over x)."""
    x = np.asarray(x, dtype=np.float64)
    q = d[0] - x
    q = np.where(np.abs(q) < pivmin, -pivmin, q)
    cnt = (q < 0).astype(np.int64)
    for i in range(1, d.shape[0]):
        q = (d[i] - x) - e2[i - 1] / q
        q = np.where(np.abs(q) < pivmin, -pivmin, q)
        cnt += q < 0
    return cnt


def top_eigvals(d, e, r, nshift=64, npass=9):
    k = d.shape[0]
    e2 = e * e
    pivmin = np.finfo(np.float64).tiny * max(1.0, float(e2.max()) if e2.size else 1.0)
    ea = np.abs(np.concatenate([[0.0], e, [0.0]]))
    gl = float(np.min(d - ea[:-1] - ea[1:]))
    gu = float(np.max(d + ea[:-1] + ea[1:]))
    span = gu - gl
    gl -= 1e-12 * span + 2 * pivmin
    gu += 1e-12 * span + 2 * pivmin
    # eigenvalue index (ascending, 0-based) m: count(x) <= m  <=>  x <= lambda_m
    ms = k - 1 - np.arange(r)
    lo = np.full(r, gl)
    hi = np.full(r, gu)
    for _ in range(npass):
        step = (hi - lo) / (nshift + 1)
        xs = lo[:, None] + step[:, None] * np.arange(1, nshift + 1)[None, :]
        cnt = sturm_count(d, e2, xs.reshape(-1), pivmin).reshape(r, nshift)
        # number of shifts with count <= m: those are <= lambda_m
        nle = np.sum(cnt <= ms[:, None], axis=1)
        new_lo = lo + step * nle
        new_hi = np.where(nle < nshift, lo + step * (nle + 1), hi)
        lo, hi = new_lo, new_hi
    return 0.5 * (lo + hi), pivmin


def twisted_vectors(d, e, lams, pivmin):
    k = d.shape[0]
    r = lams.shape[0]
    s = np.zeros((k, r))
    p = np.zeros((k, r))
    L = np.zeros((max(k - 1, 1), r))
    U = np.zeros((max(k - 1, 1), r))
    q = d[0] - lams
    q = np.where(np.abs(q) < pivmin, -pivmin, q)
    s[0] = q
    for i in range(k - 1):
        L[i] = e[i] / s[i]
        q = (d[i + 1] - lams) - L[i] * e[i]
        q = np.where(np.abs(q) < pivmin, -pivmin, q)
        s[i + 1] = q
    q = d[k - 1] - lams
    q = np.where(np.abs(q) < pivmin, -pivmin, q)
    p[k - 1] = q
    for i in range(k - 2, -1, -1):
        U[i] = e[i] / p[i + 1]
        q = (d[i] - lams) - U[i] * e[i]
        q = np.where(np.abs(q) < pivmin, -pivmin, q)
        p[i] = q
    gamma = s + p - (d[:, None] - lams[None, :])
    tw = np.argmin(np.abs(gamma), axis=0)
    z = np.zeros((k, r))
    for c in range(r):
        t = tw[c]
        z[t, c] = 1.0
        for i in range(t - 1, -1, -1):
            z[i, c] = -L[i, c] * z[i + 1, c]
        for i in range(t + 1, k):
            z[i, c] = -U[i - 1, c] * z[i - 1, c]
    z /= np.linalg.norm(z, axis=0, keepdims=True)
    return z


def back_transform(vs, tau, z):
    x = z.copy()
    for j in range(tau.shape[0] - 1, -1, -1):
        v = vs[j]
        x -= np.outer(tau[j] * v, v @ x)
    return x


def eig_top(g, r):
    """rows = r dominant eigenvectors of g (descending), eigenvalues."""
    d, e, tau, vs = sytd2(g)
    lam, pivmin = top_eigvals(d, e, r)
    z = twisted_vectors(d, e, lam, pivmin)
    x = back_transform(vs, tau, z)
    return x.T, lam


if __name__ == '__main__':
    rng = np.random.default_rng(0)
    k, n, r = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (160, 1536, 40)
    a = rng.standard_normal((k, n)).astype(np.float32).astype(np.float64)
    g = a @ a.T
    E, lam = eig_top(g, r)
    w, v = np.linalg.eigh(g)
    w, v = w[::-1], v[:, ::-1]
    print('eigenvalue rel err', np.max(np.abs(lam - w[:r])) / w[0])
    print('orthonormality', np.max(np.abs(E @ E.T - np.eye(r))))
    pr = E.T @ E
    pt = v[:, :r] @ v[:, :r].T
    print('projector err', np.linalg.norm(pr - pt))
    print('residual', np.max(np.linalg.norm(g @ E.T - E.T * lam[None, :], axis=0)) / w[0])
