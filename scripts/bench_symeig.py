"""Time tta_symeig_top_batched (csrc/trd.cu) per problem size on one B200 and check it against numpy.linalg.eigh.
usage: python scripts/bench_symeig.py [k ...]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200'))
import tta_runtime as rt  # noqa: E402

DEV = 'cuda:0'


def main():
    ks = [int(a) for a in sys.argv[1:]] or [64, 128, 256, 384, 480, 512, 608]
    rng = np.random.RandomState(0)
    for k in ks:
        r = max(1, int(0.22 * k))
        a = rng.randn(k, 9 * k).astype(np.float32).astype(np.float64)
        g_h = a @ a.T
        g = torch.from_numpy(g_h).to(DEV)
        work = torch.full((rt.symeig_work_doubles(k, r),), float("nan"), dtype=torch.float64, device=DEV)
        lam = torch.zeros(r, dtype=torch.float64, device=DEV)
        e64 = torch.zeros(r * k, dtype=torch.float64, device=DEV)
        status = torch.zeros(1, dtype=torch.int32, device=DEV)
        tab = np.zeros(1, dtype=rt.SYMEIG_TASK)
        tab[0] = (g.data_ptr(), work.data_ptr(), lam.data_ptr(), e64.data_ptr(), status.data_ptr(), k, r)
        table = rt.TaskTable(tab, DEV)
        for _ in range(3):
            rt.symeig_top(table)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        n = 20
        ev[0].record()
        for _ in range(n):
            rt.symeig_top(table)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / n
        w, v = np.linalg.eigh(g_h)
        w, v = w[::-1], v[:, ::-1]
        e = e64.cpu().numpy().reshape(r, k)
        e = e / np.linalg.norm(e, axis=1, keepdims=True)
        perr = np.linalg.norm(e.T @ e - v[:, :r] @ v[:, :r].T)
        lerr = np.max(np.abs(lam.cpu().numpy() - w[:r])) / w[0]
        # d / e of the tridiagonal matrix against the CPU model of the same algorithm
        print('k={:4d} r={:3d}  {:.3f} ms  projector err {:.2e}  eigenvalue err {:.2e}  status {}'.format(
            k, r, ms, perr, lerr, int(status.item())), flush=True)


if __name__ == '__main__':
    main()
