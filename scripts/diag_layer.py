"""Diagnostic (GPU): per-TT-step subspace error of one layer against numpy fp64 SVDs of the same carries."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200')):
    sys.path.insert(0, p)
import numpy as np
import torch

import hp_tables
import projector
import workloads
from oracle import port


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'layer3.5.conv2.weight'
    tol = float(sys.argv[2]) if len(sys.argv) > 2 else 2e-6
    refine = (sys.argv[3] != '0') if len(sys.argv) > 3 else True
    hp = hp_tables.tt_resnet50_general_3x()
    w = workloads.resnet50_weights()[name]
    dev = 'cuda:0'
    L = projector.TTLayer(name, w.shape, hp.tt_shapes[name], hp.ranks[name])
    plan = projector.TTProjectionPlan([L], dev, tol=tol, refine=refine)
    wt = w.to(dev).contiguous()
    u = torch.zeros_like(wt)
    z = torch.empty_like(wt)
    wn = w.numpy()
    un = np.zeros_like(wn)
    for it in range(3):
        plan.run([wt], [u], [z])
        torch.cuda.synchronize()
        zr = port.project_conv_tt(wn + un, hp.tt_shapes[name], list(hp.ranks[name]))
        zg = z.cpu().numpy()
        print('update', it, 'Z rel err', np.linalg.norm(zg - zr) / np.linalg.norm(zr), 'sweeps', plan.sweeps)
        ws = plan.ws[0]
        for i, st in enumerate(ws['steps']):
            m, n, k, r = st['m'], st['n'], st['k'], st['r']
            A = st['A'].t[:m * n].cpu().numpy().reshape(m, n).astype(np.float64)
            U, s, Vt = np.linalg.svd(A, full_matrices=False)
            E = st['E'].t[:r * k].cpu().numpy().reshape(r, k).astype(np.float64)
            B = U[:, :r] if m <= n else Vt[:r].T
            P_ref = B @ B.T
            P = E.T @ E
            gap = (s[r - 1] - s[r]) / s[0] if r < len(s) else float('nan')
            orth = np.abs(E @ E.T - np.eye(r)).max()
            core = st['core'].t[:m * r].cpu().numpy().reshape(m, r).astype(np.float64)
            corth = np.abs(core.T @ core - np.eye(r)).max()
            carry = st['carry'].t[:r * n].cpu().numpy().reshape(r, n).astype(np.float64)
            rec = np.linalg.norm(core @ carry - B @ (B.T @ A if m <= n else (A @ B).T).reshape(r, -1) if False else 0)
            proj_ref = (B @ (B.T @ A)) if m <= n else (A @ B) @ B.T
            print('   step', i, (m, n, r), 'proj err/sqrt(r) %.3e' % (np.linalg.norm(P - P_ref) / np.sqrt(r)),
                  'sigma gap %.2e' % gap, 'E orth %.1e core orth %.1e' % (orth, corth),
                  'core*carry vs P A: %.3e' % (np.linalg.norm(core @ carry - proj_ref) / np.linalg.norm(proj_ref)))
        if it >= 1:
            u += wt - z
            un = un + (wn - zr)


if __name__ == '__main__':
    main()
