"""Time the fused two-factor linear forward (tta_lowrank2_fwd) on the DeiT-small TTLinear shapes.

    python scripts/bench_lowrank2.py
One JSON line per shape: ms, TFLOP/s (both GEMMs), algorithmic HBM GB/s (x in + y out).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200')):
    sys.path.insert(0, p)
import torch

import tta_runtime as rt

DEV = 'cuda:0'


def main():
    M = 256 * 197
    for name, K1, N1, N2 in (('qkv b0', 384, 320, 1152), ('qkv', 384, 256, 1152), ('proj', 384, 256, 384),
                             ('fc1', 384, 256, 1536), ('fc2', 1536, 256, 384), ('fc2 b0', 1536, 320, 384)):
        for out_f32 in (True, False):
            x = torch.randn(M, K1, device=DEV).to(torch.bfloat16)
            w1 = torch.randn(N1, K1, device=DEV).to(torch.bfloat16)
            w2 = torch.randn(N2, N1, device=DEV).to(torch.bfloat16)
            bias = torch.randn(N2, device=DEV)
            y = torch.empty(M, N2, device=DEV, dtype=torch.float32 if out_f32 else torch.bfloat16)
            for _ in range(3):
                rt.lowrank2_fwd(x, w1, w2, bias, y, M, K1, N1, N2)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                rt.lowrank2_fwd(x, w1, w2, bias, y, M, K1, N1, N2)
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1) / 20
            fl = 2.0 * M * N1 * (K1 + N2)
            by = M * (K1 * 2 + N2 * (4 if out_f32 else 2))
            print(json.dumps({'shape': name, 'M': M, 'K1': K1, 'N1': N1, 'N2': N2, 'out': 'f32' if out_f32 else 'bf16',
                              'ms': ms, 'tflops': fl / ms / 1e9, 'hbm_gbs': by / ms / 1e6}), flush=True)


if __name__ == '__main__':
    main()
