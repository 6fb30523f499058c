"""Forward section of bench.py alone (decomposed-layer forwards + the tcgen05 kernels):  python scripts/bench_forward.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200')):
    sys.path.insert(0, p)
import torch

import bench

torch.cuda.set_device(0)
print(json.dumps(bench.forward_bench(torch.device('cuda', 0), bench._peaks())))
