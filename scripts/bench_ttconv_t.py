"""Chunk-size sweep of the tensor-core fused convolution (TTA_TTCONV_T = tiles per CTA) on the ttm_resnet32 stage shapes."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, os
sys.path[:0] = [%r, %r]
import torch, tta_runtime as rt
dev = 'cuda:0'
def t(fn, reps=20, iters=5):
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): g.replay()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / iters / reps * 1e3
out = []
for (cin, hw, ra, rb, cout) in ((16, 32, 16, 16, 16), (32, 16, 32, 32, 32), (64, 8, 64, 64, 64), (64, 56, 40, 40, 64)):
    x = torch.randn(128, cin, hw, hw, device=dev); y = torch.empty(128, cout, hw, hw, device=dev)
    blob = rt.ttconv_tc_pack(torch.randn(ra, cin, device=dev), torch.randn(rb, ra, 3, 3, device=dev), torch.randn(cout, rb, device=dev), None)
    out.append('%%dch %%dx%%d: %%.1f us' %% (cin, hw, hw, t(lambda: rt.ttconv_tc_fwd(x, blob, y, 128, cin, hw, hw, ra, rb, cout, 3, 1, 1))))
print('T=' + os.environ.get('TTA_TTCONV_T', 'auto'), ' | '.join(out))
''' % (ROOT, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200'))
for T in ('', '1', '2', '4', '8'):
    env = dict(os.environ)
    if T:
        env['TTA_TTCONV_T'] = T
    subprocess.run([sys.executable, '-c', CODE], env=env, check=False)
