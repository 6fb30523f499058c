"""Shared helpers for the parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
N_PROBE = 32


def probe_index(numel, seed=0):
    """Same probe positions as oracle/gen_golden.py."""
    rng = np.random.RandomState(seed + numel % 9973)
    return rng.randint(0, numel, size=N_PROBE)


def rel_fro(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def check_summary(z, gold, rtol_norm=2e-5, probe_tol=1e-4):
    """Compare a tensor with the golden (fro, sum, probe) record of the reference's tensor."""
    z = np.asarray(z, dtype=np.float32).reshape(-1)
    fro = float(np.sqrt(np.sum(z.astype(np.float64) ** 2)))
    assert abs(fro - gold['fro']) <= rtol_norm * max(gold['fro'], 1e-30), (fro, gold['fro'])
    probe = z[probe_index(z.size)].astype(np.float64)
    ref = np.asarray(gold['probe'], dtype=np.float64)
    # per-entry scale of the tensor: fro / sqrt(numel)
    scale = gold['fro'] / np.sqrt(z.size)
    assert np.max(np.abs(probe - ref)) <= probe_tol * scale * 8 + 1e-12, float(np.max(np.abs(probe - ref)) / scale)
