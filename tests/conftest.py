"""pytest wiring: sys.path for the flat drop-in package, the `gpu` marker, and the C-ABI emulator fixture."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'dnn-compression-tensor-admm_b200')
for p in (ROOT, PKG, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def emulated_backend():
    """Route the C ABI to tests/fake_tta.py (host-logic tests on a box without a GPU)."""
    import tta_runtime as rt
    from fake_tta import FakeTTA
    fake = FakeTTA()
    rt.set_backend_for_tests(fake)
    yield fake
    rt.set_backend_for_tests(None)


@pytest.fixture(scope='session')
def golden_summary():
    with open(os.path.join(GOLDEN, 'reference_summary.json')) as f:
        return json.load(f)
