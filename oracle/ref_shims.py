"""Import the UNMODIFIED reference from /root/reference (build container only; test infrastructure).

The reference's flat modules `import tensorly`, `timm`, `geoopt`, none of which is installed here.
`load_reference()` installs minimal `sys.modules` stand-ins and returns the reference modules:

  * tensorly          -> `oracle.port.partial_tucker2` / `tucker2_to_tensor` (restated tensorly<0.8
                         semantics; PARITY UNPINNED for Tucker, see oracle/__init__.py)
  * timm.models.registry.register_model -> identity decorator (resnet_cifar.py:12)
  * geoopt            -> empty module (StfTKConv.py:23, never executed)

The reference modules are registered under the prefix `refimpl_` so that they never shadow the
drop-in modules of the same bare names (`admm`, `ttd`, `TTConv`, ...).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('TTA_REFERENCE_ROOT', '/root/reference')


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'admm.py'))


def _tensorly_stub():
    from . import port

    tl = types.ModuleType('tensorly')
    tl._backend = 'numpy'

    def set_backend(name):
        tl._backend = name

    def tucker_to_tensor(tucker, *a, **k):
        core, factors = tucker
        if hasattr(core, 'detach'):  # torch tensors (TKConv.py:314, TKLinear.py:60): keep autograd
            import torch
            out = torch.tensordot(factors[0], core, dims=([1], [0]))
            out = torch.tensordot(factors[1], out, dims=([1], [1]))
            return torch.movedim(out, 0, 1)
        return port.tucker2_to_tensor(core, factors)

    dec = types.ModuleType('tensorly.decomposition')

    def partial_tucker(tensor, modes, rank=None, n_iter_max=100, init='svd', tol=10e-5, **kw):
        assert list(modes) == [0, 1] and init == 'svd'
        if hasattr(tensor, '__array_interface__') or hasattr(tensor, '__array__') and not hasattr(tensor, 'detach'):
            return port.partial_tucker2(tensor, list(rank), n_iter_max=n_iter_max, tol=tol)
        import torch
        core, factors = port.partial_tucker2(tensor.detach().cpu().numpy(), list(rank),
                                             n_iter_max=n_iter_max, tol=tol)
        return torch.from_numpy(core), [torch.from_numpy(f) for f in factors]

    def parafac(*a, **k):
        raise NotImplementedError('parafac is imported but never called by the reference hot path')

    dec.partial_tucker = partial_tucker
    dec.parafac = parafac
    tl.set_backend = set_backend
    tl.tucker_to_tensor = tucker_to_tensor
    tl.decomposition = dec
    return tl, dec


def _install_stubs():
    if 'tensorly' not in sys.modules:
        tl, dec = _tensorly_stub()
        sys.modules['tensorly'] = tl
        sys.modules['tensorly.decomposition'] = dec
    if 'timm' not in sys.modules:
        timm = types.ModuleType('timm')
        models = types.ModuleType('timm.models')
        registry = types.ModuleType('timm.models.registry')
        registry.register_model = lambda fn: fn
        models.registry = registry
        timm.models = models
        sys.modules.update({'timm': timm, 'timm.models': models, 'timm.models.registry': registry})
    if 'geoopt' not in sys.modules:
        sys.modules['geoopt'] = types.ModuleType('geoopt')


def _load(name, alias_bare=()):
    """Load /root/reference/<name>.py as module `refimpl_<name>`.

    While executing it, bare imports of sibling reference modules (`from ttd import ten2tt`) must
    resolve to the reference's own files, so those are temporarily aliased in sys.modules.
    """
    full = 'refimpl_' + name
    if full in sys.modules:
        return sys.modules[full]
    saved = {}
    for dep in alias_bare:
        saved[dep] = sys.modules.get(dep)
        sys.modules[dep] = _load(dep)
    try:
        spec = importlib.util.spec_from_file_location(full, os.path.join(REFERENCE_ROOT, name + '.py'))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[full] = mod
        spec.loader.exec_module(mod)
    finally:
        for dep, old in saved.items():
            if old is None:
                sys.modules.pop(dep, None)
            else:
                sys.modules[dep] = old
    return mod


class Reference:
    def __init__(self):
        _install_stubs()
        self.ttd = _load('ttd')
        self.admm = _load('admm', alias_bare=('ttd',))
        self.TTConv = _load('TTConv', alias_bare=('ttd',))
        self.TTLinear = _load('TTLinear', alias_bare=('ttd',))
        self.TKConv = _load('TKConv')
        self.TKLinear = _load('TKLinear')
        self.resnet_cifar = _load('resnet_cifar')
        self.SVDConv = _load('SVDConv')
        self.orthogonal = _load('orthogonal')

    def hp_class(self, module, cls):
        full = 'refimpl_hp_' + module
        if full not in sys.modules:
            spec = importlib.util.spec_from_file_location(
                full, os.path.join(REFERENCE_ROOT, 'hp_dicts', module + '.py'))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[full] = mod
            spec.loader.exec_module(mod)
        return getattr(sys.modules[full], cls)


_REF = None


def load_reference():
    global _REF
    if not reference_available():
        raise RuntimeError('reference tree not present at ' + REFERENCE_ROOT)
    if _REF is None:
        _REF = Reference()
    return _REF
