"""world_size-2 `gloo` test of the layer-sharded Z-update: each rank projects only its
LPT share, one all-gather makes every Z resident everywhere, U stays bit-identical across ranks.
Runs on CPU through the C-ABI emulator."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, key, out_dir):
    for p in (ROOT, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200'), os.path.join(ROOT, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import tta_runtime as rt
    import workloads
    from admm import ADMM
    from fake_tta import FakeTTA
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    fake = FakeTTA()
    rt.set_backend_for_tests(fake)
    wb, hb, fmt = workloads.CONFIGS[key]
    weights = wb()
    names = list(weights)[:12]
    weights = {n: weights[n] for n in names}
    a = ADMM(workloads.ParamBag(weights), 1e-3, hb(), fmt, 'cpu')
    a.update(update_u=False)
    a.update()
    # the last update goes through the host-buffer entry point: the parameters are zeroed first, so the result can
    # only be right if update_from_host uploads the local layers' weights AND the other ranks' weights arrive with
    # the weight all-gather; every rank downloads only the Z of its own layers
    host_z = {n: torch.full_like(w, float('nan')) for n, w in weights.items()}
    for n, prm in a.model.named_parameters():
        prm.data.zero_()
    a.update_from_host(weights, host_z)
    local = list(a._shard.local_names)
    for n, prm in a.model.named_parameters():
        assert torch.equal(prm.data, weights[n]), n          # W of every layer is resident on every rank
    for n in names:
        if n in local:
            assert torch.equal(host_z[n], a.z[n]), n
        else:
            assert bool(torch.isnan(host_z[n]).all()), n     # not this rank's to download
    full_z = {n: torch.empty_like(w) for n, w in weights.items()}
    a.update_from_host(weights, full_z, update_u=False, gather_host_z=True)
    for n in names:
        assert torch.equal(full_z[n], a.z[n]), n
    np.savez(os.path.join(out_dir, 'rank{}.npz'.format(rank)), local=np.array(local),
             n_proj=np.array(fake.calls.count('jacobi')),
             **{'z|' + n: a.z[n].numpy() for n in names}, **{'u|' + n: a.u[n].numpy() for n in names})
    dist.destroy_process_group()


@pytest.mark.parametrize('key', ['resnet32_tt', 'resnet32_tk'])
def test_sharded_update_matches_single_process(tmp_path, key):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, key, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(os.path.join(str(tmp_path), 'rank0.npz'))
    r1 = np.load(os.path.join(str(tmp_path), 'rank1.npz'))
    # disjoint cover of the layers
    l0, l1 = set(r0['local'].tolist()), set(r1['local'].tolist())
    assert l0 and l1 and not (l0 & l1)
    sys.path.insert(0, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200'))
    import workloads
    from oracle import port as oracle
    wb, hb, fmt = workloads.CONFIGS[key]
    weights = wb()
    names = list(weights)[:12]
    assert l0 | l1 == set(names)
    o = oracle.OracleADMM({n: weights[n].numpy() for n in names}, 1e-3, hb(), fmt)
    o.update(update_u=False)
    o.update()
    o.update()
    o.update(update_u=False)            # the workers' fourth, Z-only update (gather_host_z variant)
    for n in names:
        assert np.array_equal(r0['z|' + n], r1['z|' + n]), n          # Z identical on every rank
        assert np.array_equal(r0['u|' + n], r1['u|' + n]), n          # U bit-identical across ranks
        err = np.linalg.norm(r0['z|' + n] - o.z[n]) / np.linalg.norm(o.z[n])
        assert err <= 2e-5, (n, err)
