"""BASELINE config 5 at N GPUs: the synthetic Tucker-2 sweep (3 x 3 conv weights, 64 ... 2048 channels, ranks 0.25 / 0.5 x)
as one model whose 12 independent tensors are LPT-sharded over the ranks (the same sharding as the layers of a network,
one all-gather of Z).  Device-timed, max over ranks.

    python scripts/bench_tucker_sweep.py [--channels 64,128,...] [--steps 2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_tucker_sweep.py
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'dnn-compression-tensor-admm_b200')]
import hp_tables  # noqa: E402
import workloads  # noqa: E402
from admm import ADMM  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--channels', default='64,128,256,512,1024,2048')
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--warmup', type=int, default=1)
    args = ap.parse_args()
    channels = tuple(int(v) for v in args.channels.split(','))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = 'cuda:{}'.format(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device(dev))
    model = workloads.ParamBag(workloads.tucker_sweep_weights(0, channels), device=dev)
    admm = ADMM(model, 1e-3, hp_tables.tucker_sweep_all(channels), 'tk', dev)
    for _ in range(max(args.warmup, 1)):
        admm.update()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        admm.update()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    mine = list(admm._shard.local_names)
    names = [None] * world
    if world > 1:
        dist.all_gather_object(names, mine)
    else:
        names = [mine]
    if rank == 0:
        n = len(admm._names)
        print(json.dumps({'metric': 'tucker2_sweep_update', 'n_gpus': world, 'tensors': n, 'channels': list(channels),
                          'ms_per_update': float(ms.item()), 'tensors_per_s': n / (float(ms.item()) / 1e3),
                          'steps': args.steps, 'sharding': names,
                          'hooi_sweeps': {k: v for pl, _ in admm._plans for k, v in getattr(pl, 'hooi_sweeps', {}).items()}}),
              flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
