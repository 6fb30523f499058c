#!/usr/bin/env python
"""Benchmark of the ADMM Z+U update hot path (BASELINE.json metric: "ADMM Z+U update layers/sec
(ResNet-50 TT)").

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

A "step" is one full `ADMM.update()` (Z = Proj_TT(W + U) for all 34 listed ResNet-50 layers, then
U += W - Z) on seeded synthetic random-init weights.  `value` = layers / second with W, U, Z
resident in HBM; `e2e` = the same through the public API with HOST buffers (pinned W copied in, Z
copied out, inside the timed region).  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'dnn-compression-tensor-admm_b200')
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# ncu --set full figures of the dominant kernel are read from a committed summary (profiles/), never a literal
NCU_SUMMARY = os.path.join(ROOT, 'profiles', 'r2_ncu_trd_kernels.json')
FP64_DFMA_PEAK_TFLOPS = 36.8      # measured on this pool's B200s with scripts/ubench/fp64_rate.cu (round 1)

METRIC = 'admm_zu_update_layers_per_sec_resnet50_tt'
UNIT = 'layers/s'
WORKLOAD = 'resnet50_tt_general_3x: full-network ADMM Z+U update, 34 layers, 20,099,072 fp32 weights'


def _peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return {'hbm_gbs': p['hbm_gbs'], 'bf16_tflops': p['bf16_tflops'],
                'bf16_tflops_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling DURING the timed region (B200_PROFILING.md)."""

    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(nm)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's CPU algorithm (oracle port) on the host cores
# ------------------------------------------------------------------------------------------------
def _reference_inputs():
    import hp_tables
    import workloads
    hp = hp_tables.tt_resnet50_general_3x()
    weights = workloads.resnet50_weights(seed=0)
    return hp, {n: weights[n].numpy() for n in hp.ranks}


def reference_step_time(hp, weights, u_state):
    """Seconds for one full-network Z+U update with the reference's algorithm (`oracle.port`, numpy / LAPACK
    gesdd on the host cores): every one of the 34 listed layers, nothing extrapolated."""
    from oracle import port
    t0 = time.perf_counter()
    for name, w in weights.items():
        v = w + u_state[name]
        z = port.project_conv_tt(v, hp.tt_shapes[name], list(hp.ranks[name]))
        u_state[name] += w - z
    return time.perf_counter() - t0


def common_config():
    """Identical in both arms (the driver compares the two `config` objects)."""
    return {'workload': WORKLOAD}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    hp, weights = _reference_inputs()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    u_state = {n: np.zeros_like(w) for n, w in weights.items()}
    for _ in range(args.warmup):
        reference_step_time(hp, weights, u_state)
    times = [reference_step_time(hp, weights, u_state) for _ in range(args.steps)]
    sec = float(np.mean(times))
    value = 34.0 / sec
    desc = 'every step = all 34 layers (one full ADMM.update of the reference algorithm), weights in host memory'
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': common_config(),
            'details': {'inputs': 'host memory; numpy/LAPACK gesdd (oracle port of ttd.py/admm.py)',
                        'ms_per_step_min': float(np.min(times)) * 1e3, 'ms_per_step_max': float(np.max(times)) * 1e3},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': desc},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def golden_check(admm_cls, model_cls, weights, hp, dev):
    """Parity of THIS process's kernels against the records the unmodified reference left in
    tests/golden/reference_summary.json (Frobenius norm + 32-point probe of Z after the 1st and the 3rd update of
    every layer): run on rank 0 at every N, sharded exactly like the timed updates."""
    path = os.path.join(ROOT, 'tests', 'golden', 'reference_summary.json')
    if not os.path.isfile(path):
        return {'ok': None, 'why': 'tests/golden/reference_summary.json missing'}
    with open(path) as f:
        gold = json.load(f)['resnet50_tt']['layers']
    a = admm_cls(model_cls(weights, device=dev), 1e-3, hp, 'tt', dev)
    worst_norm, worst_probe = 0.0, 0.0

    def compare(tag):
        nonlocal worst_norm, worst_probe
        for n in a._names:
            z = a.z[n].detach().cpu().numpy().reshape(-1)
            g = gold[n][tag]
            fro = float(np.sqrt(np.sum(z.astype(np.float64) ** 2)))
            worst_norm = max(worst_norm, abs(fro - g['fro']) / max(g['fro'], 1e-30))
            rng = np.random.RandomState(z.size % 9973)        # probe positions of oracle/gen_golden.py
            probe = z[rng.randint(0, z.size, size=32)].astype(np.float64)
            scale = g['fro'] / np.sqrt(z.size)
            worst_probe = max(worst_probe, float(np.max(np.abs(probe - np.asarray(g['probe'])))) / scale)

    a.update(update_u=False)
    compare('z0')
    a.update()
    a.update()
    compare('z2')
    torch.cuda.synchronize()
    ok = worst_norm <= 2e-5 and worst_probe <= 8e-4
    return {'ok': bool(ok), 'layers': len(a._names), 'worst_rel_norm_diff': worst_norm,
            'worst_probe_diff_over_rms': worst_probe,
            'tolerance': 'norm 2e-5, probe 8e-4 of the rms entry (tests/helpers.check_summary)',
            'against': 'tests/golden/reference_summary.json (unmodified reference, oracle/gen_golden.py)'}


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import hp_tables
    import projector
    import tta_runtime as rt
    import workloads
    from admm import ADMM

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; this path has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    hp = hp_tables.tt_resnet50_general_3x()
    weights = workloads.resnet50_weights(seed=0)
    model = workloads.ParamBag(weights, device=dev)
    admm = ADMM(model, 1e-3, hp, 'tt', dev)
    n_layers = len(admm._names)
    numel = admm._total_numel

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    admm.update(update_u=False)                      # Z_0 = Proj(W), as engines.py:245
    for _ in range(max(args.warmup, 3)):
        admm.update()

    # ---- device-resident timing -----------------------------------------------------------------
    sampler = ClockSampler(local)
    launches0 = rt.launch_count()
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        admm.update()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = rt.launch_count() - launches0
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())

    # ---- the same with the Jacobi warm start disabled (every eigenproblem starts from X = G) --------
    def set_warm(on):
        for plan, _ in admm._plans:
            if hasattr(plan, 'warm_start'):
                plan.warm_start = on
    set_warm(False)
    admm.update()
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    cold_steps = max(3, args.steps // 4)
    for _ in range(cold_steps):
        admm.update()
    c1.record()
    barrier()
    cold_ms = c0.elapsed_time(c1) / cold_steps
    t = torch.tensor([cold_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cold_ms = float(t.item())
    set_warm(True)
    admm.update()

    # ---- end-to-end: host buffers in, host buffers out ------------------------------------------
    host_w = {n: weights[n].contiguous().pin_memory() for n in admm._names}
    host_z = {n: torch.empty_like(host_w[n]).pin_memory() for n in admm._names}
    params = dict(model.named_parameters())

    def e2e_step():
        # public API for host-resident weights: pinned W in, Z out (admm.ADMM.update_from_host); the copies of
        # one layer group run on that group's stream and overlap the projection of the others.  Under sharding
        # every rank uploads / downloads only its own layers (the other layers' W arrive by an NVLink all-gather),
        # so the job as a whole moves 4 * numel bytes each way per step
        admm.update_from_host(host_w, host_z)
        torch.cuda.current_stream().synchronize()

    e2e_step()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        e2e_step()
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1) / args.steps
    for n in admm._shard.local_names:      # every rank downloads the Z of the layers it projected
        if not torch.equal(host_z[n], admm.z[n].cpu()):
            raise SystemExit('bench.py: host Z of {} differs from the device tensor'.format(n))
    if world > 1:
        # the weights of the layers other ranks project must arrive by the NVLink all-gather: wipe them on the device,
        # run one more end-to-end step, compare every parameter with the host copy
        local = set(admm._shard.local_names)
        for n in admm._names:
            if n not in local:
                params[n].data.zero_()
        e2e_step()
        for n in admm._names:
            if not torch.equal(params[n].data.cpu(), host_w[n]):
                raise SystemExit('bench.py: device W of {} differs from the host tensor after update_from_host'.format(n))
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    # the same end-to-end step with every warm start disabled (what the first update of a run, or an update after
    # a full epoch of SGD, costs)
    set_warm(False)
    e2e_step()
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(cold_steps):
        e2e_step()
    g1.record()
    barrier()
    t = torch.tensor([g0.elapsed_time(g1) / cold_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_cold_ms = float(t.item())
    set_warm(True)

    # ---- per-phase profile (separate, untimed-for-the-metric pass) --------------------------------
    phases = {}
    jac_ms, jac_launches = 0.0, 0
    ew_ms = None
    prof_steps = 3
    trd_ms, trd_launches = 0.0, 0
    if rank == 0 or world > 1:
        rt.jacobi_profile(True)
        rt.jacobi_profile_read()
        rt.symeig_profile(True)
        rt.symeig_profile_read()
        for plan, _ in admm._plans:
            plan.profile = phases
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ew_acc = 0.0
        for _ in range(prof_steps):
            with torch.no_grad():
                for plan, names in admm._plans:
                    plan.run([params[n].data for n in names], [admm.u[n] for n in names], [admm.z[n] for n in names])
                admm._shard.exchange(admm.z)
                # 5 back-to-back launches per sample: one launch is ~50 us, comparable to the launch latency of
                # a ctypes call; W, Z, U (321 MB touched per launch) exceed the 126 MB L2
                ew_tab = admm._ew_table()
                rt.dual_update(ew_tab, None)
                d0.record()
                for _ in range(5):
                    rt.dual_update(ew_tab, None)
                d1.record()
                d1.synchronize()
                ew_acc += d0.elapsed_time(d1) / 5
        jac_ms, jac_launches = rt.jacobi_profile_read()
        rt.jacobi_profile(False)
        trd_ms, trd_launches = rt.symeig_profile_read()
        rt.symeig_profile(False)
        trd_ms /= prof_steps
        trd_launches //= prof_steps
        for plan, _ in admm._plans:
            plan.profile = None
        phases = {k: v / prof_steps for k, v in phases.items()}
        ew_ms = ew_acc / prof_steps
        jac_ms /= prof_steps
        jac_launches //= prof_steps

    # ---- parity of this very process (all ranks take part: the update is a collective under sharding) ----
    parity = None
    if not args.no_check:
        parity = golden_check(ADMM, workloads.ParamBag, weights, hp_tables.tt_resnet50_general_3x(), dev)
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = _peaks()
    # algorithmic work of the local share (SURVEY 8(d) accounting)
    local_names = admm._shard.local_names
    fl = np.zeros(4)
    for n in local_names:
        fl += np.array(projector.tt_step_flops(hp.tt_shapes[n], hp.ranks[n]), dtype=np.float64)
    # ---- roofline of the dominant kernel: trd_reduce_kernel (Householder tridiagonalisation of the Gram matrices
    # with 32 < k <= 608, one thread-block cluster per problem; csrc/trd.cu) -----------------------------------
    trd_ks = []
    for n in local_names:
        shp, rk = hp.tt_shapes[n], projector.clip_tt_ranks(hp.tt_shapes[n], hp.ranks[n])
        for i in range(len(shp) - 1):
            k = min(rk[i] * shp[i], projector._prod(shp[i + 1:]))
            if k != rk[i + 1] and projector.uses_trd(k, projector.default_solver()):
                trd_ks.append(k)
    n_l = max(trd_launches, 1)
    per_launch_s = (trd_ms / 1e3) / n_l
    eig_flop = float(sum(9.0 * k ** 3 for k in trd_ks))            # SURVEY 8(d) accounting of an eigensolve
    alg_bytes = float(sum(8.0 * (k * k + k * ((k + 31) // 32 * 32) + 3 * k) for k in trd_ks))   # read G, write V, d, e, tau
    dfma = float(sum(k ** 3 for k in trd_ks))                      # 3 DFMA per column element and step: sum_j 3 (k-j)^2
    ncu = {}
    if os.path.isfile(NCU_SUMMARY):
        with open(NCU_SUMMARY) as f:
            ncu = json.load(f)
    roofline = {'kernel': 'trd_reduce_kernel', 'bound': 'hbm',
                'achieved': (alg_bytes / n_l) / per_launch_s / 1e9 if trd_launches else 0.0,
                'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                'frac': (alg_bytes / n_l) / per_launch_s / 1e9 / peaks['hbm_gbs'] if trd_launches else 0.0,
                'traffic': ncu.get('trd_reduce_kernel', {}).get('dram_bytes_per_launch'),
                'traffic_source': ncu.get('trd_reduce_kernel', {}).get('source'),
                'peak_source': peaks['source'] + ' (copy bandwidth)',
                'launches_per_step': trd_launches, 'avg_launch_us': per_launch_s * 1e6,
                'problems_per_step': len(trd_ks),
                'algorithmic_bytes_per_launch': alg_bytes / n_l,
                'eig_flops_9k3_per_launch': eig_flop / n_l,
                'eig_tflops_9k3': (eig_flop / n_l) / per_launch_s / 1e12 if trd_launches else 0.0,
                'fp64_executed_tflops': 2.0 * dfma / (trd_ms / 1e3) / 1e12 if trd_ms else None,
                'fp64_dfma_peak_tflops': FP64_DFMA_PEAK_TFLOPS,
                'note': 'the kernel is neither HBM- nor tensor-bound: the matrix lives in the shared memory of a 16-CTA '
                        'cluster (DRAM traffic = one read of G) and the k - 2 Householder steps are strictly sequential, '
                        'so it is bound by the per-step latency chain (fp64 CUDA-core pipe, DSMEM exchange); the HBM '
                        'figure is reported because the contract asks for one of hbm | tensor, the fp64 figures show '
                        'what the kernel really executes (launches of different layer groups overlap, so the summed '
                        'launch time exceeds the step time)'}
    ew_bytes = 16.0 * numel
    kernels = {
        'note': 'per-phase device times of a separate profiling pass in which the layer groups run one after another '
                '(CUDA events on one stream); in the timed update the groups overlap on their own streams',
        'dual_update': {'bound': 'hbm', 'ms': ew_ms, 'achieved': ew_bytes / (ew_ms / 1e3) / 1e9 if ew_ms else None,
                        'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                        'frac': (ew_bytes / (ew_ms / 1e3) / 1e9) / peaks['hbm_gbs'] if ew_ms else None},
        'gram_tcgen05_3xtf32': {'ms': phases.get('gram'),
                                'tflops_fp32_equivalent': fl[0] / (phases['gram'] / 1e3) / 1e12 if phases.get('gram') else None,
                                'note': 'gram_tc_kernel (TMA + tcgen05, accumulator flushed every 32 indices) + gram_finish; '
                                        'round 1: fp64 DFMA kernels, 2.0 ms'},
        'eig_phase_incl_host_sync': {'ms': phases.get('eig')},
        'project_gemm': {'ms': phases.get('project'), 'tflops': fl[1] / (phases['project'] / 1e3) / 1e12 if phases.get('project') else None},
        'recon_gemm': {'ms': phases.get('recon'), 'tflops': fl[3] / (phases['recon'] / 1e3) / 1e12 if phases.get('recon') else None},
        'unfold': {'ms': phases.get('unfold')}, 'fold': {'ms': phases.get('fold')}, 'select': {'ms': phases.get('select')},
    }

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        hp_r, weights_r = _reference_inputs()
        u_state = {n: np.zeros_like(w) for n, w in weights_r.items()}
        secs = [reference_step_time(hp_r, weights_r, u_state) for _ in range(3)]
        cpu_baseline = {'value': 34.0 / min(secs), 'unit': UNIT, 'cores': os.cpu_count() or 1, 'kind': 'port',
                        'sample': 'three full 34-layer updates of the reference algorithm (oracle port, numpy / LAPACK '
                                  'on the host cores), best of 3; nothing extrapolated'}


    forward = None
    if world == 1 and not args.no_forward:
        forward = forward_bench(dev, peaks)

    sweeps = [s for v in admm.sweeps.values() for s in v]
    line = {'metric': METRIC, 'value': n_layers / (ms / 1e3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': common_config(),
            'details': {'sharding': 'layers LPT-sharded over {} rank(s) + 1 all-gather of Z'.format(world),
                        'l2': 'working set W+U+Z = {:.0f} MB plus workspaces exceeds the 126 MB L2; no explicit flush'
                        .format(3 * 4 * numel / 1e6),
                        'eigensolver': 'k <= 32: one-CTA Jacobi + fp64 refinement (warm-started from the previous update); '
                                       '32 < k <= 608: fp64 tridiagonalisation route (csrc/trd.cu), no warm start',
                        'jacobi_sweeps_max': int(max(sweeps)) if sweeps else None,
                        'cold_ms_per_step': cold_ms,
                        'cold': 'value / e2e are successive ADMM updates (U changes by W - Z each step) with the warm start '
                                'of the k <= 32 Jacobi problems; cold_* = the same with every warm start disabled'},
            'clocks': clocks,
            'e2e': {'value': n_layers / (e2e_ms / 1e3), 'unit': UNIT, 'ms_per_step': e2e_ms,
                    'h2d_bytes_per_step': 4 * numel, 'd2h_bytes_per_step': 4 * numel},
            'e2e_cold': {'value': n_layers / (e2e_cold_ms / 1e3), 'unit': UNIT, 'ms_per_step': e2e_cold_ms},
            'value_cold': n_layers / (cold_ms / 1e3),
            'parity': parity,
            'gpu_launches': int(launches), 'roofline': roofline, 'kernels': kernels, 'cpu_baseline': cpu_baseline,
            'forward': forward}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _time_cuda(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / iters


def _time_graph(fn, reps=20, iters=5):
    """Device time of one call of `fn` with the host out of the way: `reps` calls captured in a CUDA graph, the graph
    replayed `iters` times (a 15 us kernel launched from Python is bound by the ~20 us host cost of the call)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(reps):
            fn()
    return _time_cuda(graph.replay, iters=iters, warm=2) / reps


def forward_bench(dev, peaks):
    """Secondary metric of BASELINE.json ("TT-conv img/s"): decomposed-layer forwards, TT layers only.

    fused  = this repo's inference path (libtta.so kernels, bf16);
    chain  = the same contraction as an unfused torch op chain in fp32/TF32 on the same GPU (the
             reference's algorithm restated; the reference modules themselves cannot travel).
    """
    import hp_tables
    import tta_runtime as rt
    import TTConv
    import TTLinear
    out = {}
    torch.manual_seed(0)
    # ---- ttm_resnet32 TT layers, batch 128, 32x32 (config 2) ----
    hp = hp_tables.tt_resnet32_3x()
    planes = {1: 16, 2: 32, 3: 64}
    size = {1: 32, 2: 16, 3: 8}
    layers = []
    for name in hp.ranks:
        s, b, c = int(name[5]), int(name[7]), int(name[13])
        cout = planes[s]
        first = (b == 0 and c == 1 and s > 1)
        cin = planes[s - 1] if first else cout
        stride = 2 if first else 1
        hw = size[s - 1] if first else size[s]
        layer = TTConv.TTConv2dM(cin, cout, 3, stride=stride, padding=1, bias=False, hp_dict=hp, name=name).to(dev)
        layers.append((layer, torch.randn(128, cin, hw, hw, device=dev)))

    def run_fused():
        with torch.no_grad():
            for layer, x in layers:
                layer(x)

    def run_chain():
        with torch.no_grad():
            for layer, x in layers:
                layer._forward_torch(x)

    import fwd_common as fc

    def run_chain_bf16():
        with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
            for layer, x in layers:
                layer._forward_torch(x)

    ms_f, ms_c, ms_cb = _time_cuda(run_fused), _time_cuda(run_chain), _time_cuda(run_chain_bf16)
    # the same 30 launches replayed from a CUDA graph: the per-layer host cost (Python + ctypes, ~10 us) is gone
    ms_g = None
    try:
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run_fused()
        torch.cuda.current_stream().wait_stream(side)
        with torch.cuda.graph(graph):
            run_fused()
        ms_g = _time_cuda(graph.replay)
    except Exception as exc:      # capture is an optimisation of the measurement, not part of the path
        ms_g = None
        sys.stderr.write('ttm_resnet32 graph capture skipped: {}\n'.format(exc))
    # fp32 CUDA-core fused kernel (round 1) for comparison
    fc._TTCONV_TC = False
    for layer, _ in layers:
        if getattr(layer, '_folded', None) is not None:
            layer._folded.tc_ok.clear()
    ms_cc = _time_cuda(run_fused)
    fc._TTCONV_TC = True
    for layer, _ in layers:
        if getattr(layer, '_folded', None) is not None:
            layer._folded.tc_ok.clear()
    # HBM floor: every layer reads its input and writes its output activations once (fp32 NCHW, as the reference's modules)
    act_bytes = 0
    n_tc = 0
    for layer, x in layers:
        s_ = layer.stride[0] if isinstance(layer.stride, (tuple, list)) else layer.stride
        ho = (x.shape[2] + 2 - 3) // s_ + 1
        act_bytes += 4 * (x.numel() + x.shape[0] * layer.out_channels * ho * ho)
        n_tc += int(s_ == 1)
    best = min(v for v in (ms_f, ms_g) if v is not None)
    out['ttm_resnet32_tt_layers'] = {'batch': 128, 'fused_img_s': 128 / (ms_f / 1e3), 'fused_ms': ms_f,
                                     'fused_cuda_graph_ms': ms_g,
                                     'fused_fp32_cuda_core_kernel_ms': ms_cc,
                                     'torch_op_chain_img_s': 128 / (ms_c / 1e3), 'torch_op_chain_ms': ms_c,
                                     'torch_op_chain_bf16_autocast_ms': ms_cb,
                                     'kernel': 'ttconv_tc_kernel (bf16 tcgen05, csrc/ttconv_tc.cu) for all {} layers ({} of them '
                                               'stride 2: stride-1 result at every position, odd positions dropped)'
                                               .format(len(layers), len(layers) - n_tc),
                                     'roofline': {'bound': 'hbm', 'algorithmic_bytes': act_bytes,
                                                  'achieved': act_bytes / (best / 1e3) / 1e9, 'peak': peaks['hbm_gbs'],
                                                  'unit': 'GB/s', 'frac': act_bytes / (best / 1e3) / 1e9 / peaks['hbm_gbs'],
                                                  'note': 'input + output activations of the 30 layers, fp32 NCHW, over the '
                                                          'best of the eager and the graph-replayed time'}}
    # the kernel alone on the three stage shapes of ttm_resnet32 (batch 128)
    shp = {}
    for (cin, hw_, ra, rb, cout) in ((16, 32, 16, 16, 16), (32, 16, 32, 32, 32), (64, 8, 64, 64, 64)):
        x = torch.randn(128, cin, hw_, hw_, device=dev)
        a_in = torch.randn(ra, cin, device=dev)
        kern = torch.randn(rb, ra, 3, 3, device=dev)
        a_out = torch.randn(cout, rb, device=dev)
        y = torch.empty(128, cout, hw_, hw_, device=dev)
        res = {}
        blob = rt.ttconv_tc_pack(a_in, kern, a_out, None)
        for nm, fn in (('tc', lambda: rt.ttconv_tc_fwd(x, blob, y, 128, cin, hw_, hw_, ra, rb, cout, 3, 1, 1)),
                       ('cuda_core', lambda: rt.ttconv_fused_fwd(x, a_in, kern, a_out, None, y, 128, cin, hw_, hw_, ra, rb, cout,
                                                                  3, 1, 1))):
            ms = _time_graph(fn)
            by = 4.0 * (x.numel() + y.numel())
            res[nm] = {'us': ms * 1e3, 'algorithmic_hbm_gbs': by / (ms / 1e3) / 1e9,
                       'frac_of_hbm_peak': by / (ms / 1e3) / 1e9 / peaks['hbm_gbs']}
        shp['128x{}x{}x{}_r{}_{}'.format(cin, hw_, hw_, ra, rb)] = res
    out['ttconv_tc_kernel'] = {'bound': 'hbm', 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                               'io': 'fp32 NCHW in and out; algorithmic bytes = x + y; device time per launch from a CUDA '
                                     'graph of 20 launches (the Python call costs more than the kernel)', 'shapes': shp}
    del layers
    # ---- DeiT-small TTLinearM layers, batch 256 x 197 tokens (config 4) ----
    hp = hp_tables.tt_deit_small_2x()
    d = 384
    dims = {'attn.qkv.weight': (d, 3 * d), 'attn.proj.weight': (d, d), 'mlp.fc1.weight': (d, 4 * d), 'mlp.fc2.weight': (4 * d, d)}
    tokens = 256 * 197
    xs = {d: torch.randn(tokens, d, device=dev), 4 * d: torch.randn(tokens, 4 * d, device=dev)}
    lin = []
    for name in hp.ranks:
        fin, fout = dims[name.split('.', 2)[2]]
        lin.append((TTLinear.TTLinearM(fin, fout, bias=True, hp_dict=hp, name=name).to(dev), xs[fin]))

    def lin_fused():
        with torch.no_grad():
            for layer, x in lin:
                layer(x)

    def lin_chain():
        with torch.no_grad():
            for layer, x in lin:
                import fwd_common as fc
                fc.tt_apply_torch(x, list(layer.tt_cores)[layer.out_tt_order:], list(layer.tt_cores)[:layer.out_tt_order])

    xs_bf = {k: v.to(torch.bfloat16) for k, v in xs.items()}

    def lin_fused_bf16():
        with torch.no_grad():
            for layer, x in lin:
                layer(xs_bf[x.shape[1]])

    def lin_chain_bf16():
        with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
            for layer, x in lin:
                import fwd_common as fc
                fc.tt_apply_torch(x, list(layer.tt_cores)[layer.out_tt_order:], list(layer.tt_cores)[:layer.out_tt_order])

    ms_f, ms_c = _time_cuda(lin_fused, iters=3, warm=1), _time_cuda(lin_chain, iters=3, warm=1)
    ms_fb = _time_cuda(lin_fused_bf16, iters=3, warm=1)
    ms_cb = _time_cuda(lin_chain_bf16, iters=3, warm=1)
    # training step of the same layers (forward + backward with a random upstream gradient): fused path
    # (fwd_common.LowRank2Fn: forward and dX on the two-factor kernel) vs the torch op chain
    gys = {}

    def lin_train(fused):
        for layer, x in lin:
            layer.fused_training = fused
            xg = x.detach().requires_grad_(True)
            y = layer(xg)
            gy = gys.get(y.shape[1])
            if gy is None:
                gy = gys[y.shape[1]] = torch.randn_like(y)
            y.backward(gy)
        for layer, _ in lin:
            layer.zero_grad(set_to_none=True)
            layer.fused_training = False

    ms_tf = _time_cuda(lambda: lin_train(True), iters=5, warm=2)
    ms_tc = _time_cuda(lambda: lin_train(False), iters=2, warm=1)
    gys.clear()
    macs = 0
    macs_exec = 0
    n_dense = 0
    for layer, _ in lin:
        r, sh, o = layer.tt_ranks, layer.tt_shapes, layer.out_tt_order
        # per-token MACs of the 4-step chain (SURVEY 8(d) TTLinearM accounting)
        chain = sh[2] * sh[3] * r[3] + sh[2] * r[3] * r[2] + r[2] * sh[1] * r[1] + sh[1] * r[1] * sh[0]
        macs += chain
        dense = bool(getattr(layer, '_dense_first', False))
        n_dense += dense
        macs_exec += layer.in_features * layer.out_features if dense else chain
    out['deit_small_ttlinear_layers'] = {'batch': 256, 'tokens': tokens, 'fused_img_s': 256 / (ms_f / 1e3), 'fused_ms': ms_f,
                                         'torch_op_chain_img_s': 256 / (ms_c / 1e3), 'torch_op_chain_ms': ms_c,
                                         'torch_op_chain_bf16_autocast_ms': ms_cb,
                                         'fused_bf16_activations_ms': ms_fb,
                                         'fused_bf16_activations_img_s': 256 / (ms_fb / 1e3),
                                         'train_step_fused_ms': ms_tf, 'train_step_torch_op_chain_ms': ms_tc,
                                         'two_factor_fused_layers': sum(bool(getattr(l, '_fused2', False)) for l, _ in lin),
                                         'contraction_order': '{} of {} layers fold the cores into the dense weight first '
                                                              '(fewer or comparable MACs than the chain)'.format(n_dense, len(lin)),
                                         'chain_tflops_equiv': 2.0 * macs * tokens / (ms_f / 1e3) / 1e12,
                                         'executed_tflops': 2.0 * macs_exec * tokens / (ms_f / 1e3) / 1e12}
    # ---- the tcgen05 GEMM alone: the two big contractions of the block-0 qkv chain ----
    tc = {}
    for (M, N, K) in ((tokens, 1120, 320), (tokens, 320, 368)):
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        b = torch.randn(N, K, device=dev).to(torch.bfloat16)
        c = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        ms = _time_cuda(lambda: rt.gemm_bf16_tc(a, b, c, M, N, K), iters=10, warm=3)
        tf = 2.0 * M * N * K / (ms / 1e3) / 1e12
        tc['{}x{}x{}'.format(M, N, K)] = {'ms': ms, 'tflops': tf, 'frac_of_bf16_peak': tf / peaks['bf16_tflops']}
    out['gemm_bf16_tc_kernel'] = {'bound': 'tensor', 'peak': peaks['bf16_tflops'], 'unit': 'TFLOP/s', 'shapes': tc}
    # ---- the fused two-factor kernel (TMA + tcgen05, intermediate on chip) on the DeiT-small layer shapes ----
    lr = {}
    for nm, K1, N1, N2 in (('qkv_block0', 384, 320, 1152), ('qkv', 384, 256, 1152), ('proj', 384, 256, 384),
                           ('fc1', 384, 256, 1536), ('fc2', 1536, 256, 384)):
        x = torch.randn(tokens, K1, device=dev).to(torch.bfloat16)
        w1 = torch.randn(N1, K1, device=dev).to(torch.bfloat16)
        w2 = torch.randn(N2, N1, device=dev).to(torch.bfloat16)
        y = torch.empty(tokens, N2, device=dev, dtype=torch.bfloat16)
        ms = _time_cuda(lambda: rt.lowrank2_fwd(x, w1, w2, None, y, tokens, K1, N1, N2), iters=10, warm=3)
        tf = 2.0 * tokens * N1 * (K1 + N2) / (ms / 1e3) / 1e12
        gbs = tokens * 2.0 * (K1 + N2) / (ms / 1e3) / 1e9
        lr['{}_{}x{}x{}x{}'.format(nm, tokens, K1, N1, N2)] = {
            'ms': ms, 'tflops': tf, 'frac_of_bf16_peak': tf / peaks['bf16_tflops'],
            'algorithmic_hbm_gbs': gbs, 'frac_of_hbm_peak': gbs / peaks['hbm_gbs']}
    out['lowrank2_fwd_kernel'] = {'bound': 'tensor', 'peak': peaks['bf16_tflops'], 'unit': 'TFLOP/s',
                                  'io': 'bf16 in, bf16 out; algorithmic bytes = x + y (weights stay in L2)', 'shapes': lr}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-forward', action='store_true')
    ap.add_argument('--no-check', action='store_true', help='skip the golden-record parity check of the projected Z')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
