#!/usr/bin/env python
"""bench.py -- site-updates/s of batched Villain Metropolis sweeps (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (N=1): BASELINE config 2 -- Villain (phi, n), L=32, kappa=0.5, 4096 independent chains on
one B200, one checkerboard NeighborhoodUpdate sweep per step with the action / winding / wrapping
reductions fused in.  With N GPUs every rank runs that workload on its own 4096 chains (weak
scaling; chains are independent, there is no collective on the hot path; a final all_gather of
the observable records runs outside the timed region).

A step is ONE kernel launch.  The chain state (64 MiB) would fit the 126 MB L2, so the timed loop
rotates over 4 independent chain sets (256 MiB): every step streams its set from HBM.  Steps are
issued as overlapped launches (svb_villain_sweep_overlapped: programmatic dependent launch with
per-chain epochs ordering the data), the way BatchedEnsemble.generate issues them, so the ramp-up
of one launch hides under the tail of the one before; `--no-overlap` times ordinary launches.

One JSON line on stdout; see the task contract for the keys.  `--impl reference` times the CPU
restatement of the reference's numpy algorithm (oracle/villain_np.py) on all host cores.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'site-updates/sec (batched Villain Metropolis sweeps)'
UNIT = 'site-updates/s'
L, KAPPA, W_CONSTRAINT, CHAINS = 32, 0.5, 1, 4096
BYTES_PER_SITE_UPDATE = 32          # fp64 phi + 2 x int32 n, one read + one write
ROTATE = 4                           # chain sets rotated through so every step comes from HBM, not L2
THERMALISE = 200                     # untimed sweeps applied to each synthetic hot start before the warm-up
WORKLOAD = 'config2: Villain (phi,n) L=32 kappa=0.5 W=1, 4096 chains/GPU, NeighborhoodUpdate checkerboard sweep + action/winding/wrapping'


def traffic_from_profile(override):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the sweep kernel, from the committed ncu capture."""
    if override is not None:
        return override
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)['dram_bytes_per_launch'])
    return None


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md, 6.65 TB/s)'


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle's port of the reference's numpy NeighborhoodUpdate, one chain per process
# ---------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, sweeps = args
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    from oracle import villain_np as V
    rng = np.random.default_rng(seed)
    phi, n = V.hot_start(np.random.default_rng(1000 + seed), L)
    phi, n = V.neighborhood_step(phi, n, KAPPA, W_CONSTRAINT, rng)       # untimed first sweep (imports, caches)
    t0 = time.perf_counter()
    for _ in range(sweeps):
        phi, n = V.neighborhood_step(phi, n, KAPPA, W_CONSTRAINT, rng)
    return time.perf_counter() - t0


def cpu_reference_rate(procs, seconds):
    """site-updates/s of `procs` processes each sweeping its own L=32 chain for about `seconds` of wall clock.

    Returns (rate, wall, sweeps_per_process).  The sweep count is calibrated on a short untimed run so
    the timed sample is bounded."""
    ctx = mp.get_context('spawn')
    with ctx.Pool(procs) as pool:
        pool.map(_cpu_worker, [(i, 1) for i in range(procs)])            # warm the pool (imports)
        probe = max(pool.map(_cpu_worker, [(i, 20) for i in range(procs)])) / 20
        sweeps = max(10, int(seconds / max(probe, 1e-6)))
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(i, sweeps) for i in range(procs)])
        wall = time.perf_counter() - t0
    return procs * sweeps * L * L / wall, wall, sweeps


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    cores = host_cores()
    per_step = []
    # bounded: the whole --steps/--warmup run stays within a few minutes whatever K and W are
    steps, warmup = min(args.steps, 20), min(args.warmup, 3)
    seconds = min(3.0, 60.0 / max(steps, 1))
    for _ in range(warmup):
        cpu_reference_rate(cores, 0.2)
    t_all = time.perf_counter()
    sweeps = 0
    for _ in range(steps):
        rate, wall, sweeps = cpu_reference_rate(cores, seconds)
        per_step.append((rate, wall))
    total_wall = time.perf_counter() - t_all
    value = float(np.mean([r for r, _ in per_step]))
    sample = (f'{steps} timed steps; each step = {cores} processes x ~{sweeps} sweeps of one L=32 kappa=0.5 chain each '
              f'(~{seconds:.1f} s; numpy port of neighborhood.py:59-137)')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total_wall / max(args.steps, 1),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'L': L, 'kappa': KAPPA, 'chains_per_gpu': CHAINS},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons every 100 ms from ONE long-lived `nvidia-smi -lms` process (started before the
    warm-up: forking inside the timed region costs the launching thread milliseconds)."""
    QUERY = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []               # (timestamp, fields)
        self.stop_flag = threading.Event()
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.QUERY}',
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                parts = [p.strip() for p in line.strip().split(',')]
                if len(parts) >= 6:
                    self.samples.append((time.time(), parts))
                if self.stop_flag.is_set():
                    break
        except Exception:
            pass

    def stop(self):
        self.stop_flag.set()
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        self.join(timeout=3)

    def summary(self, t_from=None, t_to=None):
        rows = [f for (t, f) in self.samples if (t_from is None or t >= t_from) and (t_to is None or t <= t_to)]
        if not rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unsampled']}
        sm = [float(r[0]) for r in rows if r[0].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith('active') for r in rows)]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': float(rows[0][1]),
                'reasons': reasons, 'samples': len(rows)}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    import supervillain_b200 as svb
    from supervillain_b200._lib import VOBS_COUNT
    from supervillain_b200.generator.villain import NeighborhoodUpdate
    from supervillain_b200.hostpath import HostStepper
    from supervillain_b200 import sharding

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)

    S = svb.Villain(svb.Lattice2D(L), KAPPA, W=W_CONSTRAINT)
    G = NeighborhoodUpdate(S, seed=20260101)
    # weak scaling: 4096 chains per GPU; global chain ids make the draws independent of the GPU count
    chain0, n_chains = sharding.shard_chains(world * CHAINS, world, rank)
    assert n_chains == CHAINS
    sets = []
    for r in range(ROTATE):
        E = svb.BatchedEnsemble(S, CHAINS, chain0=chain0)
        sets.append(E._start('hot', 20260101 + 7919 * (rank * ROTATE + r)))
    obs_sets = [torch.zeros((CHAINS, VOBS_COUNT), dtype=torch.float64, device=dev) for _ in sets]   # one record per set
    for phi, n in sets:                                       # untimed thermalisation of the synthetic hot starts
        G.sweep_device(phi, n, THERMALISE, chain0=chain0)

    if args.no_overlap:
        plans = [G.plan_device(phi, n, obs=o, chain0=chain0) for (phi, n), o in zip(sets, obs_sets)]

        def step(k):
            plans[k % ROTATE](args.sweeps_per_step)
    else:
        # the way BatchedEnsemble.generate steps: launch k writes its counters to record k and completes the state columns
        # of record k - 1 (the observables of the chains as they arrive ride along with the pass that builds the residuals)
        steppers = [G.overlapped_device(phi, n, chain0=chain0) for phi, n in sets]
        prev_sets = [torch.zeros_like(o) for o in obs_sets]

        def step(k):
            r = k % ROTATE
            steppers[r](args.sweeps_per_step, obs_sets[r], prev_sets[r])
            obs_sets[r], prev_sets[r] = prev_sets[r], obs_sets[r]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()                                           # one nvidia-smi process for the whole run, forked before the warm-up
    for k in range(args.warmup):
        step(k)
    barrier()
    t_load0 = time.time()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for k in range(args.steps):
        step(args.warmup + k)
    stop.record()
    barrier()
    ms = start.elapsed_time(stop)
    if ms < 1000:             # a short timed region sees few 100 ms samples: keep the same load on until a second has passed
        k = 0
        while time.time() < t_load0 + 1.0:
            step(k); k += 1
        torch.cuda.synchronize()
    t_load1 = time.time()
    sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    updates_per_step = CHAINS * L * L * args.sweeps_per_step
    value = world * updates_per_step * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (the step IS one launch of villain_smem_kernel) ----
    peak, peak_src = measured_peaks()
    launch_s = (ms * 1e-3) / args.steps
    achieved = BYTES_PER_SITE_UPDATE * CHAINS * L * L * args.sweeps_per_step / launch_s / 1e9
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic_from_profile(args.traffic), 'kernel': 'villain_smem_filtered_kernel<N=32, 128 threads>',
                'launches': 'ordinary' if args.no_overlap else 'overlapped (programmatic dependent launch + per-chain epochs)',
                'algorithmic_bytes_per_launch': BYTES_PER_SITE_UPDATE * CHAINS * L * L,
                'peak_source': peak_src, 'sweeps_per_launch': args.sweeps_per_step}

    # ---- end to end through the host-buffer API: H2D of the fields, sweep, D2H of fields + observables ----
    stepper = HostStepper(G, CHAINS, chain0=chain0)
    host_sets = [stepper.pinned_fields(from_device=s) for s in sets[:2]]
    for k in range(2):
        stepper.step(*host_sets[k % 2], n_sweeps=args.sweeps_per_step)
    barrier()
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    t0 = time.perf_counter()
    pending = None
    for k in range(e2e_steps):                               # two steps in flight, on alternating host buffer pairs
        issued = stepper.step_async(*host_sets[k % 2], n_sweeps=args.sweeps_per_step)
        if pending is not None:
            pending.wait()
        pending = issued
    pending.wait()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * updates_per_step * e2e_steps / float(t.item())
    e2e = {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': stepper.h2d_bytes, 'd2h_bytes_per_step': stepper.d2h_bytes,
           'steps': e2e_steps, 'api': 'HostStepper.step_async(phi_host, n_host): pinned host fields in and out + observables, two steps in flight on '
                  'alternating host buffer pairs'}

    # ---- final gather of observables (outside the timed region; the only inter-GPU traffic) ----
    torch.cuda.synchronize()
    obs = svb.ops.villain_observables(sets[0][0], sets[0][1], KAPPA)      # final state of set 0
    all_obs = sharding.gather_columns(obs)                    # (world * CHAINS, VOBS_COUNT) on every rank
    mean_action_density = float(all_obs[:, 0].mean().item()) / (L * L)

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = host_cores()
            rate, wall, sweeps = cpu_reference_rate(cores, 12.0)
            cpu = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                   'sample': f'{cores} processes x {sweeps} sweeps of one L=32 kappa=0.5 chain each, {wall:.1f} s wall '
                             f'(oracle/villain_np.py port of neighborhood.py:59-137)'}
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'L': L, 'kappa': KAPPA, 'W': W_CONSTRAINT, 'chains_per_gpu': CHAINS,
                       'sweeps_per_step': args.sweeps_per_step, 'start': f'hot (phi~U(-pi,pi), n~integers(-2,3)) + {THERMALISE} untimed sweeps',
                       'rng': 'philox4x32-10 in-kernel',
                       'l2': f'inputs larger than L2: {ROTATE} chain sets ({ROTATE * CHAINS * L * L * 16 >> 20} MiB) rotated',
                       'parallelism': f'chains sharded over {world} GPU(s), no hot-path collective'},
            'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': args.steps,
            'clocks': sampler.summary(t_load0, t_load1), 'check': {'mean_action_density': mean_action_density, 'gathered_chains': int(all_obs.shape[0])},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2000)
    ap.add_argument('--warmup', type=int, default=50)
    ap.add_argument('--impl', default='svb200', choices=['svb200', 'reference'])
    ap.add_argument('--sweeps-per-step', type=int, default=1)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--e2e-steps', type=int, default=20, help='timed steps of the host-buffer (e2e) leg')
    ap.add_argument('--no-overlap', action='store_true', help='ordinary launches instead of overlapped ones')
    ap.add_argument('--traffic', type=float, default=None,
                    help='dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun when asked for several GPUs directly
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={args.gpus}',
               '--master-addr', '127.0.0.1', '--master-port', '29517', os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_gpu(args)


if __name__ == '__main__':
    sys.exit(main())
