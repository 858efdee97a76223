#!/usr/bin/env python
"""bench.py -- site-updates/s of batched Villain Metropolis sweeps (BASELINE.json metric), every named shape.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--configs c2,c3,c4,c5]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Headline workload (the top-level keys of the JSON line): BASELINE config 2 -- Villain (phi, n), L=32,
kappa=0.5, 4096 independent chains per B200, one checkerboard NeighborhoodUpdate sweep per step with the
action / winding / wrapping reductions fused in.  The `configs` array of the same line carries the other
named shapes at their per-GPU shard sizes -- c3 (worldline L=64, 1024 chains/GPU = 8192 over 8 GPUs),
c4 (L=128 kappa scan, 8 kappa x 1024 chains/GPU = 64 kappa over 8 GPUs), c5 (one L=4096 lattice; replicas
with N GPUs) -- each with hot AND cold starts, its own roofline and, on rank 0 at N=1, the reference's own
generators timed on the host cores.  With N GPUs every rank runs the same per-GPU shard on its own chains
(weak scaling; chains are independent, there is no collective on the hot path; a final all_gather of the
observable records runs outside the timed region).

Timing protocol (every shape): the clock sampler is started and its FIRST sample awaited (NVML start-up
is then over); ranks meet at a barrier; the step is issued back to back for >= 1 s (the warm-up: at least
W steps, and the clocks settle under the power cap); WITHOUT draining the stream, 5 windows of exactly K
steps follow, each bracketed by CUDA events on the launching stream; after a final barrier + synchronize
the per-window times are reduced with MAX over ranks and the BEST window is reported (all five are in
`windows_ms_per_step`).  A step is ONE kernel launch; inputs are larger than L2 (c2 / c3 rotate over 4 chain
sets, c4 / c5 are larger than L2 by themselves), so every step streams from HBM.

One JSON line on stdout; see the task contract for the keys.  `--impl reference` times the UNMODIFIED
reference's NeighborhoodUpdate.step (staged under oracle/_ref by oracle/stage_reference.py) on all host
cores; where the staged copy is missing it falls back to the oracle's numpy port and says so (`kind`).
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
KAPPA, W_CONSTRAINT = 0.5, 1
WINDOWS = 5                          # timed windows of --steps steps; the best one is the record
SUSTAIN_S = 1.0                      # seconds of the same load issued before the first window
SEED = 20260101

# name -> the shape (per GPU).  bytes: algorithmic bytes per update; rotate: independent chain sets cycled
# through so that every step streams from HBM; cap: most steps per window (long steps: keep the default run short)
SHAPES = {
    'c2': dict(kind='villain', L=32, chains=4096, bytes=32, rotate=4, thermalise=200, cap=None,
               kernel='villain_smem_filtered_kernel<N=32, 128 threads>',
               workload='config2: Villain (phi,n) L=32 kappa=0.5 W=1, 4096 chains/GPU, NeighborhoodUpdate checkerboard sweep + action/winding/wrapping'),
    'c3': dict(kind='worldline', L=64, chains=1024, bytes=24, rotate=4, thermalise=200, cap=None,
               kernel='worldline_smem_table_kernel<N=64>',
               workload='config3: Worldline (m,v) L=64 kappa=0.5 W=1, 1024 chains/GPU (8192 over 8 GPUs), PlaquetteUpdate move in checkerboard order + observables'),
    'c4': dict(kind='villain', L=128, chains=8192, bytes=32, rotate=1, thermalise=20, cap=100, kappas_per_gpu=8,
               kernel='villain_strips_kernel<N=128, one chain per CTA of 512 threads, phi and n streamed through a ring of 8-row strips>',
               workload='config4: Villain L=128 kappa scan, 8 kappa x 1024 chains/GPU (64 kappa in [0.3,1.2] over 8 GPUs), NeighborhoodUpdate sweep + observables'),
    'c5': dict(kind='villain', L=4096, chains=1, bytes=32, rotate=1, thermalise=20, cap=200,
               kernel='villain_tile_pass_kernel (TMA tensor-staged 16 x 128 tiles, one launch per colour, in place)',
               workload='config5: Villain L=4096 kappa=0.5, one lattice per GPU (replicas only), NeighborhoodUpdate sweep + observables'),
}
WORKLOAD = SHAPES['c2']['workload']


def config_dict(world, sweeps_per_step):
    """The headline `config`: the same dict in both arms (`--impl reference` describes the workload it is the baseline of)."""
    s = SHAPES['c2']
    return {'workload': WORKLOAD, 'L': s['L'], 'kappa': KAPPA, 'W': W_CONSTRAINT, 'chains_per_gpu': s['chains'],
            'sweeps_per_step': sweeps_per_step,
            'start': f'hot (phi~U(-pi,pi), n~integers(-2,3)) + {s["thermalise"]} untimed sweeps',
            'rng': 'philox4x32-10 in-kernel',
            'l2': f'inputs larger than L2: {s["rotate"]} chain sets ({s["rotate"] * s["chains"] * s["L"] ** 2 * 16 >> 20} MiB) rotated',
            'parallelism': f'chains sharded over {world} GPU(s), no hot-path collective'}


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
# CPU arm: the reference's own generators (oracle/_ref), one chain per process on every host core
# ---------------------------------------------------------------------------------------------
def reference_staged():
    from oracle import refimport
    return refimport.available()


def _cpu_worker(args):
    """Sweeps of ONE chain by the reference's generator `what` at lattice size L; returns the seconds `sweeps` sweeps took.

    what: 'neighborhood' (generator/villain/neighborhood.py:59-137), 'plaquette' (generator/worldline/plaquette.py:35-104,
    the sequential sweep the API names), 'vortex+coexact' (worldline/vortex.py:51-136 + coexact.py:53-128, the reference's
    own checkerboard form of the same move), or 'port' (oracle/villain_np.py, only when the reference is not staged)."""
    what, L, seed, sweeps, threads = args
    os.environ['NUMBA_NUM_THREADS'] = str(threads)
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    rng = np.random.default_rng(1000 + seed)
    if what == 'port':
        from oracle import villain_np as V
        draw = np.random.default_rng(seed)
        phi, n = V.hot_start(rng, L)
        step = lambda c: V.neighborhood_step(c[0], c[1], KAPPA, W_CONSTRAINT, draw)      # noqa: E731
        cfg = (phi, n)
    else:
        import warnings
        warnings.simplefilter('ignore')
        from oracle import refimport
        sv = refimport.import_reference()
        lattice = sv.lattice.Lattice2D(L)
        if what == 'neighborhood':
            S = sv.action.Villain(lattice, KAPPA, W_CONSTRAINT)
            G = sv.generator.villain.NeighborhoodUpdate(S)
            cfg = {'phi': lattice.form(0), 'n': lattice.form(1, dtype=int)}
            cfg['phi'][...] = rng.uniform(-np.pi, np.pi, cfg['phi'].shape)
            cfg['n'][...] = rng.integers(-2, 3, cfg['n'].shape)
        else:
            S = sv.action.Worldline(lattice, KAPPA, W_CONSTRAINT)
            if what == 'plaquette':
                G = sv.generator.worldline.PlaquetteUpdate(S)
            else:
                G = sv.generator.combining.Sequentially((sv.generator.worldline.VortexUpdate(S), sv.generator.worldline.CoexactUpdate(S)))
            cfg = S.configurations(1)[0]                  # cold: delta m = 0 holds trivially; the sweep cost does not depend on the state
        step = G.step
    cfg = step(cfg)                                       # untimed first sweep (imports, JIT, colour tables)
    t0 = time.perf_counter()
    for _ in range(sweeps):
        cfg = step(cfg)
    return time.perf_counter() - t0


def cpu_rate(what, L, procs, seconds, threads=1, pool=None):
    """updates/s of `procs` processes each sweeping its own chain for about `seconds` of wall clock (calibrated on a short
    untimed run, so the timed sample is bounded).  Returns (rate, wall, sweeps_per_process)."""
    own = pool is None
    if own:
        pool = mp.get_context('spawn').Pool(procs)
    try:
        if L >= 1024:
            sweeps = 1                                    # a sweep takes seconds: no probe, one timed sweep after the untimed first
        else:
            probe = max(pool.map(_cpu_worker, [(what, L, i, 3, threads) for i in range(procs)])) / 3
            sweeps = max(1, int(seconds / max(probe, 1e-6)))
        t0 = time.perf_counter()
        inner = pool.map(_cpu_worker, [(what, L, i, sweeps, threads) for i in range(procs)])
        wall = time.perf_counter() - t0
    finally:
        if own:
            pool.close()
            pool.join()
    # the rate of the sweeps themselves: each worker's own clock around its loop (the pool's wall clock also holds the import
    # and JIT of the reference, which are not part of any sweep)
    return sum(sweeps * L * L / t for t in inner), wall, sweeps


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def headline_cpu_baseline(cores, seconds):
    """config 2 on the host: the reference's NeighborhoodUpdate, one chain per core."""
    if reference_staged():
        rate, wall, sweeps = cpu_rate('neighborhood', 32, cores, seconds)
        return {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'reference',
                'sample': f'{cores} processes (NUMBA_NUM_THREADS=1) x {sweeps} sweeps of one L=32 kappa=0.5 chain each through the unmodified '
                          f'supervillain.generator.villain.NeighborhoodUpdate.step (oracle/_ref), {wall:.1f} s wall incl. import + JIT'}
    rate, wall, sweeps = cpu_rate('port', 32, cores, seconds)
    return {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': f'{cores} processes x {sweeps} sweeps of one L=32 kappa=0.5 chain each, {wall:.1f} s wall (oracle/villain_np.py port of '
                      f'neighborhood.py:59-137; the reference is not staged under oracle/_ref)'}


def shape_cpu_baselines(name, cores):
    """The reference's generators at the shape's lattice size (bounded samples); c2's is the headline `cpu_baseline`."""
    if not reference_staged():
        return None
    L = SHAPES[name]['L']
    out = {}
    if name == 'c3':
        for what, secs in (('plaquette', 1.0), ('vortex+coexact', 3.0)):
            rate, wall, sweeps = cpu_rate(what, L, cores, secs)
            out[what] = {'value': rate, 'unit': 'plaquette-updates/s', 'cores': cores, 'kind': 'reference',
                         'sample': f'{cores} processes x {sweeps} sweeps of one L={L} chain each, {wall:.1f} s wall incl. import + JIT'}
    elif name == 'c4':
        rate, wall, sweeps = cpu_rate('neighborhood', L, cores, 3.0)
        out['neighborhood'] = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'reference',
                               'sample': f'{cores} processes x {sweeps} sweeps of one L={L} chain each, {wall:.1f} s wall incl. import + JIT'}
    elif name == 'c5':
        # one lattice: ONE process, the reference's numba kernels run parallel above 30 000 sites (lattice/_kernels.py:10,25)
        rate, wall, sweeps = cpu_rate('neighborhood', L, 1, 1.0, threads=cores)
        out['neighborhood'] = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'reference',
                               'sample': f'1 process, NUMBA_NUM_THREADS={cores}, {sweeps} sweep(s) of the L={L} lattice, {wall:.1f} s wall incl. '
                                         f'import, JIT and the colour tables'}
    return out


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    cores = host_cores()
    # bounded: the whole --steps/--warmup run stays within a few minutes whatever K and W are
    steps, warmup = min(args.steps, 20), min(args.warmup, 3)
    seconds = min(3.0, 60.0 / max(steps, 1))
    what = 'neighborhood' if reference_staged() else 'port'
    kind = 'reference' if what == 'neighborhood' else 'port'
    with mp.get_context('spawn').Pool(cores) as pool:             # one pool: the reference is imported and JIT-compiled once per core
        for _ in range(warmup):
            cpu_rate(what, 32, cores, 0.2, pool=pool)
        t_all = time.perf_counter()
        rates, sweeps = [], 0
        for _ in range(steps):
            rate, _, sweeps = cpu_rate(what, 32, cores, seconds, pool=pool)
            rates.append(rate)
        total_wall = time.perf_counter() - t_all
    value = float(np.mean(rates))
    source = ('the unmodified supervillain.generator.villain.NeighborhoodUpdate.step, staged under oracle/_ref' if kind == 'reference'
              else 'oracle/villain_np.py, the numpy port of neighborhood.py:59-137 (oracle/_ref is not staged)')
    sample = (f'{steps} timed steps; each step = {cores} processes (NUMBA_NUM_THREADS=1) x ~{sweeps} sweeps of one L=32 kappa=0.5 chain each '
              f'(~{seconds:.1f} s) through {source}')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total_wall / max(args.steps, 1),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': config_dict(int(os.environ.get('WORLD_SIZE', '1')), args.sweeps_per_step),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons every 100 ms from ONE long-lived `nvidia-smi -lms` process.  `start_and_wait` returns
    only once the first sample has arrived: NVML start-up (hundreds of ms of driver locks) is then behind us and cannot
    land inside a timed window."""
    QUERY = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []               # (timestamp, fields)
        self.first = threading.Event()
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
                    self.first.set()
                if self.stop_flag.is_set():
                    break
        except Exception:
            pass
        self.first.set()

    def start_and_wait(self, timeout=20.0):
        self.start()
        self.first.wait(timeout)

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
class Shape:
    """One named shape resident on this rank's GPU: `step(k)` issues step k (one kernel launch), `cold()` zeroes the fields."""

    def __init__(self, name, world, rank, dev, sweeps_per_step, overlap=True):
        import torch
        import supervillain_b200 as svb
        from supervillain_b200 import ops, sharding
        from supervillain_b200.generator.villain import NeighborhoodUpdate
        from supervillain_b200.generator.worldline import PlaquetteUpdate
        self.torch, self.ops, self.name, self.spec = torch, ops, name, SHAPES[name]
        s = self.spec
        self.L, self.chains, self.sweeps_per_step = s['L'], s['chains'], sweeps_per_step
        self.updates_per_step = self.chains * self.L * self.L * sweeps_per_step
        self.kappa_chain = None
        self.kappas = None
        if name == 'c4':
            per_kappa = self.chains // s['kappas_per_gpu']
            self.kappas = 0.3 + 0.9 * np.arange(s['kappas_per_gpu'] * world) / max(s['kappas_per_gpu'] * world - 1, 1)
            self.chain0, count, kc = sharding.kappa_scan(self.kappas, per_kappa, world, rank)
            self.kappa_chain = torch.as_tensor(np.asarray(kc, dtype=np.float64)).to(dev)
        else:
            self.chain0, count = sharding.shard_chains(world * self.chains, world, rank)
        assert count == self.chains
        lattice = svb.Lattice2D(self.L)
        if s['kind'] == 'villain':
            self.action = svb.Villain(lattice, KAPPA, W=W_CONSTRAINT)
            self.G = NeighborhoodUpdate(self.action, seed=SEED)
            nobs = ops.VOBS_COUNT
        else:
            self.action = svb.Worldline(lattice, KAPPA, W=W_CONSTRAINT)
            self.G = PlaquetteUpdate(self.action, seed=SEED)
            nobs = ops.WOBS_COUNT
        self.sets = []
        for r in range(s['rotate']):
            E = svb.BatchedEnsemble(self.action, self.chains, chain0=self.chain0)
            self.sets.append(E._start('hot', SEED + 7919 * (rank * s['rotate'] + r)))
        self.obs = [torch.zeros((self.chains, nobs), dtype=torch.float64, device=dev) for _ in self.sets]
        self.prev = [torch.zeros_like(o) for o in self.obs]
        self.mode = 'ordinary'
        self.steppers = None
        G, kc = self.G, self.kappa_chain
        for a, b in self.sets:                                    # untimed thermalisation of the synthetic hot starts
            G.sweep_device(a, b, s['thermalise'], chain0=self.chain0, kappa_chain=kc)
        if overlap:
            self.mode = ('in place, one launch per colour pass + one over n for sum (dn)^2 (svb_villain_sweep_inplace)' if name == 'c5'
                         else 'overlapped (programmatic dependent launch + per-chain epochs)')
            self.steppers = [G.overlapped_device(a, b, chain0=self.chain0, kappa_chain=kc) for a, b in self.sets]
        else:
            self.plans = [G.plan_device(a, b, obs=o, chain0=self.chain0, kappa_chain=kc) for (a, b), o in zip(self.sets, self.obs)]

    def step(self, k):
        r = k % len(self.sets)
        sps = self.sweeps_per_step
        if self.steppers is None:
            self.plans[r](sps)
        elif self.spec['kind'] == 'villain':
            # the way BatchedEnsemble.generate steps: launch k writes its counters to record k and completes the state
            # columns of record k - 1 (the observables of the chains as they arrive ride along with the residual build)
            self.steppers[r](sps, self.obs[r], self.prev[r])
            self.obs[r], self.prev[r] = self.prev[r], self.obs[r]
        else:
            self.steppers[r](sps, self.obs[r])

    def cold(self):
        """All-zero fields (the reference's 'cold' start, ensemble.py:80-81); the next launch waits for the memsets."""
        for a, b in self.sets:
            a.zero_()
            b.zero_()
        for st in self.steppers or ():
            fence = getattr(st, 'fence', None)
            if fence is not None:
                fence()


def time_windows(torch, dist, world, dev, step, steps, warmup, sustain_s, before_window=None):
    """The timing protocol of the module docstring.  Returns (best ms per window, [ms per window], t_load0, t_load1)."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    t_load0 = time.time()
    k = 0
    if before_window is None:
        while k < warmup or time.time() < t_load0 + sustain_s:    # the warm-up: the same load, back to back, not drained
            step(k)
            k += 1
    events = []
    for _ in range(WINDOWS):
        if before_window is not None:                             # cold starts: reset, then W warm-up steps, per window
            before_window()
            for _ in range(warmup):
                step(k)
                k += 1
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            step(k)
            k += 1
        b.record()
        events.append((a, b))
    barrier()
    t_load1 = time.time()
    t = torch.tensor([a.elapsed_time(b) for a, b in events], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = [float(x) for x in t.tolist()]
    return min(ms), ms, t_load0, t_load1


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import supervillain_b200 as svb
    from supervillain_b200.hostpath import HostStepper
    from supervillain_b200 import sharding

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)
    peak, peak_src = measured_peaks()
    names = [c for c in args.configs.split(',') if c]
    unknown = [c for c in names if c not in SHAPES]
    if unknown or 'c2' not in names:
        raise SystemExit(f'--configs must list c2 (the headline) and only {sorted(SHAPES)}; got {args.configs!r}')

    sampler = ClockSampler(local)
    sampler.start_and_wait()                                     # NVML start-up is over before anything is timed

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(shape, steps, start):
        before = shape.cold if start == 'cold' else None
        if before is not None:
            shape.cold()
        best, ms, t0, t1 = time_windows(torch, dist, world, dev, shape.step, steps, args.warmup, SUSTAIN_S, before)
        launch_s = best * 1e-3 / steps
        achieved = shape.spec['bytes'] * shape.updates_per_step / launch_s / 1e9
        return {'start': start, 'value': world * shape.updates_per_step / launch_s,
                'ms_per_step': best / steps, 'steps': steps, 'windows_ms_per_step': [m / steps for m in ms],
                'roofline_frac': achieved / peak, 'achieved_gbs': achieved, 'clocks': sampler.summary(t0, t1)}

    # ---- headline: config 2, hot start ----
    c2 = Shape('c2', world, rank, dev, args.sweeps_per_step, overlap=not args.no_overlap)
    head = measure(c2, args.steps, 'hot')
    ms_per_step = head['ms_per_step']
    value = head['value']
    spec = SHAPES['c2']
    roofline = {'bound': 'hbm', 'achieved': head['achieved_gbs'], 'peak': peak, 'unit': 'GB/s', 'frac': head['roofline_frac'],
                'traffic': traffic_from_profile(args.traffic), 'kernel': spec['kernel'], 'launches': c2.mode,
                'algorithmic_bytes_per_launch': spec['bytes'] * c2.updates_per_step, 'peak_source': peak_src,
                'sweeps_per_launch': args.sweeps_per_step}
    clocks = head['clocks']
    # ---- gather of observables of the thermalised chains (outside the timed region; the only inter-GPU traffic) ----
    obs = svb.ops.villain_observables(c2.sets[0][0], c2.sets[0][1], KAPPA)      # state of set 0 after the hot windows
    all_obs = sharding.gather_columns(obs)                    # (world * CHAINS, VOBS_COUNT) on every rank
    mean_action_density = float(all_obs[:, 0].mean().item()) / (spec['L'] ** 2)
    c2_cold = measure(c2, args.steps, 'cold')

    # ---- end to end through the host-buffer API: H2D of the fields, sweep, D2H of fields + observables ----
    G = c2.G
    stepper = HostStepper(G, c2.chains, chain0=c2.chain0)
    host_sets = [stepper.pinned_fields(from_device=s) for s in c2.sets[:2]]
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
    e2e_value = world * c2.updates_per_step * e2e_steps / float(t.item())
    e2e = {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': stepper.h2d_bytes, 'd2h_bytes_per_step': stepper.d2h_bytes,
           'steps': e2e_steps, 'api': 'HostStepper.step_async(phi_host, n_host): pinned host fields in and out + observables, two steps in flight on '
                  'alternating host buffer pairs'}

    torch.cuda.synchronize()
    del stepper, host_sets

    # ---- every named shape: hot and cold, its own roofline ----
    def entry(name, shape, hot, cold):
        s = SHAPES[name]
        unit = UNIT if s['kind'] == 'villain' else 'plaquette-updates/s'
        e = {'name': name, 'workload': s['workload'], 'unit': unit, 'value': hot['value'], 'ms_per_step': hot['ms_per_step'],
             'steps': hot['steps'], 'windows_ms_per_step': hot['windows_ms_per_step'],
             'L': s['L'], 'chains_per_gpu': s['chains'], 'chains_total': s['chains'] * world, 'launches': shape.mode,
             'roofline': {'bound': 'hbm', 'achieved': hot['achieved_gbs'], 'peak': peak, 'unit': 'GB/s', 'frac': hot['roofline_frac'],
                          'bytes_per_update': s['bytes'], 'algorithmic_bytes_per_launch': s['bytes'] * shape.updates_per_step,
                          'kernel': s['kernel']},
             'start': 'hot', 'clocks': hot['clocks'],
             'cold': {k: cold[k] for k in ('value', 'ms_per_step', 'steps', 'windows_ms_per_step', 'roofline_frac')},
             'l2': (f'{s["rotate"]} chain sets rotated' if s['rotate'] > 1 else 'one set') +
                   f' ({s["rotate"] * s["chains"] * s["L"] ** 2 * (s["bytes"] // 2) >> 20} MiB of state per GPU: larger than L2)'}
        if shape.kappas is not None:
            e['kappas_total'] = len(shape.kappas)
        return e

    shapes_out = [entry('c2', c2, head, c2_cold)]
    del c2, G
    torch.cuda.empty_cache()
    for name in names:
        if name == 'c2':
            continue
        cap = SHAPES[name]['cap']
        steps = args.steps if cap is None else min(args.steps, cap)
        shape = Shape(name, world, rank, dev, args.sweeps_per_step)
        hot = measure(shape, steps, 'hot')
        cold = measure(shape, steps, 'cold')
        shapes_out.append(entry(name, shape, hot, cold))
        del shape
        torch.cuda.empty_cache()
    sampler.stop()

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = host_cores()
            cpu = headline_cpu_baseline(cores, 12.0)
            for e in shapes_out:
                if e['name'] != 'c2' and args.cpu_baselines == 'all':
                    e['cpu_baseline'] = shape_cpu_baselines(e['name'], cores)
            shapes_out[0]['cpu_baseline'] = cpu
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': config_dict(world, args.sweeps_per_step),
            'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': args.steps,
            'clocks': clocks,
            'timing': {'windows': WINDOWS, 'reported': 'best window (min over windows of the max over ranks)',
                       'windows_ms_per_step': head['windows_ms_per_step'],
                       'warmup': f'>= {args.warmup} steps and >= {SUSTAIN_S} s of the same load issued back to back; the stream is not drained '
                                 f'before the first window; clock sampler started (first sample awaited) before the warm-up'},
            'configs': shapes_out,
            'check': {'mean_action_density': mean_action_density, 'gathered_chains': int(all_obs.shape[0]),
                      'state': 'set 0 after the hot-start windows'},
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
    ap.add_argument('--configs', default='c2,c3,c4,c5', help='named shapes to run (c2, the headline, always)')
    ap.add_argument('--sweeps-per-step', type=int, default=1)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-baselines', default='all', choices=['all', 'headline'],
                    help='reference generators timed on the host for every shape, or for config 2 only')
    ap.add_argument('--e2e-steps', type=int, default=20, help='timed steps of the host-buffer (e2e) leg')
    ap.add_argument('--no-overlap', action='store_true', help='ordinary launches instead of overlapped ones (config 2)')
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
