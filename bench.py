#!/usr/bin/env python
"""Headline benchmark: AA and GPNH outer (alternating) iterations per second at HadISST shape.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload both|aa|gpnh] [--no-stress] [--no-strong] [--no-numba]

A "step" is one outer alternating iteration (dictionary update + weights update + cost
bookkeeping) of the reference's loops -- `_iterate_aa` (archetypal_analysis.py:534-670, the
workload that owns the outer-level SPG; the drivers' `dictionary_solver_kwargs =
dict(max_iterations=1)`, bin/run_hadisst_aa.py:160-166) and `_iterate_gpnh_convex_coding`
(gpnh_convex_coding.py:282-402) -- on a synthetic anomaly matrix of HadISST shape (1620
training months x 44 000 ocean cells, fp64, k = 8): BASELINE.json configs[0] / configs[1].

One JSON line (rank 0).  The top-level keys describe the AA workload (BASELINE.json's metric
names it first); the GPNH workload sits beside it under "gpnh" with the same keys:
  value         whole-job outer iterations / second with X resident in HBM, device timed
  e2e           the same through the public NumPy-in / NumPy-out call: host buffers, H2D of X
                and D2H of the factors inside the timed region
  roofline      the slower streaming pass over X against the measured HBM peak
  cpu_baseline  the CPU oracle port of the reference loop timed on this box's host cores
  parity        CUDA path vs the CPU port after the same number of iterations from the same
                start (asserted: the run exits non-zero when it fails)
  time_to_converge   one public fit to the drivers' stopping rule
With N > 1 the sample axis is sharded.  "value" is WEAK scaling (every rank owns a full
1620-row slab; the job is one fit of N x 1620 samples); "strong_scaling" adds the fixed-size
figures: the 1620-row HadISST matrix and the 18 000 x 44 000, k = 64 stress shape
(BASELINE.json configs[4]) split over the N ranks.  "--impl reference" runs the CPU arm on
the same total number of samples.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, 'matrix-factorization-case-studies_b200')
for _p in (ROOT, PKG_DIR):
    if _p not in sys.path:
        sys.path.insert(0, _p)

if '--impl' in sys.argv and sys.argv[sys.argv.index('--impl') + 1:][:1] == ['reference']:
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every
    # host core, so the BLAS thread count is set explicitly before
    # NumPy loads
    for _v in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np   # noqa: E402

T_ROWS, N_FEATURES, N_COMPONENTS = 1620, 44000, 8
STRESS_ROWS, STRESS_COMPONENTS, STRESS_BLOCKS = 18000, 64, 8
LAMBDA_W = 0.0
PARITY_STEPS = 5
PARITY_COST_RTOL = 1e-8
PARITY_FACTOR_ATOL = 2e-5


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--workload', choices=('both', 'aa', 'gpnh'), default='both')
    ap.add_argument('--impl', choices=('b200', 'reference'), default='b200')
    ap.add_argument('--rows', type=int, default=T_ROWS)
    ap.add_argument('--features', type=int, default=N_FEATURES)
    ap.add_argument('--components', type=int, default=N_COMPONENTS)
    ap.add_argument('--formulation', choices=('stream', 'gram'), default='stream',
                    help="AA only: 'gram' builds K = X X' once and iterates on K (opt-in "
                         "algorithmic variant; the headline numbers use 'stream')")
    ap.add_argument('--cpu-steps', type=int, default=PARITY_STEPS,
                    help='outer iterations of the CPU baseline / parity sample (0 disables it)')
    ap.add_argument('--min-timed-ms', type=float, default=50.0,
                    help='the K-step block is repeated until the timed region is this long')
    ap.add_argument('--no-stress', action='store_true',
                    help='skip the 18 000 x 44 000, k = 64 stress shape (BASELINE configs[4])')
    ap.add_argument('--no-strong', action='store_true',
                    help='N > 1: skip the strong-scaling (fixed total size) measurements')
    ap.add_argument('--no-kmeans', action='store_true',
                    help='skip the k-means block (BASELINE configs[2])')
    ap.add_argument('--no-gram', action='store_true',
                    help="skip the nested AA block measured through formulation='gram'")
    ap.add_argument('--no-numba', action='store_true',
                    help='reference arm: skip timing the real (Numba) reference from baseline/_ref')
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh).get('hbm_gbs', 6650.0), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """SM clock / throttle-reason samples while the GPU is under load.

    NVML (a few hundred samples per second) when `pynvml` works, otherwise one
    `nvidia-smi` query every 0.2 s (the recipe's clocks line)."""

    QUERY = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
    REASON_BITS = {'sw_power_cap': 0x4, 'hw_slowdown': 0x8, 'sw_thermal_slowdown': 0x20,
                   'hw_thermal_slowdown': 0x40}

    def __init__(self, index=0):
        self.index = index
        self.sm, self.sm_max, self.reasons = [], [], set()
        self.source = None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def _open_nvml(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            index = self.index
            visible = os.environ.get('CUDA_VISIBLE_DEVICES')
            if visible:
                index = int(visible.split(',')[self.index])
            handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            reasons_fn = getattr(pynvml, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
                getattr(pynvml, 'nvmlDeviceGetCurrentClocksThrottleReasons')
            smax = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)
            return pynvml, handle, reasons_fn, smax
        except Exception:
            return None

    def _loop_nvml(self):
        pynvml, handle, reasons_fn, smax = self._nvml
        while not self._stop.is_set():
            try:
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)))
                self.sm_max.append(smax)
                mask = int(reasons_fn(handle))
                for name, bit in self.REASON_BITS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.005)

    def _loop_smi(self):
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        while not self._stop.is_set():
            try:
                out = subprocess.check_output(
                    ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                     '--format=csv,noheader,nounits'], timeout=5).decode().strip()
                vals = [v.strip() for v in out.split(',')]
                self.sm.append(float(vals[0]))
                self.sm_max.append(float(vals[1]))
                for name, val in zip(names, vals[2:6]):
                    if val.lower().startswith('active'):
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._nvml = self._open_nvml()
        self.source = 'nvml' if self._nvml else 'nvidia-smi'
        self._thread = threading.Thread(target=self._loop_nvml if self._nvml else self._loop_smi,
                                        daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self):
        return {'sm_mhz': float(np.median(self.sm)) if self.sm else None,
                'sm_max_mhz': float(np.max(self.sm_max)) if self.sm_max else None,
                'reasons': sorted(self.reasons), 'samples': len(self.sm), 'source': self.source}


# ---------------------------------------------------------------------------
# problems (host generated, identical for the CUDA path and the CPU arm)
# ---------------------------------------------------------------------------

def slab(args, rank):
    """Rank `rank`'s slab of the weak-scaling problem: a full HadISST-shaped matrix."""
    from convex_dim_red.datasets import synthetic_field
    return synthetic_field(args.rows, args.features, seed=rank)


def initial_factors(workload, T, d, k, rank, world):
    """(weights of this rank's rows, replicated dictionary) in the single-process draw order."""
    from convex_dim_red.stochastic_matrices import right_stochastic_matrix
    Z0 = right_stochastic_matrix((T, k), np.random.RandomState(1000 + rank))
    if workload == 'gpnh':
        return Z0, np.sqrt(0.4 / k) * np.random.RandomState(0).randn(d, k)
    return Z0, right_stochastic_matrix((k, T * world), np.random.RandomState(7))


def stacked_problem(args, workload, world):
    """The whole weak-scaling job on one host: the slabs of all ranks stacked."""
    X = np.concatenate([slab(args, r) for r in range(world)], axis=0) if world > 1 \
        else slab(args, 0)
    parts = [initial_factors(workload, args.rows, args.features, args.components, r, world)
             for r in range(world)]
    Z0 = np.concatenate([p[0] for p in parts], axis=0)
    return X, Z0, parts[0][1]


def stress_block(block, d, n_rows):
    """Rows [block * n_rows, (block + 1) * n_rows) of the stress matrix (block-seeded so that
    every split over 1 / 2 / 4 / 8 ranks sees the same 18 000 x 44 000 matrix)."""
    from convex_dim_red.stochastic_matrices import right_stochastic_matrix
    sources = np.random.RandomState(499).standard_normal((12, d))
    rs = np.random.RandomState(500 + block)
    x = right_stochastic_matrix((n_rows, 12), rs).dot(sources)
    x += 0.5 * rs.standard_normal((n_rows, d))
    return x


def stress_blocks(rank, world):
    per = STRESS_BLOCKS // world
    return range(rank * per, (rank + 1) * per), STRESS_ROWS // STRESS_BLOCKS


def stress_rows(rank, world, d):
    """This rank's rows of the stress matrix."""
    blocks, rows = stress_blocks(rank, world)
    return np.concatenate([stress_block(b, d, rows) for b in blocks], axis=0)


def stress_factors(workload, rank, world, d):
    from convex_dim_red.stochastic_matrices import right_stochastic_matrix
    k = STRESS_COMPONENTS
    blocks, rows = stress_blocks(rank, world)
    Z0 = np.concatenate([right_stochastic_matrix((rows, k), np.random.RandomState(2000 + b))
                         for b in blocks], axis=0)
    if workload == 'gpnh':
        return Z0, np.sqrt(0.4 / k) * np.random.RandomState(0).randn(d, k)
    return Z0, right_stochastic_matrix((k, STRESS_ROWS), np.random.RandomState(7))


# ---------------------------------------------------------------------------
# CPU arm (oracle port of the reference loop)
# ---------------------------------------------------------------------------

def cpu_steps(workload, X, Z0, F0, n_steps):
    """The reference loop (oracle port: NumPy/BLAS contractions + C per-sample QPs, same pass
    structure as the reference) for n_steps outer iterations from the given start."""
    from oracle import convex_oracle as orc
    trace = float(np.einsum('ij,ij->', X, X))
    times = []
    if workload == 'gpnh':
        out = orc.iterate_gpnh(X, Z0.copy(), F0.copy(), lambda_W=LAMBDA_W, tolerance=0.0,
                               max_iterations=n_steps, trace_XtX=trace,
                               require_monotonic_cost_decrease=False, iter_times_out=times)
        return times, float(out[2]), out[0], out[1]
    out = orc.iterate_aa(X, Z0.copy(), F0.copy(), np.ones(F0.shape[0]), tolerance=0.0,
                         max_iterations=n_steps, trace_XXt=trace,
                         dictionary_solver_kwargs=dict(max_iterations=1),
                         require_monotonic_cost_decrease=False, iter_times_out=times)
    return times, float(out[3]), out[0], out[1]


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p['num_threads'] for p in threadpool_info() if p.get('user_api') == 'blas']
        return max(n) if n else os.cpu_count()
    except Exception:
        return os.cpu_count()


PORT_NOTE = ('oracle port of the reference loop (NumPy/OpenBLAS passes over X with all host '
             'threads + C per-sample QPs); the real Numba reference makes 11+ passes per AA '
             'iteration and is ~30x slower (see numba_reference / profiles)')


def metric_name(workload):
    return '%s_outer_iterations_per_sec_hadisst' % workload


def config_dict(args, workload, world):
    return {'workload': '%s k=%d on synthetic HadISST-shaped anomalies, %d x %d fp64 per GPU '
                        '(BASELINE.json configs[%d])'
                        % ('GPNH convex coding' if workload == 'gpnh' else
                           'archetypal analysis (dictionary SPG max_iterations=1, as '
                           'bin/run_hadisst_aa.py:160-166)',
                           args.components, args.rows, args.features,
                           1 if workload == 'gpnh' else 0),
            'n_samples_total': args.rows * world, 'n_features': args.features,
            'n_components': args.components, 'lambda_W': LAMBDA_W,
            'dictionary_solver_kwargs': None if workload == 'gpnh' else {'max_iterations': 1},
            'formulation': ('streaming (X read from HBM every pass; 2 passes/iter GPNH, 4 AA)'
                            if args.formulation == 'stream' or workload == 'gpnh' else
                            'gram (K = X X^T built once, 2 passes over the L2-resident K per '
                            'iteration; no HBM roofline applies)'),
            'l2_policy': 'X (%.0f MB per GPU) is larger than the 126 MB L2; no explicit flush'
                         % (8e-6 * args.rows * args.features),
            'sharding': 'sample axis, %d rank(s)' % world}


def numba_reference(args, workload):
    """Time the real reference (installed unmodified under baseline/_ref) in a subprocess."""
    script = os.path.join(ROOT, 'baseline', 'time_reference.py')
    if args.no_numba or not os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'convex_dim_red')):
        return {'unavailable': 'baseline/_ref not installed or --no-numba'}
    try:
        out = subprocess.run([sys.executable, script, '--workload', workload, '--rows',
                              str(args.rows), '--features', str(args.features), '--components',
                              str(args.components), '--iterations', '2'],
                             capture_output=True, text=True, timeout=900)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as exc:              # the baseline arm must not take the bench line down
        return {'unavailable': 'baseline/time_reference.py failed: %r' % (exc,)}


def reference_line(args, workload, world):
    X, Z0, F0 = stacked_problem(args, workload, world)
    # `warmup + steps` outer iterations from the same start as the CUDA arm; the last `steps`
    # are timed (per-iteration wall times from the loop itself)
    times, cost, _, _ = cpu_steps(workload, X, Z0, F0, args.warmup + args.steps)
    timed = max(sum(times[args.warmup:]), 1e-9)
    value = args.steps / timed
    cores = blas_threads()
    line = {
        'impl': 'reference', 'metric': metric_name(workload), 'value': value,
        'unit': 'iterations/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * timed / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': config_dict(args, workload, world),
        'cpu_baseline': {'value': value, 'unit': 'iterations/s', 'cores': cores, 'kind': 'port',
                         'host_cpus': os.cpu_count(),
                         'sample': '%d outer iterations after %d untimed on all %d x %d samples; %s'
                                   % (args.steps, args.warmup, X.shape[0], X.shape[1], PORT_NOTE)},
        'e2e': {'value': value, 'unit': 'iterations/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'final_cost': cost,
    }
    return line


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    world = args.gpus
    workloads = ('aa', 'gpnh') if args.workload == 'both' else (args.workload,)
    lines = {w: reference_line(args, w, world) for w in workloads}
    if world == 1:
        for w in workloads:
            lines[w]['numba_reference'] = numba_reference(args, w)
    head = lines[workloads[0]]
    if len(workloads) > 1:
        head['gpnh'] = lines['gpnh']
    print(json.dumps(head))


# ---------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------

def parity_block(args, workload, X, Z0, F0, cpu_result):
    """CUDA path vs the CPU port after the same number of outer iterations from the same
    start, through the public `_iterate_*` call."""
    import bench_harness as bh
    times, cpu_cost, cpu_Z, cpu_F = cpu_result
    n = len(times)
    Zg, Fg, gpu_cost = bh.gpu_fit(workload, X, Z0, F0, n)
    dz = float(np.max(np.abs(Zg - cpu_Z)))
    df = float(np.max(np.abs(Fg - cpu_F)))
    rel = abs(gpu_cost - cpu_cost) / abs(cpu_cost)
    ok = bool(rel <= PARITY_COST_RTOL and dz <= PARITY_FACTOR_ATOL and df <= PARITY_FACTOR_ATOL)
    return {'ok': ok, 'iterations': n, 'gpu_cost': gpu_cost, 'cpu_cost': cpu_cost,
            'cost_rel_diff': rel, 'cost_rtol': PARITY_COST_RTOL,
            'weights_max_abs_diff': dz, 'dictionary_max_abs_diff': df,
            'factor_atol': PARITY_FACTOR_ATOL,
            'against': 'CPU oracle port, same host-generated inputs and start'}


def sharded_parity_block(args, workload, Xl, Z0l, F0, rank, world, comm):
    """N > 1: the sharded fit against a one-GPU fit of the stacked rows on rank 0."""
    import bench_harness as bh
    _, _, cost_sharded = bh.gpu_fit(workload, Xl, Z0l, F0, PARITY_STEPS, comm=comm)
    block = None
    if rank == 0:
        X, Z0, F0s = stacked_problem(args, workload, world)
        _, _, cost_one = bh.gpu_fit(workload, X, Z0, F0s, PARITY_STEPS)
        rel = abs(cost_sharded - cost_one) / abs(cost_one)
        block = {'ok': bool(rel <= PARITY_COST_RTOL), 'iterations': PARITY_STEPS,
                 'sharded_cost': cost_sharded, 'one_gpu_cost': cost_one, 'cost_rel_diff': rel,
                 'cost_rtol': PARITY_COST_RTOL,
                 'against': 'one-GPU fit of the %d stacked rows on rank 0' % X.shape[0]}
    bh.barrier(world)
    return block


def workload_block(args, workload, X, Xd, rank, world, comm, sampler, hbm_peak, peak_src):
    import bench_harness as bh
    T, d, k = args.rows, args.features, args.components
    Z0, F0 = initial_factors(workload, T, d, k, rank, world)
    res = bh.run_workload(args, workload, X, Z0, F0, Xd, rank, world, comm)
    eng, step = res.pop('engine'), res.pop('step')
    ms = res['ms_per_step']
    gram_mode = workload == 'aa' and args.formulation == 'gram'
    block = {
        'metric': metric_name(workload), 'value': world * 1e3 / ms, 'unit': 'iterations/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
        'timed_steps': res['timed_steps'], 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': config_dict(args, workload, world),
        'gpu_launches': res['launches_per_step'] * res['timed_steps'],
        'launches_per_step': res['launches_per_step'], 'final_cost': res['final_cost'],
    }
    pass_bytes = 8.0 * T * d
    if not gram_mode:
        t_s, t_f, t_qp, passes = bh.time_passes(args, workload, eng, T, d, k)
        slow, name = max((t_s, 'reduce_samples_tma_kernel'), (t_f, 'reduce_features_strip_kernel'))
        roof = {'kernel': name, 'bound': 'hbm', 'achieved': pass_bytes / (slow * 1e-3) / 1e9,
                'peak': hbm_peak, 'unit': 'GB/s', 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': pass_bytes, 'ms_per_launch': slow, 'traffic': None}
        roof['frac'] = roof['achieved'] / hbm_peak
        tpath = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
        if os.path.exists(tpath):
            with open(tpath) as fh:
                tr = json.load(fh)
            roof['traffic'] = tr['dram_bytes_per_launch'].get(name)
            roof['traffic_source'] = tr['source']
        block['roofline'] = roof
        block['kernels'] = {
            'reduce_samples_ms': t_s, 'reduce_features_ms': t_f,
            'reduce_samples_gbs': pass_bytes / (t_s * 1e-3) / 1e9,
            'reduce_features_gbs': pass_bytes / (t_f * 1e-3) / 1e9,
            'qp_batched_ms': t_qp, 'passes_per_step': passes, 'step_ms': ms,
            'non_streaming_ms_per_step': ms - passes / 2.0 * (t_s + t_f),
            'streaming_share_of_step': passes / 2.0 * (t_s + t_f) / ms,
            'whole_step_gbs': passes * pass_bytes / (ms * 1e-3) / 1e9,
            'whole_step_frac_of_hbm_peak': passes * pass_bytes / (ms * 1e-3) / 1e9 / hbm_peak}
    else:
        from convex_dim_red import _backend as be
        t_gram = bh.time_launches(lambda: be.gram(Xd, T, d), reps=3)
        block['roofline'] = None
        block['kernels'] = {'gram_build_ms': t_gram,
                            'gram_build_tflops': 2.0 * T * T * d / (t_gram * 1e-3) / 1e12,
                            'step_ms': ms}
    # keep the GPU busy a little longer so that several clock samples land under load (a fixed
    # count: every rank must issue the same collectives)
    for _ in range(300):
        step()
    bh.barrier(world)
    del eng, step
    block['e2e'] = bh.run_e2e(workload, X, Z0, F0, args.steps, world, comm)
    if world == 1:
        block['time_to_converge'] = bh.run_to_convergence(workload, X, Z0, F0,
                                                          formulation=args.formulation)
        if args.cpu_steps > 0 and rank == 0:
            cpu = cpu_steps(workload, X, Z0, F0, args.cpu_steps)
            times = cpu[0]
            block['cpu_baseline'] = {
                'value': len(times) / sum(times), 'unit': 'iterations/s', 'cores': blas_threads(),
                'host_cpus': os.cpu_count(), 'kind': 'port',
                'sample': 'first %d outer iterations of the same workload from the same start; %s'
                          % (len(times), PORT_NOTE),
                'final_cost': cpu[1]}
            block['parity'] = parity_block(args, workload, X, Z0, F0, cpu)
            block['time_to_converge']['cpu_estimate_seconds'] = \
                block['time_to_converge']['iterations'] * sum(times) / len(times)
    else:
        block['parity'] = sharded_parity_block(args, workload, X, Z0, F0, rank, world, comm)
    return block


def gram_block(args, X, Xd):
    """AA through the opt-in Gram formulation: K = X X' is built once by
    the SYRK kernel and the alternating loop runs on the L2-resident K -- the same iteration
    in exact arithmetic, so it is reported next to the streaming numbers, never as them."""
    import copy
    import bench_harness as bh
    from convex_dim_red import _backend as be
    T, d, k = args.rows, args.features, args.components
    Z0, C0 = initial_factors('aa', T, d, k, 0, 1)
    gargs = copy.copy(args)
    gargs.formulation = 'gram'
    res = bh.run_workload(gargs, 'aa', X, Z0, C0, Xd, 0, 1, None)
    res.pop('engine')
    res.pop('step')
    t_gram = bh.time_launches(lambda: be.gram(Xd, T, d), reps=3)
    ttc = bh.run_to_convergence('aa', X, Z0, C0, formulation='gram')
    ref = bh.gpu_fit('aa', X, Z0, C0, 5)
    from convex_dim_red import archetypal_analysis as aa
    out = aa._iterate_aa(X, Z0, C0, np.ones(k), tolerance=0.0, max_iterations=5,
                         dictionary_solver_kwargs=bh.DICT_KW, require_monotonic_cost_decrease=False,
                         formulation='gram')
    return {'what': "archetypal analysis with formulation='gram': K = X X' (T x T) built once, "
                    'iterations on K (general kernel sequence)',
            'value': 1e3 / res['ms_per_step'], 'unit': 'iterations/s', 'ms_per_step': res['ms_per_step'],
            'launches_per_step': res['launches_per_step'], 'gram_build_ms': t_gram,
            'gram_build_tflops': 2.0 * T * T * d / (t_gram * 1e-3) / 1e12,
            'gram_build_executed_tflops': 1.0 * T * T * d / (t_gram * 1e-3) / 1e12,   # upper triangle
            'time_to_converge': ttc,
            'agreement_with_streaming': {
                'iterations': 5, 'streaming_cost': ref[2], 'gram_cost': float(out[3]),
                'cost_rel_diff': abs(float(out[3]) - ref[2]) / abs(ref[2]),
                'weights_max_abs_diff': float(np.max(np.abs(out[0] - ref[0])))}}


def fixed_size_block(args, name, workload, X, Xd, Z0, F0, rank, world, comm, flops_per_step):
    """A fixed-total-size (strong scaling) measurement: this rank's rows of the problem."""
    import bench_harness as bh
    res = bh.run_workload(args, workload, X, Z0, F0, Xd, rank, world, comm)
    res.pop('engine')
    res.pop('step')
    ms = res['ms_per_step']
    out = {'workload': name, 'value': 1e3 / ms, 'unit': 'iterations/s', 'ms_per_step': ms,
           'timed_steps': res['timed_steps'], 'rows_per_gpu': int(X.shape[0]), 'n_gpus': world,
           'launches_per_step': res['launches_per_step'], 'final_cost': res['final_cost'],
           'tflops': flops_per_step / (ms * 1e-3) / 1e12,
           'gbs_streaming': flops_per_step / (2.0 * Z0.shape[1]) * 8.0 / (ms * 1e-3) / 1e9}
    return out


def strong_scaling(args, rank, world, comm):
    """Fixed-size problems split over the ranks: the HadISST matrix (1620 rows, k = 8) and --
    unless --no-stress -- the 18 000 x 44 000, k = 64 stress shape (BASELINE configs[4])."""
    from convex_dim_red import _backend as be
    from convex_dim_red._dist import shard_bounds
    from convex_dim_red.datasets import synthetic_field
    out = {}
    d, k = args.features, args.components
    if world > 1 and not args.no_strong:
        X = synthetic_field(args.rows, d, seed=0)
        lo, hi = shard_bounds(args.rows, world, rank)
        Xl = np.ascontiguousarray(X[lo:hi])
        del X
        Xd = be.to_device_padded(Xl)
        for wl in ('aa', 'gpnh'):
            Zfull, F0 = initial_factors(wl, args.rows, d, k, 0, 1)
            passes = 4 if wl == 'aa' else 2
            out['hadisst_' + wl] = fixed_size_block(
                args, '%s k=%d, %d x %d split over %d GPUs' % (wl, k, args.rows, d, world), wl,
                Xl, Xd, np.ascontiguousarray(Zfull[lo:hi]), F0, rank, world, comm,
                passes * 2.0 * k * args.rows * d)
        del Xd
        be.torch_mod().cuda.empty_cache()
    if not args.no_stress and STRESS_BLOCKS % world == 0:
        Xs = stress_rows(rank, world, d)
        Xd = be.to_device_padded(Xs)
        for wl in ('gpnh', 'aa'):
            Z0, F0 = stress_factors(wl, rank, world, d)
            passes = 4 if wl == 'aa' else 2
            out['stress_' + wl] = fixed_size_block(
                args, '%s k=%d, %d x %d split over %d GPU(s) (BASELINE.json configs[4])'
                % (wl, STRESS_COMPONENTS, STRESS_ROWS, d, world), wl, Xs, Xd, Z0, F0, rank,
                world, comm, passes * 2.0 * STRESS_COMPONENTS * STRESS_ROWS * d)
        del Xd, Xs
        be.torch_mod().cuda.empty_cache()
    return out


def kmeans_block(args, hbm_peak, peak_src):
    """BASELINE.json configs[2]: k-means k = 8 with FurthestSum initialisation on a synthetic
    JRA-55 hgt500-shaped field (700 months x 41 800 grid points), scikit-learn's KMeans with the
    same initial centres beside it (the third-party code the drivers call,
    bin/run_hadisst_kmeans.py:128-131)."""
    import torch
    from sklearn.cluster import KMeans as SkKMeans
    from convex_dim_red.datasets import synthetic_field
    from convex_dim_red.kmeans import furthest_sum_centres, kmeans_lloyd
    T, d, k = 700, 41800, 8
    X = synthetic_field(T, d, seed=5)
    start = int(np.random.RandomState(0).randint(T))
    picks = furthest_sum_centres(X, k, start, 10)
    t0 = time.perf_counter()
    picks = furthest_sum_centres(X, k, start, 10)
    torch.cuda.synchronize()
    t_init = time.perf_counter() - t0
    init = X[picks].copy()
    kmeans_lloyd(X, init, tol=1e-4, max_iter=10000)               # warm-up
    fits = []
    for _ in range(3):
        stats = {}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        labels, centres, inertia, n_iter = kmeans_lloyd(X, init, tol=1e-4, max_iter=10000,
                                                        stats=stats)
        torch.cuda.synchronize()
        fits.append((time.perf_counter() - t0, stats))
    fit_s, stats = sorted(fits, key=lambda f: f[0])[1]
    # steady-state Lloyd iterations (the synthetic field converges in a handful, so the
    # iteration throughput is timed on 200 re-armed iterations from the initial centres)
    ms_iter = kmeans_lloyd(X, init, tol=1e-4, max_iter=10000, _time_iterations=200)
    pass_bytes = 8.0 * T * d
    t0 = time.perf_counter()
    sk = SkKMeans(n_clusters=k, init=init, n_init=1, algorithm='lloyd', tol=1e-4,
                  max_iter=10000).fit(X)
    t_sk = time.perf_counter() - t0
    return {
        'workload': 'k-means k=%d, FurthestSum init, %d x %d fp64 (BASELINE.json configs[2])' % (k, T, d),
        'metric': 'kmeans_lloyd_iterations_per_sec_jra55', 'value': 1e3 / ms_iter,
        'unit': 'iterations/s', 'ms_per_lloyd_iteration': ms_iter, 'n_iter': int(n_iter),
        'device_loop': stats['device_loop'], 'fit_loop_ms': stats['loop_ms'],
        'timed': '200 graph-replayed Lloyd iterations (4 kernels each), CUDA events',
        'roofline': {'bound': 'hbm', 'kernel': 'the two streaming passes of a Lloyd iteration',
                     'algorithmic_bytes_per_iteration': 2 * pass_bytes,
                     'achieved': 2 * pass_bytes / (ms_iter * 1e-3) / 1e9, 'peak': hbm_peak,
                     'unit': 'GB/s', 'frac': 2 * pass_bytes / (ms_iter * 1e-3) / 1e9 / hbm_peak,
                     'peak_source': peak_src,
                     'note': 'whole iteration (4 kernels: two streaming passes, assignment, centre update); '
                             'X (234 MB) is larger than the L2'},
        'e2e': {'value': n_iter / fit_s, 'unit': 'iterations/s', 'fit_seconds': fit_s,
                'h2d_bytes_per_step': X.nbytes / max(n_iter, 1),
                'd2h_bytes_per_step': (labels.nbytes + centres.nbytes) / max(n_iter, 1),
                'call': 'kmeans_lloyd(X, init): upload, centring, Lloyd iterations, labels and '
                        'centres back; median of 3'},
        'furthest_sum_init_seconds': t_init,
        'cpu_baseline': {'value': sk.n_iter_ / t_sk, 'unit': 'iterations/s', 'kind': 'reference',
                         'cores': os.cpu_count(), 'fit_seconds': t_sk, 'n_iter': int(sk.n_iter_),
                         'sample': 'sklearn.cluster.KMeans(init=same centres, n_init=1, '
                                   "algorithm='lloyd').fit(X), the whole fit"},
        'parity': {'ok': bool(np.array_equal(labels, sk.labels_) and n_iter == sk.n_iter_),
                   'labels_equal': bool(np.array_equal(labels, sk.labels_)),
                   'n_iter': [int(n_iter), int(sk.n_iter_)],
                   'inertia_rel_diff': abs(inertia - sk.inertia_) / sk.inertia_,
                   'against': 'scikit-learn 1.9.0 KMeans, same initial centres'}}


def run_b200(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        # the exchanged payloads are small (k x d = 2.8 MB, k x k): the low-latency protocol
        # measured 40 us vs 52 us for the 2.8 MB all-reduce on 8 GPUs (profiles/bench_allreduce.py)
        os.environ.setdefault('NCCL_PROTO', 'LL')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from convex_dim_red import _backend as be
    from convex_dim_red._dist import Comm
    comm = Comm() if world > 1 else None
    hbm_peak, peak_src = peaks()
    workloads = ('aa', 'gpnh') if args.workload == 'both' else (args.workload,)

    X = slab(args, rank)
    Xd = be.to_device_padded(X)
    sampler = ClockSampler(local_rank)
    sampler.__enter__()
    blocks = {}
    for wl in workloads:
        blocks[wl] = workload_block(args, wl, X, Xd, rank, world, comm, sampler, hbm_peak, peak_src)
    sampler.__exit__(None, None, None)
    gram = None
    if world == 1 and 'aa' in workloads and args.formulation == 'stream' and not args.no_gram:
        gram = gram_block(args, X, Xd)
    del Xd
    torch.cuda.empty_cache()
    strong = strong_scaling(args, rank, world, comm)
    kmeans = kmeans_block(args, hbm_peak, peak_src) if (world == 1 and not args.no_kmeans) else None

    if rank == 0:
        head = blocks[workloads[0]]
        head['clocks'] = sampler.summary()
        if len(workloads) > 1:
            head['gpnh'] = blocks['gpnh']
            head['gpu_launches'] += blocks['gpnh']['gpu_launches']
        if gram:
            blocks['aa']['gram_formulation'] = gram
        if strong:
            head['strong_scaling'] = strong
        if kmeans:
            head['kmeans'] = kmeans
        print(json.dumps(head))
        sys.stdout.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        bad = [w for w in workloads if blocks[w].get('parity') and not blocks[w]['parity']['ok']]
        if bad:
            sys.stderr.write('bench.py: PARITY FAILED for %s\n' % ', '.join(bad))
            sys.exit(3)


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
