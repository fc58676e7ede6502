#!/usr/bin/env python
"""Headline benchmark: GPNH / AA alternating iterations per second at HadISST shape.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload gpnh|aa] [--impl b200|reference]

A "step" is one outer alternating iteration (dictionary update + weights update +
cost bookkeeping) of the reference's loop (`_iterate_gpnh_convex_coding`,
gpnh_convex_coding.py:282-402, or `_iterate_aa`, archetypal_analysis.py:534-670) on a
synthetic anomaly matrix of HadISST shape (1620 training months x 44 000 ocean
cells, fp64, k = 8), BASELINE.json configs[1] (GPNH, default) / configs[0] (AA).

Printed JSON (one line, rank 0):
  value    whole-job outer iterations / second with X resident in HBM, device timed
  e2e      the same through the public NumPy-in / NumPy-out call, host buffers, H2D of
           X and D2H of the factors inside the timed region
  roofline the dominant kernel (the streaming pass over X) against the measured HBM peak
  cpu_baseline  the CPU oracle port of the reference loop timed on this box's host cores
With N > 1 the sample axis is sharded (weak scaling: every rank owns a full
1620-row slab; the job is a fit of N x 1620 samples).
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

import numpy as np   # noqa: E402

T_ROWS, N_FEATURES, N_COMPONENTS = 1620, 44000, 8
LAMBDA_W = 0.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--workload', choices=('gpnh', 'aa'), default='gpnh')
    ap.add_argument('--impl', choices=('b200', 'reference'), default='b200')
    ap.add_argument('--rows', type=int, default=T_ROWS)
    ap.add_argument('--features', type=int, default=N_FEATURES)
    ap.add_argument('--components', type=int, default=N_COMPONENTS)
    ap.add_argument('--formulation', choices=('stream', 'gram'), default='stream',
                    help="AA only: 'gram' builds K = X X' once and iterates on K (opt-in "
                         "algorithmic variant; the headline numbers use 'stream')")
    ap.add_argument('--cpu-steps', type=int, default=4,
                    help='outer iterations of the CPU baseline sample (0 disables it)')
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


def make_problem(args, rank=0, world=1):
    """Host-generated inputs, identical for the CUDA path and the CPU arm
    (BASELINE.md section 3).  Each rank owns a full slab of `rows` samples."""
    from convex_dim_red.datasets import synthetic_field
    from convex_dim_red.stochastic_matrices import right_stochastic_matrix
    T, d, k = args.rows, args.features, args.components
    X = synthetic_field(T, d, seed=rank)
    rs = np.random.RandomState(1000 + rank)
    if args.workload == 'gpnh':
        # the dictionary is replicated: every rank draws the same one
        W0 = np.sqrt(0.4 / k) * np.random.RandomState(0).randn(d, k)
        Z0 = right_stochastic_matrix((T, k), rs)
        return X, Z0, W0
    # the dictionary (k x total samples) is replicated; the weights are this rank's rows
    C0 = right_stochastic_matrix((k, T * world), np.random.RandomState(7))
    Z0 = right_stochastic_matrix((T, k), rs)
    return X, Z0, C0


def cpu_steps(args, X, Z0, F0, n_steps):
    """The reference loop (oracle port: NumPy/BLAS contractions + C per-sample QPs,
    same pass structure as the reference) for n_steps outer iterations."""
    from oracle import convex_oracle as orc
    trace = float(np.sum(X * X))
    times = []
    if args.workload == 'gpnh':
        out = orc.iterate_gpnh(X, Z0.copy(), F0.copy(), lambda_W=LAMBDA_W, tolerance=0.0,
                               max_iterations=n_steps, trace_XtX=trace,
                               require_monotonic_cost_decrease=False, iter_times_out=times)
        cost = out[2]
    else:
        out = orc.iterate_aa(X, Z0.copy(), F0.copy(), np.ones(F0.shape[0]), tolerance=0.0,
                             max_iterations=n_steps, trace_XXt=trace,
                             dictionary_solver_kwargs=dict(max_iterations=1),
                             require_monotonic_cost_decrease=False, iter_times_out=times)
        cost = out[3]
    return times, cost


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p['num_threads'] for p in threadpool_info() if p.get('user_api') == 'blas']
        return max(n) if n else os.cpu_count()
    except Exception:
        return os.cpu_count()


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    X, Z0, F0 = make_problem(args)
    # run `warmup + steps` outer iterations from the same start as the CUDA arm and
    # time the last `steps` of them (per-iteration wall times from the loop itself)
    times, cost = cpu_steps(args, X, Z0, F0, args.warmup + args.steps)
    timed = max(sum(times[args.warmup:]), 1e-9)
    value = args.steps / timed
    cores = blas_threads()
    line = {
        'impl': 'reference', 'metric': metric_name(args), 'value': value, 'unit': 'iterations/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * timed / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': config_dict(args, 1),
        'cpu_baseline': {'value': value, 'unit': 'iterations/s', 'cores': cores, 'kind': 'port',
                         'sample': '%d outer iterations after %d untimed, oracle port of the '
                                   'reference loop (NumPy/OpenBLAS passes + C per-sample QPs)'
                                   % (args.steps, args.warmup)},
        'e2e': {'value': value, 'unit': 'iterations/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'final_cost': cost,
    }
    print(json.dumps(line))


def metric_name(args):
    return ('gpnh_outer_iterations_per_sec_hadisst' if args.workload == 'gpnh'
            else 'aa_outer_iterations_per_sec_hadisst')


def config_dict(args, world):
    return {'workload': '%s k=%d on synthetic HadISST-shaped anomalies, %d x %d fp64 per GPU '
                        '(BASELINE.json configs[%d])'
                        % ('GPNH convex coding' if args.workload == 'gpnh' else
                           'archetypal analysis (dictionary SPG max_iterations=1)',
                           args.components, args.rows, args.features,
                           1 if args.workload == 'gpnh' else 0),
            'n_samples_total': args.rows * world, 'n_features': args.features,
            'n_components': args.components, 'lambda_W': LAMBDA_W,
            'formulation': ('streaming (X read from HBM every pass; 2 passes/iter GPNH, 4 AA)'
                            if args.formulation == 'stream' else
                            'gram (K = X X^T built once, 2 passes over the L2-resident K per '
                            'iteration; no HBM roofline applies)'),
            'l2_policy': 'X (570 MB per GPU) is larger than the 126 MB L2; no explicit flush',
            'sharding': 'sample axis, %d rank(s)' % world}


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
    from convex_dim_red import _bench_support as bs

    X, Z0, F0 = make_problem(args, rank, world)
    result = bs.run_benchmark(args, X, Z0, F0, rank, world, ClockSampler(local_rank))
    if rank == 0:
        peak, peak_src = peaks()
        roof = result['roofline']
        if roof is not None:
            roof.update({'bound': 'hbm', 'peak': peak, 'unit': 'GB/s',
                         'frac': roof['achieved'] / peak, 'peak_source': peak_src})
            tpath = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
            if os.path.exists(tpath):
                with open(tpath) as fh:
                    tr = json.load(fh)
                roof['traffic'] = tr['dram_bytes_per_launch'].get(roof['kernel'])
                roof['traffic_source'] = tr['source']
        line = {
            'metric': metric_name(args), 'value': result['value'], 'unit': 'iterations/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': result['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': config_dict(args, world), 'clocks': result['clocks'],
            'e2e': result['e2e'], 'gpu_launches': result['gpu_launches'],
            'roofline': roof, 'kernels': result['kernels'], 'final_cost': result['final_cost'],
            'time_to_converge': result['time_to_converge'],
        }
        if args.cpu_steps > 0 and world == 1:
            times, cost = cpu_steps(args, X, Z0, F0, args.cpu_steps)
            line['cpu_baseline'] = {
                'value': args.cpu_steps / sum(times), 'unit': 'iterations/s',
                'cores': blas_threads(), 'kind': 'port',
                'sample': 'first %d outer iterations of the same workload from the same start '
                          '(oracle port of the reference loop)' % args.cpu_steps,
                'final_cost': cost}
            ttc = line.get('time_to_converge')
            if ttc:
                # not run: iterations of the converged CUDA fit x the CPU time per iteration
                ttc['cpu_estimate_seconds'] = ttc['iterations'] * sum(times) / args.cpu_steps
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
