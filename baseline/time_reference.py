#!/usr/bin/env python
"""Times the REAL reference (azedarach/matrix-factorization-case-studies, unmodified) on the
host cores of the box it runs on.  Measurement infrastructure only.

The reference is installed, git-ignored, under ``baseline/_ref`` with

    cp -r /root/reference /tmp/refcopy && sed -i /setup_requires/d /tmp/refcopy/setup.py
    python -m pip install --no-index --no-build-isolation --no-deps \
        --target baseline/_ref /tmp/refcopy

(`setup_requires=['pytest-runner', ...]` cannot be resolved offline; the library sources are
untouched).  It must run in its own process: the reference package and the B200 package are
both called ``convex_dim_red``.  One NumPy-2 shim is needed before import: ``np.NaN``
(spg.py:310).  Prints one JSON line.

    python baseline/time_reference.py --workload aa|gpnh [--rows T] [--features d]
        [--components k] [--iterations n] [--threads N]

AA runs the reference's own ``_iterate_aa`` (archetypal_analysis.py:534-670).  GPNH at
d = 44 000 cannot call ``_iterate_gpnh_convex_coding`` as it stands -- its first statement
group forms the d x d product ``X.T.dot(X)`` only for its trace (gpnh_convex_coding.py:302;
15.5 GB, OpenBLAS dsyrk crashes on it) -- so its loop body (:339-399) is executed here
through the reference's own step functions with ``trace_XtX = ||X||_F^2`` supplied.
"""

import argparse
import json
import os
import sys
import time


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', choices=('aa', 'gpnh'), default='aa')
    ap.add_argument('--rows', type=int, default=1620)
    ap.add_argument('--features', type=int, default=44000)
    ap.add_argument('--components', type=int, default=8)
    ap.add_argument('--iterations', type=int, default=2)
    ap.add_argument('--threads', type=int, default=0, help='0 = all host cores')
    return ap.parse_args()


ARGS = parse_args()
NTHREADS = ARGS.threads if ARGS.threads > 0 else (os.cpu_count() or 1)
for _v in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS', 'NUMBA_NUM_THREADS'):
    os.environ[_v] = str(NTHREADS)

import numpy as np   # noqa: E402

np.NaN = np.nan          # spg.py:310 (removed in NumPy 2.0)
np.product = np.prod

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '_ref'))

import convex_dim_red                                          # noqa: E402  (the reference)
from convex_dim_red import archetypal_analysis as raa          # noqa: E402
from convex_dim_red import gpnh_convex_coding as rgp           # noqa: E402

assert os.path.realpath(convex_dim_red.__file__).startswith(os.path.realpath(HERE)), \
    'imported the wrong convex_dim_red: %s' % convex_dim_red.__file__


def stochastic_rows(shape, rs):
    a = rs.uniform(size=shape)
    return a / a.sum(axis=1)[:, np.newaxis]


def synthetic_field(n_samples, n_features, seed=0, n_sources=12, sigma=0.5):
    """Same generator and draw order as convex_dim_red.datasets.synthetic_field of the B200
    package (bench.py's inputs)."""
    rs = np.random.RandomState(seed)
    mix = stochastic_rows((n_samples, n_sources), rs)
    sources = rs.standard_normal((n_sources, n_features))
    x = mix.dot(sources)
    step = max(1, (1 << 24) // max(n_features, 1))
    for lo in range(0, n_samples, step):
        hi = min(n_samples, lo + step)
        x[lo:hi] += sigma * rs.standard_normal((hi - lo, n_features))
    x -= x.mean(axis=0)
    return np.ascontiguousarray(x)


def factors(workload, T, d, k):
    Z0 = stochastic_rows((T, k), np.random.RandomState(1000))
    if workload == 'gpnh':
        return Z0, np.sqrt(0.4 / k) * np.random.RandomState(0).randn(d, k)
    return Z0, stochastic_rows((k, T), np.random.RandomState(7))


def run_aa(X, Z0, C0, n_iter):
    out = raa._iterate_aa(X, Z0.copy(), C0.copy(), np.ones(C0.shape[0]), tolerance=0.0,
                          max_iterations=n_iter, require_monotonic_cost_decrease=False,
                          dictionary_solver_kwargs=dict(max_iterations=1))
    return float(out[3]), float(out[5])


def run_gpnh(X, Z0, W0, n_iter):
    """gpnh_convex_coding.py:339-399 through the reference's own step functions."""
    T, k = Z0.shape
    d = X.shape[1]
    Z, W = Z0.copy(), W0.copy()
    ZtZ = Z.T.dot(Z)
    GW = 4.0 / (d * k * (k - 1)) * (k * np.eye(k) - 1)
    trace = float(np.einsum('ij,ij->', X, X))
    times = []
    cost = None
    for _ in range(n_iter):
        t0 = time.perf_counter()
        W = rgp._update_gpnh_dictionary(X, Z, ZtZ, GW, lambda_W=0)
        WtXt = W.T.dot(X.T)
        WtW = W.T.dot(W)
        cost = 0.5 * (trace - 2 * WtXt.dot(Z).trace() + ZtZ.dot(WtW).trace()) / T
        Z = rgp._update_gpnh_weights(X, Z, W)
        ZtZ = Z.T.dot(Z)
        cost = 0.5 * (trace - 2 * WtXt.dot(Z).trace() + ZtZ.dot(WtW).trace()) / T
        times.append(time.perf_counter() - t0)
    return float(cost), float(np.mean(times))


def main():
    a = ARGS
    run = run_aa if a.workload == 'aa' else run_gpnh
    # warm-up on a small problem: pays the Numba JIT of the same signatures
    t0 = time.perf_counter()
    Xs = synthetic_field(64, 256, seed=1)
    Zs, Fs = factors(a.workload, 64, 256, a.components)
    run(Xs, Zs, Fs, 1)
    jit_s = time.perf_counter() - t0
    X = synthetic_field(a.rows, a.features, seed=0)
    Z0, F0 = factors(a.workload, a.rows, a.features, a.components)
    t0 = time.perf_counter()
    cost, per_iter = run(X, Z0, F0, a.iterations)
    wall = time.perf_counter() - t0
    print(json.dumps({
        'impl': 'reference-numba', 'workload': a.workload, 'value': 1.0 / per_iter,
        'unit': 'iterations/s', 'seconds_per_iteration': per_iter, 'iterations': a.iterations,
        'wall_seconds': wall, 'jit_warmup_seconds': jit_s, 'threads': NTHREADS,
        'host_cpus': os.cpu_count(), 'final_cost': cost,
        'shape': [a.rows, a.features, a.components],
        'what': ('reference _iterate_aa, avg_time_per_iter' if a.workload == 'aa' else
                 'reference loop body gpnh_convex_coding.py:339-399 via _update_gpnh_dictionary / '
                 '_update_gpnh_weights (the d x d trace at :302 replaced by ||X||_F^2)')}))


if __name__ == '__main__':
    main()
