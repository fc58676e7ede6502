#!/usr/bin/env python
"""Short timing / sanity runs of the other BASELINE.json configs (3, 4, 5) on one GPU.

    python profiles/bench_configs.py [3] [4] [5] [--cpu]
Prints one JSON line per measurement.  `--cpu` adds the oracle (CPU port) timing where it
finishes in seconds.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'matrix-factorization-case-studies_b200'))

import numpy as np   # noqa: E402
import torch         # noqa: E402
import warnings      # noqa: E402
warnings.simplefilter('ignore')

from convex_dim_red import _backend as be                                   # noqa: E402
from convex_dim_red import archetypal_analysis as aa                        # noqa: E402
from convex_dim_red import gpnh_convex_coding as gp                         # noqa: E402
from convex_dim_red.datasets import synthetic_field                         # noqa: E402
from convex_dim_red.kmeans import furthest_sum_centres, kmeans_lloyd        # noqa: E402


def emit(**kw):
    print(json.dumps(kw), flush=True)


def wall(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return time.perf_counter() - t0, out


def config3(cpu):
    """k-means k=8 with FurthestSum init on a JRA-55 hgt500-shaped field (700 x 41 800)."""
    X = synthetic_field(700, 41800, seed=0)
    start = np.random.RandomState(0).randint(700)
    t_fs, picks = wall(lambda: furthest_sum_centres(X, 8, start, 10))
    t_km, (labels, centres, inertia, n_iter) = wall(lambda: kmeans_lloyd(X, X[picks], tol=1e-4, max_iter=10000))
    t_km2, _ = wall(lambda: kmeans_lloyd(X, X[picks], tol=1e-4, max_iter=10000))
    emit(config=3, what='furthest_sum init (Gram + dissimilarities + picks)', seconds=t_fs, picks=[int(p) for p in picks])
    emit(config=3, what='kmeans_lloyd incl. upload', seconds=t_km2, first_call_seconds=t_km, n_iter=int(n_iter),
         ms_per_lloyd_iteration=1e3 * t_km2 / n_iter, inertia=inertia)
    if cpu:
        from sklearn.cluster import KMeans
        from oracle import convex_oracle as orc
        t0 = time.perf_counter()
        K = X.dot(X.T)
        ref_picks = orc.furthest_sum(np.nan_to_num(orc.dissimilarity_from_kernel(K)), 8, start, [], 10)
        t_fs_cpu = time.perf_counter() - t0
        t0 = time.perf_counter()
        km = KMeans(n_clusters=8, init=X[picks].copy(), n_init=1, algorithm='lloyd', tol=1e-4, max_iter=10000).fit(X.copy())
        t_cpu = time.perf_counter() - t0
        emit(config=3, what='cpu: sklearn KMeans same init', seconds=t_cpu, n_iter=int(km.n_iter_),
             labels_equal=bool(np.array_equal(km.labels_, labels)), inertia=float(km.inertia_),
             furthest_sum_cpu_seconds=t_fs_cpu, picks_equal=bool(np.array_equal(ref_picks, picks)))


def config4(cpu):
    """PCA reduction of a JRA-55-shaped field (700 x 41 800 -> 167 components) followed by AA
    k = 4..20 on the scores, both solvers 1 iteration,
    rel_delta_f 1e-6 (bin/run_jra55_pca_aa.py:119-133)."""
    from convex_dim_red.pca import PCA
    field = synthetic_field(700, 41800, seed=0)
    PCA(n_components=167).fit_transform(field)                     # warm-up (library load, eigh)
    t_pca, X = wall(lambda: PCA(n_components=167).fit_transform(field))
    emit(config=4, what='PCA 700 x 41800 -> 167 components (upload, centring, SYRK Gram, eigh, scores)',
         seconds=t_pca, shape=list(X.shape))
    for k in (4, 8, 20):
        def fit():
            m = aa.ArchetypalAnalysis(n_components=k, init='random', tolerance=1e-6, max_iterations=10000,
                                      random_state=0, stopping_criterion='rel_delta_f',
                                      dictionary_solver_kwargs=dict(max_iterations=1),
                                      weights_solver_kwargs=dict(max_iterations=1))
            m.fit_transform(X)
            return m
        fit()
        t, m = wall(fit)
        emit(config=4, k=k, what='AA fit to convergence', seconds=t, n_iter=int(m.n_iter) + 1, cost=float(m.cost),
             ms_per_iteration=1e3 * t / (m.n_iter + 1))
        if cpu and k <= 8:
            from oracle import convex_oracle as orc
            rng = np.random.RandomState(0)
            C, Z = orc.init_kernel_aa(X.dot(X.T), k, 'random', rng)
            t0 = time.perf_counter()
            out = orc.iterate_aa(X, Z, C, np.ones(k), tolerance=1e-6, max_iterations=10000,
                                 stopping_criterion='rel_delta_f', dictionary_solver_kwargs=dict(max_iterations=1),
                                 weights_solver_kwargs=dict(max_iterations=1))
            tc = time.perf_counter() - t0
            emit(config=4, k=k, what='cpu oracle same fit', seconds=tc, n_iter=int(out[4]) + 1, cost=float(out[3]))


def config5(cpu):
    """Stress: AA and GPNH k = 64 on an 18 000 x 44 000 matrix (6.3 GB), a few iterations."""
    T, d, k = 18000, 44000, 64
    g = torch.Generator(device='cuda').manual_seed(0)
    Xd = torch.randn((T, be.round_up(d)), dtype=torch.float64, device='cuda', generator=g)
    rs = np.random.RandomState(0)
    Z0 = rs.uniform(size=(T, k)); Z0 /= Z0.sum(axis=1)[:, None]
    W0 = 0.2 * rs.randn(d, k)
    C0 = rs.uniform(size=(k, T)); C0 /= C0.sum(axis=1)[:, None]

    class Shape:                      # the engines only read .shape of the host array when a device copy is given
        shape = (T, d)
    from convex_dim_red.gpnh_convex_coding import _GpnhEngine
    from convex_dim_red.archetypal_analysis import _AaEngine
    for name, make in (('gpnh', lambda: _GpnhEngine(Shape, Z0, W0, tolerance=0.0, max_iterations=100,
                                                   require_monotonic_cost_decrease=False, X_device=Xd)),
                       ('aa', lambda: _AaEngine(Shape, Z0, C0, np.ones(k), 'feature', tolerance=0.0,
                                                max_iterations=100, require_monotonic_cost_decrease=False,
                                                dictionary_solver_kwargs=dict(max_iterations=1), data_device=Xd))):
        eng = make()
        eng.initial_cost()
        eng.iteration()
        t, _ = wall(lambda: [eng.iteration() for _ in range(3)])
        st = eng.state.read()
        emit(config=5, what=name + ' outer iteration, k=64, T=18000', ms_per_iteration=1e3 * t / 3, cost=st.cost,
             n_iter=st.n_iter, flops_per_iteration=(4 if name == 'gpnh' else 8) * k * T * d,
             tflops=(4 if name == 'gpnh' else 8) * k * T * d / (t / 3) / 1e12)
        del eng


if __name__ == '__main__':
    which = [a for a in sys.argv[1:] if a.isdigit()] or ['3', '4', '5']
    cpu = '--cpu' in sys.argv
    for c in which:
        {'3': config3, '4': config4, '5': config5}[c](cpu)
