#!/usr/bin/env python
"""The streaming contractions and the batched QP at the stress shape (BASELINE.json configs[4]:
k = 64, 18 000 x 44 000), each timed alone with CUDA events, against the fp64 tensor-pipe
(DMMA) peak measured in the same process.

    python profiles/bench_stress_kernels.py [T] [k]      -> JSON lines
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'matrix-factorization-case-studies_b200'))

import numpy as np   # noqa: E402
import torch         # noqa: E402

from convex_dim_red import _backend as be   # noqa: E402
import bench_harness as bh                  # noqa: E402


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 18000
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    d = 44000
    peak = be.dmma_peak_tflops()
    g = torch.Generator(device='cuda').manual_seed(0)
    X = torch.randn((T, be.round_up(d)), dtype=torch.float64, device='cuda', generator=g)
    X[:, d:] = 0
    ldt = be.round_up(T)
    Z = torch.rand((T, k), dtype=torch.float64, device='cuda', generator=g)
    Z /= Z.sum(dim=1, keepdim=True)
    D = torch.randn((k, ldt), dtype=torch.float64, device='cuda', generator=g)
    D[:, T:] = 0
    M = torch.randn((k, be.round_up(d)), dtype=torch.float64, device='cuda', generator=g)
    M[:, d:] = 0
    E = torch.eye(k, dtype=torch.float64, device='cuda') * 0.5
    out_kd = be.zeros(k, be.round_up(d))
    out_kt = be.zeros(k, ldt)
    ws = be.Workspace(T, d, k)
    flops = 2.0 * k * T * d

    def emit(what, ms, **kw):
        line = {'what': what, 'T': T, 'd': d, 'k': k, 'ms': ms, 'tflops': flops / (ms * 1e-3) / 1e12,
                'dmma_peak_tflops': peak}
        line['frac_of_dmma_peak'] = line['tflops'] / peak
        line.update(kw)
        print(json.dumps(line), flush=True)

    for name, env in (('gemm64', '0'), ('round-1 direct-load', '1')):
        os.environ['CDR_DISABLE_GEMM64'] = env
        t = bh.time_launches(lambda: be.reduce_samples(Z, 1, k, X, T, d, k, out_kd, ws, E=E), reps=4)
        emit('reduce_samples  W^T = E Z^T X   [%s]' % name, t)
        t = bh.time_launches(lambda: be.reduce_samples(D, ldt, 1, X, T, d, k, out_kd, ws), reps=4)
        emit('reduce_samples  D X             [%s]' % name, t)
        t = bh.time_launches(lambda: be.reduce_features(M, X, T, d, k, out_kt, ws), reps=4)
        emit('reduce_features M X^T           [%s]' % name, t)
    os.environ['CDR_DISABLE_GEMM64'] = '0'
    # numerical cross-check of the two paths
    a = be.zeros(k, be.round_up(d)); b = be.zeros(k, be.round_up(d))
    be.reduce_samples(Z, 1, k, X, T, d, k, a, ws, E=E)
    os.environ['CDR_DISABLE_GEMM64'] = '1'
    be.reduce_samples(Z, 1, k, X, T, d, k, b, ws, E=E)
    os.environ['CDR_DISABLE_GEMM64'] = '0'
    print(json.dumps({'what': 'max rel diff gemm64 vs direct-load (samples)',
                      'value': float((a - b).abs().max() / b.abs().max())}))
    # the per-sample QPs at this shape (A well conditioned: a few dozen SPG iterations)
    A = M[:, :d] @ M[:, :d].T / d
    B = (X[:, :d] @ M[:, :d].T / d).contiguous()          # T x k
    params = be.make_spg_params({})
    zsave = Z.clone()
    n_it = torch.zeros(T, dtype=torch.int32, device='cuda')
    t = bh.time_launches(lambda: (Z.copy_(zsave), be.quad_simplex_spg_batched(
        A.contiguous(), None, B, k, 1, Z, T, k, params, n_iter=n_it)), reps=2)
    print(json.dumps({'what': 'quad_simplex_spg_batched', 'T': T, 'k': k, 'ms': t,
                      'mean_spg_iterations': float(n_it.double().mean().item()),
                      'max_spg_iterations': int(n_it.max().item())}))


if __name__ == '__main__':
    main()
