#!/usr/bin/env python
"""Micro-benchmark of the two streaming passes (CUDA events, inputs larger than L2).

    python profiles/bench_passes.py [T d k]      # default 1620 44000 8
Prints one JSON line per kernel variant; CDR_DISABLE_TMA=1 selects the direct-load kernels.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'matrix-factorization-case-studies_b200'))

import numpy as np   # noqa: E402
import torch         # noqa: E402
from convex_dim_red import _backend as be   # noqa: E402


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    T, d, k = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (1620, 44000, 8)
    g = torch.Generator(device='cuda').manual_seed(0)
    ld = be.round_up(d)
    X = torch.zeros((T, ld), dtype=torch.float64, device='cuda')
    X[:, :d] = torch.randn((T, d), dtype=torch.float64, device='cuda', generator=g)
    Z = torch.rand((T, k), dtype=torch.float64, device='cuda', generator=g)
    L = torch.zeros((k, be.round_up(T)), dtype=torch.float64, device='cuda')
    L[:, :T] = torch.randn((k, T), dtype=torch.float64, device='cuda', generator=g)
    E = torch.randn((k, k), dtype=torch.float64, device='cuda', generator=g)
    M = torch.zeros((k, ld), dtype=torch.float64, device='cuda')
    M[:, :d] = torch.randn((k, d), dtype=torch.float64, device='cuda', generator=g)
    out_kd = be.zeros(k, ld)
    out_kt = be.zeros(k, be.round_up(T))
    ws = be.Workspace(T, d, k)
    nbytes = 8.0 * T * d
    variants = {
        'reduce_samples(Z^T X, E)': lambda: be.reduce_samples(Z, 1, k, X, T, d, k, out_kd, ws, E=E),
        'reduce_samples(L X)': lambda: be.reduce_samples(L, L.stride(0), 1, X, T, d, k, out_kd, ws),
        'reduce_features(M X^T)': lambda: be.reduce_features(M, X, T, d, k, out_kt, ws),
    }
    for name, fn in variants.items():
        ms = timeit(fn)
        print(json.dumps({'kernel': name, 'T': T, 'd': d, 'k': k, 'ms': round(ms, 5),
                          'GBps': round(nbytes / ms / 1e6, 1),
                          'tma': os.environ.get('CDR_DISABLE_TMA', '0') != '1'}))


if __name__ == '__main__':
    main()
