#!/usr/bin/env python
"""Gram matrix K = X X' (archetypal_analysis.py:1032): SYRK kernel vs the round-1 slab path,
CUDA-event timed, with the fp64 tensor (DMMA) peak measured in the same process.

    python profiles/bench_gram.py            -> one JSON line per shape
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
    peak = be.dmma_peak_tflops() if hasattr(be, 'dmma_peak_tflops') else None
    for T, d in ((1620, 44000), (700, 41800), (4096, 44000)):
        g = torch.Generator(device='cuda').manual_seed(0)
        X = torch.randn((T, be.round_up(d)), dtype=torch.float64, device='cuda', generator=g)
        X[:, d:] = 0
        t_syrk = bh.time_launches(lambda: be.gram(X, T, d), reps=5)
        t_slab = bh.time_launches(lambda: be.gram_slabs(X, T, d), reps=2)
        K1, K2 = be.gram(X, T, d), be.gram_slabs(X, T, d)
        err = float((K1 - K2).abs().max() / K2.abs().max())
        line = {'what': 'gram K = X X^T', 'T': T, 'd': d, 'syrk_ms': t_syrk, 'slab_ms': t_slab,
                'syrk_tflops_algorithmic': T * T * d / (t_syrk * 1e-3) / 1e12,
                'syrk_tflops_executed': (T + 127) // 128 * ((T + 127) // 128 + 1) / 2 * 128 * 128 * 2.0 * d
                / (t_syrk * 1e-3) / 1e12,
                'slab_tflops_executed': 2.0 * T * T * d / (t_slab * 1e-3) / 1e12,
                'max_rel_diff_vs_slab': err, 'dmma_peak_tflops': peak}
        if peak:
            line['syrk_frac_of_dmma_peak'] = line['syrk_tflops_executed'] / peak
        print(json.dumps(line), flush=True)


if __name__ == '__main__':
    main()
