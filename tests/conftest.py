"""Shared pytest configuration.

* registers the ``gpu`` marker (tests that need a B200; the driver runs
  ``-m "not gpu"`` on a CPU-only box and ``-m gpu`` on the GPU box);
* puts the product package (``matrix-factorization-case-studies_b200/``, which
  holds the drop-in ``convex_dim_red`` package) and the repo root (for
  ``oracle``) on ``sys.path``;
* loads the golden fixtures generated from the real reference by
  ``tests/golden/make_golden.py``.
"""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, 'matrix-factorization-case-studies_b200')
for p in (ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (B200)')


class Golden:
    def __init__(self, path):
        self._z = np.load(path, allow_pickle=False)

    def __getitem__(self, key):
        return self._z[key]

    def scalar(self, key):
        return float(self._z[key])

    def keys(self):
        return list(self._z.keys())


@pytest.fixture(scope='session')
def golden():
    return Golden(os.path.join(ROOT, 'tests', 'golden', 'golden_v1.npz'))


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:      # pragma: no cover
        return False
