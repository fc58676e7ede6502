"""Synthetic inputs of the shapes the reference's drivers feed the hot path.

There is no network on the build or GPU boxes, so HadISST / JRA-55 are replaced
by seeded surrogates with the same shapes: a noisy
convex mixture ``X = Z0 A0 + sigma E`` with the column mean removed
("anomalies"), which makes the solvers do real work (iid noise converges in two
k-means iterations).
"""

import numpy as np

from .stochastic_matrices import right_stochastic_matrix

# name -> (n_samples, n_features): training rows of bin/run_hadisst_aa.py:205
# (0.9 * 1800 months), JRA-55 hgt500 grid, its 167-EOF reduction, 10x stress.
SHAPES = {
    'hadisst': (1620, 44000),
    'jra55': (700, 41800),
    'jra55_pca': (700, 167),
    'hadisst_x10': (18000, 44000),
}


def synthetic_field(n_samples, n_features, seed=0, n_sources=12, sigma=0.5,
                    dtype=np.float64):
    """Return a C-contiguous (n_samples, n_features) anomaly matrix."""
    rs = np.random.RandomState(seed)
    mix = right_stochastic_matrix((n_samples, n_sources), rs)
    sources = rs.standard_normal((n_sources, n_features))
    x = mix.dot(sources)
    # add the noise in row blocks so the 18000 x 44000 case needs no second
    # full-size temporary
    step = max(1, (1 << 24) // max(n_features, 1))
    for lo in range(0, n_samples, step):
        hi = min(n_samples, lo + step)
        x[lo:hi] += sigma * rs.standard_normal((hi - lo, n_features))
    x -= x.mean(axis=0)
    return np.ascontiguousarray(x, dtype=dtype)


def synthetic_named(name, seed=0, scale=1.0):
    """Surrogate for one of ``SHAPES``; ``scale`` shrinks both axes (tests)."""
    n_samples, n_features = SHAPES[name]
    return synthetic_field(max(8, int(round(n_samples * scale))),
                           max(8, int(round(n_features * scale))), seed=seed)
