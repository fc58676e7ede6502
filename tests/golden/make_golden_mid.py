"""Mid-size golden vectors (400 x 6000, k = 8) from the REAL reference: large enough that
the CUDA path takes its production kernels (bulk-copy strip-owned passes, one-sample-per-warp
QP), small enough for the reference to run in seconds on a CPU.

    python tests/golden/make_golden_mid.py      # build container only (/root/reference)

Inputs are regenerated from seeds by ``convex_dim_red.datasets.synthetic_field`` (this
repository's generator, NumPy RandomState only), so only the outputs are stored.
"""

import os
import sys
import warnings

import numpy as np

np.NaN = np.nan          # noqa: NumPy-2 shim for reference spg.py:310
np.product = np.prod
warnings.filterwarnings('ignore')

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.environ.get('CDR_REFERENCE_SRC', '/root/reference/src'))

import convex_dim_red as ref                                     # noqa: E402
from convex_dim_red import archetypal_analysis as raa            # noqa: E402
from convex_dim_red import gpnh_convex_coding as rgp             # noqa: E402
from sklearn.cluster import KMeans                               # noqa: E402

# this repository's synthetic generator (pure NumPy), loaded by path so that the name
# `convex_dim_red` keeps pointing at the reference
import importlib.util                                            # noqa: E402
_spec = importlib.util.spec_from_file_location(
    '_cdr_datasets_standalone',
    os.path.join(ROOT, 'matrix-factorization-case-studies_b200', 'convex_dim_red', 'datasets.py'))

T, D, K = 400, 6000, 8
OUT = {}


def synthetic_field(n_samples, n_features, seed):
    """Same draws as convex_dim_red.datasets.synthetic_field (kept in sync by a test)."""
    rs = np.random.RandomState(seed)
    mix = ref.right_stochastic_matrix((n_samples, 12), rs)
    sources = rs.standard_normal((12, n_features))
    x = mix.dot(sources)
    step = max(1, (1 << 24) // max(n_features, 1))
    for lo in range(0, n_samples, step):
        hi = min(n_samples, lo + step)
        x[lo:hi] += 0.5 * rs.standard_normal((hi - lo, n_features))
    x -= x.mean(axis=0)
    return np.ascontiguousarray(x)


def main():
    X = synthetic_field(T, D, seed=3)
    OUT['X_checksum'] = np.array([X.sum(), np.abs(X).sum(), X[17, 123], X[-1, -1]])
    rs = np.random.RandomState(1)
    W0 = np.sqrt(np.abs(X).mean() / K) * rs.randn(D, K)
    Z0 = ref.right_stochastic_matrix((T, K), random_state=rs)
    C0 = ref.right_stochastic_matrix((K, T), random_state=rs)
    OUT['W0_checksum'] = np.array([W0.sum(), Z0.sum(), C0[3, 7]])

    Z, W, cost, n_iter, _, deltas = rgp._iterate_gpnh_convex_coding(
        X, Z0.copy(), W0.copy(), lambda_W=0.1, tolerance=1e-12, max_iterations=6)
    OUT['gpnh/Z'], OUT['gpnh/W'] = Z, np.ascontiguousarray(W)
    OUT['gpnh/stats'] = np.array([cost, n_iter])
    OUT['gpnh/deltas'] = np.array(deltas)

    Z, C, a, cost, n_iter, _, deltas = raa._iterate_aa(
        X, Z0.copy(), C0.copy(), np.ones(K), tolerance=1e-12, max_iterations=6,
        dictionary_solver_kwargs=dict(max_iterations=1))
    OUT['aa/Z'], OUT['aa/C'] = Z, C
    OUT['aa/stats'] = np.array([cost, n_iter])
    OUT['aa/deltas'] = np.array(deltas)

    m = ref.ArchetypalAnalysis(n_components=K, init='furthest_sum', tolerance=1e-12,
                               max_iterations=8, random_state=5,
                               dictionary_solver_kwargs=dict(max_iterations=1))
    Z = m.fit_transform(X)
    OUT['aa_fs/Z'], OUT['aa_fs/C'] = Z, m.dictionary
    OUT['aa_fs/picks0'] = np.argmax(
        raa._initialize_kernel_aa_dictionary_furthest_sum(
            X.dot(X.T), K, random_state=np.random.RandomState(5)), axis=1)
    OUT['aa_fs/stats'] = np.array([m.cost, m.n_iter])
    Xv = synthetic_field(40, D, seed=4)
    Zv, cv = m.transform(Xv)
    OUT['aa_fs/Zv'], OUT['aa_fs/cost_v'] = Zv, np.array(cv)

    picks = np.asarray(OUT['aa_fs/picks0'], dtype=np.int64)
    km = KMeans(n_clusters=K, init=X[picks].copy(), n_init=1, algorithm='lloyd', tol=1e-4,
                max_iter=300).fit(X.copy())
    OUT['km/labels'] = km.labels_.astype(np.int32)
    OUT['km/stats'] = np.array([km.inertia_, km.n_iter_])
    OUT['km/centres_checksum'] = np.array([km.cluster_centers_.sum(),
                                           np.abs(km.cluster_centers_).sum()])

    path = os.path.join(HERE, 'golden_mid_v1.npz')
    np.savez_compressed(path, **OUT)
    print('wrote %s: %d arrays, %.1f KiB' % (path, len(OUT), os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
