"""Mid-size (400 x 6000, k = 8) golden vectors generated from the real reference
(tests/golden/make_golden_mid.py): the oracle on CPU, the CUDA path on GPU.  At this size
the CUDA path runs its production kernels (bulk-copy strip-owned passes)."""

import os
import warnings

import numpy as np
import pytest

from conftest import ROOT, Golden

T, D, K = 400, 6000, 8


@pytest.fixture(scope='module')
def mid():
    return Golden(os.path.join(ROOT, 'tests', 'golden', 'golden_mid_v1.npz'))


@pytest.fixture(scope='module')
def problem(mid):
    from convex_dim_red.datasets import synthetic_field
    from convex_dim_red.stochastic_matrices import right_stochastic_matrix
    X = synthetic_field(T, D, seed=3)
    # the fixture generator and this repository's generator must draw the same numbers
    np.testing.assert_allclose([X.sum(), np.abs(X).sum(), X[17, 123], X[-1, -1]],
                               mid['X_checksum'], rtol=1e-12, atol=1e-9)
    rs = np.random.RandomState(1)
    W0 = np.sqrt(np.abs(X).mean() / K) * rs.randn(D, K)
    Z0 = right_stochastic_matrix((T, K), random_state=rs)
    C0 = right_stochastic_matrix((K, T), random_state=rs)
    np.testing.assert_allclose([W0.sum(), Z0.sum(), C0[3, 7]], mid['W0_checksum'], rtol=1e-12)
    return X, Z0, W0, C0


def close(a, b, rtol, atol=0.0):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def _check_gpnh(mid, Z, W, cost, n_iter, deltas):
    gcost, gn = mid['gpnh/stats']
    assert n_iter == int(gn)
    close(cost, gcost, rtol=1e-8)
    close(deltas, mid['gpnh/deltas'], rtol=1e-5, atol=1e-9)
    close(Z, mid['gpnh/Z'], rtol=0, atol=2e-5)
    close(W, mid['gpnh/W'], rtol=0, atol=2e-5)


def _check_aa(mid, Z, C, cost, n_iter, deltas):
    gcost, gn = mid['aa/stats']
    assert n_iter == int(gn)
    close(cost, gcost, rtol=1e-8)
    close(deltas, mid['aa/deltas'], rtol=1e-5, atol=1e-9)
    close(Z, mid['aa/Z'], rtol=0, atol=2e-5)
    close(C, mid['aa/C'], rtol=0, atol=2e-5)


# ---------------------------------------------------------------- oracle (CPU)
def test_oracle_mid_gpnh(mid, problem):
    from oracle import convex_oracle as orc
    X, Z0, W0, C0 = problem
    Z, W, cost, n_iter, _, deltas = orc.iterate_gpnh(X, Z0.copy(), W0.copy(), lambda_W=0.1,
                                                     tolerance=1e-12, max_iterations=6)
    _check_gpnh(mid, Z, W, cost, n_iter, deltas)


def test_oracle_mid_aa(mid, problem):
    from oracle import convex_oracle as orc
    X, Z0, W0, C0 = problem
    Z, C, a, cost, n_iter, _, deltas = orc.iterate_aa(
        X, Z0.copy(), C0.copy(), np.ones(K), tolerance=1e-12, max_iterations=6,
        dictionary_solver_kwargs=dict(max_iterations=1))
    _check_aa(mid, Z, C, cost, n_iter, deltas)


def test_oracle_mid_furthest_sum_and_kmeans(mid, problem):
    from oracle import convex_oracle as orc
    X = problem[0]
    rng = np.random.RandomState(5)
    C, _ = orc.init_kernel_aa(X.dot(X.T), K, 'furthest_sum', rng)
    picks = np.argmax(C, axis=1)
    assert np.array_equal(picks, mid['aa_fs/picks0'])
    labels, centres, inertia, n_iter = orc.kmeans_lloyd(X, X[picks], tol=1e-4, max_iter=300)
    assert np.array_equal(labels, mid['km/labels'])
    close([inertia, n_iter], mid['km/stats'], rtol=1e-10)
    close([centres.sum(), np.abs(centres).sum()], mid['km/centres_checksum'], rtol=1e-9)


# ---------------------------------------------------------------- CUDA path (GPU)
def _need_gpu():
    torch = pytest.importorskip('torch')
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')


@pytest.mark.gpu
def test_cuda_mid_gpnh(mid, problem):
    _need_gpu()
    from convex_dim_red import gpnh_convex_coding as gp
    X, Z0, W0, C0 = problem
    Z, W, cost, n_iter, _, deltas = gp._iterate_gpnh_convex_coding(
        X, Z0.copy(), W0.copy(), lambda_W=0.1, tolerance=1e-12, max_iterations=6)
    _check_gpnh(mid, Z, W, cost, n_iter, deltas)


@pytest.mark.gpu
def test_cuda_mid_aa(mid, problem):
    _need_gpu()
    from convex_dim_red import archetypal_analysis as aa
    X, Z0, W0, C0 = problem
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        Z, C, a, cost, n_iter, _, deltas = aa._iterate_aa(
            X, Z0.copy(), C0.copy(), np.ones(K), tolerance=1e-12, max_iterations=6,
            dictionary_solver_kwargs=dict(max_iterations=1))
        _check_aa(mid, Z, C, cost, n_iter, deltas)
        # the opt-in Gram formulation follows the same trajectory
        Zg, Cg, _, costg, ng, _, deltas_g = aa._iterate_aa(
            X, Z0.copy(), C0.copy(), np.ones(K), tolerance=1e-12, max_iterations=6,
            dictionary_solver_kwargs=dict(max_iterations=1), formulation='gram')
        _check_aa(mid, Zg, Cg, costg, ng, deltas_g)


@pytest.mark.gpu
def test_cuda_mid_estimator_furthest_sum_and_kmeans(mid, problem):
    _need_gpu()
    import convex_dim_red as cdr
    from convex_dim_red.datasets import synthetic_field
    from convex_dim_red.kmeans import kmeans_lloyd
    X = problem[0]
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        m = cdr.ArchetypalAnalysis(n_components=K, init='furthest_sum', tolerance=1e-12,
                                   max_iterations=8, random_state=5,
                                   dictionary_solver_kwargs=dict(max_iterations=1))
        Z = m.fit_transform(X)
    gcost, gn = mid['aa_fs/stats']
    assert m.n_iter == int(gn)
    close(m.cost, gcost, rtol=1e-8)
    close(Z, mid['aa_fs/Z'], rtol=0, atol=2e-5)
    close(m.dictionary, mid['aa_fs/C'], rtol=0, atol=2e-5)
    Zv, cv = m.transform(synthetic_field(40, D, seed=4))
    close(Zv, mid['aa_fs/Zv'], rtol=0, atol=2e-5)
    close(cv, float(mid['aa_fs/cost_v']), rtol=1e-7)
    picks = np.asarray(mid['aa_fs/picks0'], dtype=np.int64)
    labels, centres, inertia, n_iter = kmeans_lloyd(X, X[picks], tol=1e-4, max_iter=300)
    assert np.array_equal(labels, mid['km/labels'])
    close([inertia, n_iter], mid['km/stats'], rtol=1e-10)
    close([centres.sum(), np.abs(centres).sum()], mid['km/centres_checksum'], rtol=1e-9)
