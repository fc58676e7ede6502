"""The property tests of the reference's own suite, restated against the CUDA path.

Same properties, problem sizes, seeds and tolerances as the reference's
``tests/test_archetypal_analysis.py``, ``tests/test_gpnh_convex_coding.py``,
``tests/test_simplex_projection.py`` (each test cites the lines it mirrors); the
dictionary solver runs with its default 10 000-iteration cap exactly as there.
"""

import warnings

import numpy as np
import pytest
from sklearn.utils import check_random_state

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
if not torch.cuda.is_available():          # pragma: no cover
    pytest.skip('needs a CUDA device', allow_module_level=True)

from convex_dim_red import KernelAA, right_stochastic_matrix, simplex_project_rows       # noqa: E402
from convex_dim_red.archetypal_analysis import (                                          # noqa: E402
    _iterate_kernel_aa, _kernel_aa_cost, _update_kernel_aa_dictionary,
    _update_kernel_aa_weights)
from convex_dim_red.gpnh_convex_coding import (                                           # noqa: E402
    _gpnh_cost, _iterate_gpnh_convex_coding, _update_gpnh_dictionary, _update_gpnh_weights)


@pytest.fixture(autouse=True)
def _quiet():
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        yield


def _random_kernel_problem(n_features, n_components, n_samples, seed=0):
    rs = check_random_state(seed)
    X = rs.uniform(size=(n_samples, n_features))
    K = X.dot(X.T)
    C = right_stochastic_matrix((n_components, n_samples), random_state=rs)
    Z = right_stochastic_matrix((n_samples, n_components), random_state=rs)
    return rs, X, K, C, Z


def _exact_aa_solution(n_features, n_components, n_samples, seed=0):
    """Samples that are exact convex combinations of n_components of them
    (reference tests/test_archetypal_analysis.py:97-140)."""
    rs = check_random_state(seed)
    basis = rs.uniform(size=(n_components, n_features))
    Z = right_stochastic_matrix((n_samples, n_components), random_state=rs)
    picks = []
    while len(picks) < n_components:
        idx = rs.randint(low=0, high=n_samples)
        if idx not in picks:
            picks.append(idx)
    C = np.zeros((n_components, n_samples))
    for comp, idx in enumerate(picks):
        C[comp, idx] = 1.0
        Z[idx] = 0.0
        Z[idx, comp] = 1.0
    X = Z.dot(basis)
    return X, X.dot(X.T), C, Z


# ---------------------------------------------------------------- AA: dictionary update
@pytest.mark.parametrize('delta', [0.0, 0.1])
def test_single_dictionary_update_reduces_cost(delta):
    # tests/test_archetypal_analysis.py:20-95
    rs, X, K, C, Z = _random_kernel_problem(10, 5, 400)
    alpha = np.ones(5) if delta == 0 else rs.uniform(low=1 - delta, high=1 + delta, size=5)
    initial = _kernel_aa_cost(K, Z, C, alpha)
    C_new = _update_kernel_aa_dictionary(K, C, alpha, np.trace(K), K.dot(Z), Z.T.dot(Z))
    final = _kernel_aa_cost(K, Z, C_new, alpha)
    assert final <= initial
    assert np.allclose(C_new.sum(axis=1), 1, 1e-12)


def test_exact_solution_is_dictionary_update_fixed_point():
    # tests/test_archetypal_analysis.py:97-165
    X, K, C, Z = _exact_aa_solution(10, 6, 100)
    alpha = np.ones(6)
    initial = _kernel_aa_cost(K, Z, C, alpha)
    C_new = _update_kernel_aa_dictionary(K, C, alpha, np.trace(K), K.dot(Z), Z.T.dot(Z))
    final = _kernel_aa_cost(K, Z, C_new, alpha)
    assert abs(final - initial) < 1e-12
    assert np.allclose(C_new.sum(axis=1), 1, 1e-12)
    assert np.allclose(C_new, C, 1e-12)


@pytest.mark.parametrize('delta', [0.0, 0.2])
def test_repeated_dictionary_updates_converge(delta):
    # tests/test_archetypal_analysis.py:168-256 (20 features, 15 components, 600 samples)
    rs, X, K, C, Z = _random_kernel_problem(20, 15, 600)
    alpha = np.ones(15) if delta == 0 else rs.uniform(low=1 - delta, high=1 + delta, size=15)
    initial = _kernel_aa_cost(K, Z, C, alpha)
    Z_new, C_new, alpha_new, _, n_iter = _iterate_kernel_aa(
        K, Z, C, alpha, delta=delta, update_weights=False, update_dictionary=True,
        update_scale_factors=False, tolerance=1e-6, max_iterations=1000,
        require_monotonic_cost_decrease=True)[:5]
    final = _kernel_aa_cost(K, Z_new, C_new, alpha_new)
    assert final <= initial
    assert n_iter < 1000
    assert np.allclose(Z_new, Z, 1e-12)
    assert np.allclose(alpha_new, alpha, 1e-12)
    assert np.allclose(C_new.sum(axis=1), 1, 1e-12)


# ---------------------------------------------------------------- AA: weights update
@pytest.mark.parametrize('delta', [0.0, 0.1])
def test_single_weights_update_reduces_cost(delta):
    # tests/test_archetypal_analysis.py:259-330
    rs, X, K, C, Z = _random_kernel_problem(13, 7, 300)
    alpha = np.ones(7) if delta == 0 else rs.uniform(low=1 - delta, high=1 + delta, size=7)
    initial = _kernel_aa_cost(K, Z, C, alpha)
    CK = C.dot(K)
    Z_new = _update_kernel_aa_weights(Z, alpha, CK, CK.dot(C.T))
    final = _kernel_aa_cost(K, Z_new, C, alpha)
    assert final <= initial
    assert np.allclose(Z_new.sum(axis=1), 1, 1e-12)


def test_exact_solution_is_weights_update_fixed_point():
    # tests/test_archetypal_analysis.py:333-402
    X, K, C, Z = _exact_aa_solution(10, 6, 100)
    alpha = np.ones(6)
    initial = _kernel_aa_cost(K, Z, C, alpha)
    CK = C.dot(K)
    Z_new = _update_kernel_aa_weights(Z, alpha, CK, CK.dot(C.T))
    final = _kernel_aa_cost(K, Z_new, C, alpha)
    assert abs(final - initial) < 1e-12
    assert np.allclose(Z_new.sum(axis=1), 1, 1e-12)
    assert np.allclose(Z_new, Z, 1e-12)


@pytest.mark.parametrize('delta', [0.0, 0.2])
def test_repeated_weights_updates_converge(delta):
    # tests/test_archetypal_analysis.py:405-493
    rs, X, K, C, Z = _random_kernel_problem(30, 10, 500)
    alpha = np.ones(10) if delta == 0 else rs.uniform(low=1 - delta, high=1 + delta, size=10)
    initial = _kernel_aa_cost(K, Z, C, alpha)
    Z_new, C_new, alpha_new, _, n_iter = _iterate_kernel_aa(
        K, Z, C, alpha, delta=delta, update_weights=True, update_dictionary=False,
        update_scale_factors=False, tolerance=1e-6, max_iterations=100,
        require_monotonic_cost_decrease=True)[:5]
    final = _kernel_aa_cost(K, Z_new, C_new, alpha_new)
    assert final <= initial
    assert n_iter < 100
    assert np.allclose(C_new, C, 1e-12)
    assert np.allclose(alpha_new, alpha, 1e-12)
    assert np.allclose(Z_new.sum(axis=1), 1, 1e-12)


# ---------------------------------------------------------------- AA: vertex recovery
def test_finds_elements_of_3_point_convex_hull():
    # tests/test_archetypal_analysis.py:496-543 -- index-exact
    rs = check_random_state(0)
    n_samples, k = 50, 3
    basis = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
    expected_Z = right_stochastic_matrix((n_samples, k), random_state=rs)
    assignments = np.array([5, 27, 32])
    for i in range(k):
        expected_Z[assignments[i]] = np.zeros(k)
        expected_Z[assignments[i], i] = 1
    X = expected_Z.dot(basis)
    K = X.dot(X.T)
    C = right_stochastic_matrix((k, n_samples), random_state=rs)
    Z = right_stochastic_matrix((n_samples, k), random_state=rs)
    aa = KernelAA(n_components=k, delta=0, init='custom', max_iterations=500, tolerance=1e-6)
    sol_Z = aa.fit_transform(K, dictionary=C, weights=Z, alpha=np.ones(k))
    assert aa.n_iter < 500
    assert np.allclose(aa.dictionary.sum(axis=1), 1, 1e-12)
    assert np.allclose(sol_Z.sum(axis=1), 1, 1e-12)
    assert sorted(aa.dictionary.argmax(axis=1)) == list(assignments)


def test_finds_elements_of_4_point_convex_hull():
    # tests/test_archetypal_analysis.py:546-606 -- index-exact
    rs = check_random_state(0)
    n_samples, k = 123, 4
    basis = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=float)
    expected_Z = right_stochastic_matrix((n_samples, k), random_state=rs)
    assignments = np.array([8, 9, 56, 90])
    for i in range(k):
        expected_Z[assignments[i]] = np.zeros(k)
        expected_Z[assignments[i], i] = 1
    X = expected_Z.dot(basis)
    K = X.dot(X.T)
    C = right_stochastic_matrix((k, n_samples), random_state=rs)
    Z = right_stochastic_matrix((n_samples, k), random_state=rs)
    aa = KernelAA(n_components=k, delta=0, init='custom', max_iterations=500, tolerance=1e-12)
    sol_Z = aa.fit_transform(K, dictionary=C, weights=Z, alpha=np.ones(k))
    assert np.allclose(aa.dictionary.sum(axis=1), 1, 1e-12)
    assert np.allclose(sol_Z.sum(axis=1), 1, 1e-12)
    assert sorted(aa.dictionary.argmax(axis=1)) == list(assignments)


# ---------------------------------------------------------------- GPNH
def _gw(n_features, k):
    return 4.0 / (n_features * k * (k - 1)) * (k * np.eye(k) - 1)


def test_cost_is_zero_for_perfect_reconstruction():
    # tests/test_gpnh_convex_coding.py:18-42
    rs = check_random_state(0)
    W = rs.uniform(size=(5, 3))
    Z = right_stochastic_matrix((30, 3), random_state=rs)
    assert abs(_gpnh_cost(Z.dot(W.T), Z, W, lambda_W=0)) < 1e-14


@pytest.mark.parametrize('n_features,k,n_samples,lam', [(7, 5, 450, 0.0), (11, 6, 230, 3.2)])
def test_single_gpnh_dictionary_update_reduces_cost(n_features, k, n_samples, lam):
    # tests/test_gpnh_convex_coding.py:45-106
    rs = check_random_state(0)
    X = rs.uniform(size=(n_samples, n_features))
    W = rs.uniform(size=(n_features, k))
    Z = right_stochastic_matrix((n_samples, k), random_state=rs)
    initial = _gpnh_cost(X, Z, W, lambda_W=lam)
    W_new = _update_gpnh_dictionary(X, Z, Z.T.dot(Z), _gw(n_features, k), lambda_W=lam)
    assert _gpnh_cost(X, Z, W_new, lambda_W=lam) <= initial


def test_exact_solution_is_gpnh_dictionary_update_fixed_point():
    # tests/test_gpnh_convex_coding.py:109-144
    rs = check_random_state(0)
    n_features, k, n_samples = 10, 6, 100
    W = rs.uniform(size=(n_features, k))
    Z = right_stochastic_matrix((n_samples, k), random_state=rs)
    X = Z.dot(W.T)
    W_new = _update_gpnh_dictionary(X, Z, Z.T.dot(Z), _gw(n_features, k), lambda_W=0)
    assert abs(_gpnh_cost(X, Z, W_new, 0) - _gpnh_cost(X, Z, W, 0)) < 1e-6
    assert np.allclose(W_new, W, 1e-6)


@pytest.mark.parametrize('lam', [0.0, 4.2])
def test_repeated_gpnh_dictionary_updates_converge(lam):
    # tests/test_gpnh_convex_coding.py:147-216
    rs = check_random_state(0)
    n_features, k, n_samples = 13, 3, 50
    X = rs.uniform(size=(n_samples, n_features))
    W = rs.uniform(size=(n_features, k))
    Z = right_stochastic_matrix((n_samples, k), random_state=rs)
    initial = _gpnh_cost(X, Z, W, lambda_W=lam)
    Z_new, W_new, _, n_iter = _iterate_gpnh_convex_coding(
        X, Z, W, lambda_W=lam, update_weights=False, update_dictionary=True,
        tolerance=1e-6, max_iterations=100, require_monotonic_cost_decrease=True)[:4]
    assert _gpnh_cost(X, Z_new, W_new, lambda_W=lam) <= initial
    assert n_iter < 100
    assert np.allclose(Z_new, Z, 1e-12)


@pytest.mark.parametrize('lam', [0.0, 3.2])
def test_single_gpnh_weights_update_reduces_cost(lam):
    # tests/test_gpnh_convex_coding.py:219-272
    rs = check_random_state(0)
    n_features, k, n_samples = 25, 6, 300
    X = rs.uniform(size=(n_samples, n_features))
    W = rs.uniform(size=(n_features, k))
    Z = right_stochastic_matrix((n_samples, k), random_state=rs)
    initial = _gpnh_cost(X, Z, W, lambda_W=lam)
    Z_new = _update_gpnh_weights(X, Z, W)
    assert _gpnh_cost(X, Z_new, W, lambda_W=lam) <= initial
    assert np.allclose(Z_new.sum(axis=1), 1, 1e-12)


def test_exact_solution_is_gpnh_weights_update_fixed_point():
    # tests/test_gpnh_convex_coding.py:275-340
    rs = check_random_state(0)
    n_features, k, n_samples = 10, 5, 150
    W = rs.uniform(size=(n_features, k))
    Z = right_stochastic_matrix((n_samples, k), random_state=rs)
    X = Z.dot(W.T)
    Z_new = _update_gpnh_weights(X, Z, W)
    assert abs(_gpnh_cost(X, Z_new, W, 0) - _gpnh_cost(X, Z, W, 0)) < 1e-6
    assert np.allclose(Z_new.sum(axis=1), 1, 1e-12)
    assert np.allclose(Z_new, Z, 1e-6)


@pytest.mark.parametrize('lam', [0.0, 3.2])
def test_repeated_gpnh_weights_updates_converge(lam):
    # tests/test_gpnh_convex_coding.py:343-412
    rs = check_random_state(0)
    n_features, k, n_samples = 10, 5, 200
    X = rs.uniform(size=(n_samples, n_features))
    W = rs.uniform(size=(n_features, k))
    Z = right_stochastic_matrix((n_samples, k), random_state=rs)
    initial = _gpnh_cost(X, Z, W, lambda_W=lam)
    Z_new, W_new, _, n_iter = _iterate_gpnh_convex_coding(
        X, Z, W, lambda_W=lam, update_weights=True, update_dictionary=False,
        tolerance=1e-6, max_iterations=100, require_monotonic_cost_decrease=True)[:4]
    assert _gpnh_cost(X, Z_new, W_new, lambda_W=lam) <= initial
    assert n_iter < 100
    assert np.allclose(W_new, W, 1e-12)
    assert np.allclose(Z_new.sum(axis=1), 1, 1e-12)


# ---------------------------------------------------------------- simplex feasibility
@pytest.mark.parametrize('shape', [(1, 5), (10, 3), (341, 317)])
def test_projected_rows_are_feasible(shape):
    # tests/test_simplex_projection.py:179-208
    rs = check_random_state(0)
    A = rs.uniform(low=-5, high=5, size=shape)
    P = simplex_project_rows(A)
    assert np.all(P >= 0)
    assert np.allclose(P.sum(axis=1), 1, 1e-14)
