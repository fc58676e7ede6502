"""Pins the CPU oracle (``oracle/``) against the real reference.

Every expected value below comes from ``tests/golden/golden_v1.npz`` (produced by
importing the unmodified reference, see ``tests/golden/make_golden.py``) or from
the known-answer vectors of the reference's own tests (cited per test).
"""

import numpy as np
import pytest

from oracle import convex_oracle as orc

RT = 1e-9


def close(a, b, rtol=RT, atol=1e-12):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


# ---------------------------------------------------------------- simplex
@pytest.mark.parametrize('n', [1, 2, 3, 5, 8, 17, 64, 317])
def test_simplex_vector_golden(golden, n):
    x = golden['simplex/vec%d/x' % n]
    close(orc.simplex_project_vector(x), golden['simplex/vec%d/out' % n], rtol=1e-13, atol=1e-15)
    close(orc.simplex_project_vector_py(x), golden['simplex/vec%d/out' % n], rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize('name', ['feasible', 'ties'])
def test_simplex_special_golden(golden, name):
    close(orc.simplex_project_vector(golden['simplex/%s/x' % name]),
          golden['simplex/%s/out' % name], rtol=1e-13, atol=1e-15)


def test_simplex_known_answers():
    # reference tests/test_simplex_projection.py:13-57 and :166-176
    close(orc.simplex_project_vector(np.array([0.8, 0.8])), [0.5, 0.5], atol=1e-15)
    close(orc.simplex_project_vector(np.array([0.5, -0.5])), [1.0, 0.0], atol=1e-15)
    close(orc.simplex_project_vector(np.array([1.0])), [1.0], atol=1e-15)
    close(orc.simplex_project_vector(np.array([3.0])), [1.0], atol=1e-15)
    A = np.array([[0.5, 0.5], [0.5, 1.0], [0.0, -0.5]])
    close(orc.simplex_project_rows(A), [[0.5, 0.5], [0.25, 0.75], [0.75, 0.25]], atol=1e-15)


def test_simplex_rows_cols_golden(golden):
    A = golden['simplex/rows/A']
    close(orc.simplex_project_rows(A), golden['simplex/rows/out'], rtol=1e-13, atol=1e-15)
    close(orc.simplex_project_columns(A), golden['simplex/cols/out'], rtol=1e-13, atol=1e-15)
    A = golden['simplex/rows_wide/A']
    out = orc.simplex_project_rows(A)
    close(out, golden['simplex/rows_wide/out'], rtol=1e-12, atol=1e-15)
    assert np.all(out >= 0)
    close(out.sum(axis=1), 1.0, atol=1e-13)


# ---------------------------------------------------------------- QP
@pytest.mark.parametrize('k', [2, 3, 8, 16])
def test_qp_golden(golden, k):
    A, b, x0 = (golden['qp/k%d/%s' % (k, n)] for n in ('A', 'b', 'x0'))
    close(orc.quad_simplex_spg(A, b, x0, max_iterations=3), golden['qp/k%d/x_it3' % k], rtol=1e-10)
    # converged solutions agree to the solver tolerance (epsilon_two = 1e-6)
    close(orc.quad_simplex_spg(A, b, x0), golden['qp/k%d/x' % k], rtol=0, atol=5e-6)
    close(orc.quad_simplex_spg_py(A, b, x0), orc.quad_simplex_spg(A, b, x0), rtol=0, atol=5e-6)


def test_qp_batched_golden(golden):
    A, CK, Z0 = golden['qp/batch/A'], golden['qp/batch/CK'], golden['qp/batch/Z0']
    one = np.ones(A.shape[0])
    close(orc.update_kernel_aa_weights(Z0, one, CK, A, max_iterations=1),
          golden['qp/batch/Z_aa_it1'], rtol=1e-10)
    close(orc.update_kernel_aa_weights(Z0, one, CK, A), golden['qp/batch/Z_aa'], rtol=0, atol=5e-6)
    close(orc.update_kernel_aa_weights(Z0, golden['qp/batch/alpha2'], CK, A),
          golden['qp/batch/Z_aa_alpha2'], rtol=0, atol=5e-6)


# ---------------------------------------------------------------- generic SPG
def test_spg_golden(golden):
    # reference tests/test_spg.py:37-90: quartic on [-1, 0.5] -> x = 0, f = 1
    x, fx, n_it, n_fe = orc.spg(lambda x: x ** 4 + 2 * x ** 2 + 1,
                                lambda x: 4 * x ** 3 + 4 * x, 0.4,
                                project=lambda x: min(max(x, -1.0), 0.5))
    g = golden['spg/quartic/out']
    assert abs(x) < 1e-6 and abs(fx - 1) < 1e-6
    close([x, fx, n_it, n_fe], g, rtol=1e-12, atol=1e-14)
    M, y, x0 = golden['spg/box/M'], golden['spg/box/y'], golden['spg/box/x0']
    x, fx, n_it, n_fe = orc.spg(lambda v: 0.5 * (M.dot(v) - y).dot(M.dot(v) - y),
                                lambda v: M.T.dot(M.dot(v) - y), x0,
                                project=lambda v: np.fmin(np.fmax(v, 0.0), 0.3))
    close(x, golden['spg/box/x'], rtol=1e-9)
    close([fx, n_it, n_fe], golden['spg/box/stats'], rtol=1e-12)


# ---------------------------------------------------------------- furthest sum
def test_furthest_sum_golden(golden):
    D = golden['fs/D']
    for i, (k, start, extra) in enumerate(golden['fs/cases']):
        sel = orc.furthest_sum(D, int(k), int(start), list(golden['fs/case%d/exclude' % i]), int(extra))
        assert np.array_equal(np.asarray(sel, dtype=np.int64), golden['fs/case%d/selected' % i])
    # reference tests/test_furthest_sum.py:142-161: picks {0, 2} on the 3x3 matrix
    K3 = golden['fs/k3/D']
    for start in range(3):
        for extra in range(1, 11):
            assert set(orc.furthest_sum(K3, 2, start, [], extra)) == {0, 2}
    assert np.array_equal(orc.furthest_sum(K3, 2, 1, [], 3), golden['fs/k3/sel'])


def test_furthest_sum_edge_cases():
    D = np.zeros((4, 4))
    assert list(orc.furthest_sum(D, 0, 0)) == []
    with pytest.raises(ValueError):
        orc.furthest_sum(np.zeros((3, 4)), 2, 0)
    with pytest.raises(ValueError):
        orc.furthest_sum(D, 2, 7)
    with pytest.raises(ValueError):
        orc.furthest_sum(D, 2, 1, [1])
    with pytest.raises(ValueError):
        orc.furthest_sum(D, 4, 0, [1])
    # unit square corners: reference tests/test_furthest_sum.py:164-194
    pts = np.array([[0, 0], [.5, .5], [.2, .7], [.4, .1], [1, 0], [.6, .6], [0, 1],
                    [.3, .3], [.7, .2], [1, 1]], dtype=float)
    Dm = np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1))
    assert set(orc.furthest_sum(Dm, 4, 1, [], 10)) == {0, 4, 6, 9}


# ---------------------------------------------------------------- AA
def _aa_inputs(golden):
    X, C0, Z0 = golden['aa/X'], golden['aa/C0'], golden['aa/Z0']
    K = X.dot(X.T)
    return X, K, C0, Z0, np.ones(C0.shape[0])


def test_aa_single_functions(golden):
    X, K, C0, Z0, alpha = _aa_inputs(golden)
    ZtZ, KZ, trK = Z0.T.dot(Z0), K.dot(Z0), K.trace()
    close(orc.kernel_aa_cost(K, Z0, C0, alpha), golden.scalar('aa/kernel_cost'), rtol=1e-12)
    close(orc.aa_dictionary_cost(X, C0, trK, KZ, ZtZ), golden.scalar('aa/dict_cost'), rtol=1e-12)
    close(orc.aa_dictionary_gradient(X, C0, KZ, ZtZ), golden['aa/dict_grad'], rtol=1e-10)
    close(orc.kernel_aa_dictionary_cost(K, C0, trK, KZ, ZtZ), golden.scalar('aa/kdict_cost'), rtol=1e-12)
    close(orc.kernel_aa_dictionary_gradient(K, C0, KZ, ZtZ), golden['aa/kdict_grad'], rtol=1e-10)
    for it in (1, 2, 5):
        close(orc.update_kernel_aa_dictionary(K, C0, alpha, trK, KZ, ZtZ, max_iterations=it),
              golden['aa/kdict_update_it%d' % it], rtol=1e-8, atol=1e-11)
        close(orc.update_aa_dictionary(X, C0, alpha, trK, KZ, ZtZ, max_iterations=it),
              golden['aa/dict_update_it%d' % it], rtol=1e-8, atol=1e-11)
    CK = C0.dot(K)
    close(orc.update_kernel_aa_weights(Z0, alpha, CK, CK.dot(C0.T)),
          golden['aa/weights_update'], rtol=0, atol=5e-6)


@pytest.mark.parametrize('mode', ['kaa', 'faa'])
@pytest.mark.parametrize('tag', ['d1', 'd3w5', 'rel'])
def test_aa_iterate_golden(golden, mode, tag):
    X, K, C0, Z0, alpha = _aa_inputs(golden)
    kw = {'d1': dict(dictionary_solver_kwargs=dict(max_iterations=1)),
          'd3w5': dict(dictionary_solver_kwargs=dict(max_iterations=3),
                       weights_solver_kwargs=dict(max_iterations=5)),
          'rel': dict(dictionary_solver_kwargs=dict(max_iterations=1),
                      weights_solver_kwargs=dict(max_iterations=1),
                      stopping_criterion='rel_delta_f')}[tag]
    fn, data = (orc.iterate_kernel_aa, K) if mode == 'kaa' else (orc.iterate_aa, X)
    Z, C, a, cost, n_iter, _, deltas = fn(data, Z0.copy(), C0.copy(), alpha.copy(),
                                          tolerance=1e-9, max_iterations=12, **kw)
    gcost, gn = golden['aa/%s_%s/stats' % (mode, tag)]
    assert n_iter == int(gn)
    close(cost, gcost, rtol=1e-7)
    close(deltas, golden['aa/%s_%s/deltas' % (mode, tag)], rtol=1e-4, atol=1e-9)
    close(Z, golden['aa/%s_%s/Z' % (mode, tag)], rtol=0, atol=2e-5)
    close(C, golden['aa/%s_%s/C' % (mode, tag)], rtol=0, atol=2e-5)


def test_aa_frozen_and_delta(golden):
    X, K, C0, Z0, alpha = _aa_inputs(golden)
    Z, C, a, cost, n_iter, _, _ = orc.iterate_kernel_aa(
        K, Z0.copy(), C0.copy(), alpha.copy(), update_dictionary=False,
        tolerance=1e-9, max_iterations=5)
    assert np.array_equal(C, C0)
    close(Z, golden['aa/kaa_frozenC/Z'], rtol=0, atol=5e-6)
    close([cost, n_iter], golden['aa/kaa_frozenC/stats'], rtol=1e-7)
    Z, C, a, cost, n_iter, _, _ = orc.iterate_kernel_aa(
        K, Z0.copy(), C0.copy(), alpha.copy(), delta=0.2, tolerance=1e-9,
        max_iterations=6, dictionary_solver_kwargs=dict(max_iterations=1))
    close(a, golden['aa/kaa_delta/alpha'], rtol=1e-6)
    close(C, golden['aa/kaa_delta/C'], rtol=0, atol=2e-5)
    close(Z, golden['aa/kaa_delta/Z'], rtol=0, atol=2e-5)
    close([cost, n_iter], golden['aa/kaa_delta/stats'], rtol=1e-7)


@pytest.mark.parametrize('init', ['random', 'furthest_sum'])
def test_aa_estimator_flow_golden(golden, init):
    """Replays ArchetypalAnalysis.fit_transform / transform (archetypal_analysis.py:
    1026-1199) with the oracle pieces, including the RNG draw order."""
    X = golden['aa/X']
    k = golden['aa/C0'].shape[0]
    rng = np.random.RandomState(0)
    K = X.dot(X.T)
    C, Z = orc.init_kernel_aa(K, k, init, rng)
    Z, C, a, cost, n_iter, _, deltas = orc.iterate_aa(
        X, Z, C, np.ones(k), tolerance=1e-6, max_iterations=40,
        dictionary_solver_kwargs=dict(max_iterations=1))
    gcost, gn = golden['aa/est_%s/stats' % init]
    assert n_iter == int(gn)
    close(cost, gcost, rtol=1e-6)
    close(Z, golden['aa/est_%s/Z' % init], rtol=0, atol=1e-4)
    close(C, golden['aa/est_%s/C' % init], rtol=0, atol=1e-4)
    close(C.dot(X), golden['aa/est_%s/archetypes' % init], rtol=0, atol=1e-4)
    Xv = golden['aa/est_%s/Xv' % init]
    Z0v = orc.right_stochastic_matrix((Xv.shape[0], k), rng)
    Zv, cv = orc.aa_transform(golden['aa/est_%s/archetypes' % init], Xv, Z0v, 40)
    close(Zv, golden['aa/est_%s/Zv' % init], rtol=0, atol=1e-5)
    close(cv, golden.scalar('aa/est_%s/cost_v' % init), rtol=1e-7)


@pytest.mark.parametrize('init', ['random', 'furthest_sum'])
def test_kernel_aa_estimator_flow_golden(golden, init):
    X = golden['aa/X']
    k = golden['aa/C0'].shape[0]
    rng = np.random.RandomState(0)
    K = X.dot(X.T)
    C, Z = orc.init_kernel_aa(K, k, init, rng)
    Z, C, a, cost, n_iter, _, _ = orc.iterate_kernel_aa(
        K, Z, C, np.ones(k), tolerance=1e-6, max_iterations=40,
        dictionary_solver_kwargs=dict(max_iterations=2))
    gcost, gn = golden['aa/kest_%s/stats' % init]
    assert n_iter == int(gn)
    close(cost, gcost, rtol=1e-6)
    close(Z, golden['aa/kest_%s/Z' % init], rtol=0, atol=1e-4)
    close(C, golden['aa/kest_%s/C' % init], rtol=0, atol=1e-4)


def test_kernel_aa_triangle_vertices(golden):
    # index-exact pin in the style of reference tests/test_archetypal_analysis.py:496-606
    X = golden['aa/tri/X']
    K = X.dot(X.T)
    rng = np.random.RandomState(0)
    C, Z = orc.init_kernel_aa(K, 3, 'furthest_sum', rng)
    Z, C, a, cost, n_iter, _, _ = orc.iterate_kernel_aa(
        K, Z, C, np.ones(3), tolerance=1e-8, max_iterations=200,
        dictionary_solver_kwargs=dict(max_iterations=20))
    assert np.array_equal(np.sort(np.argmax(C, axis=1)), golden['aa/tri/picks'])
    assert np.array_equal(golden['aa/tri/picks'], [5, 27, 32])
    close(cost, golden['aa/tri/stats'][0], rtol=0, atol=1e-7)


# ---------------------------------------------------------------- GPNH
def test_gpnh_single_functions(golden):
    X, W0, Z0 = golden['gpnh/X'], golden['gpnh/W0'], golden['gpnh/Z0']
    T, d = X.shape
    k = W0.shape[1]
    close(orc.gpnh_regularization(W0), golden.scalar('gpnh/reg'), rtol=1e-12)
    close(orc.gpnh_cost(X, Z0, W0, 0.0), golden.scalar('gpnh/cost0'), rtol=1e-12)
    close(orc.gpnh_cost(X, Z0, W0, 3.2), golden.scalar('gpnh/cost_l'), rtol=1e-12)
    ZtZ, GW = Z0.T.dot(Z0), orc.gpnh_GW(d, k)
    for lam in (0.0, 3.2):
        close(orc.update_gpnh_dictionary(X, Z0, ZtZ, GW, lambda_W=lam),
              golden['gpnh/dict_update_l%g' % lam], rtol=1e-10)
    close(orc.update_gpnh_weights(X, Z0, W0, max_iterations=2),
          golden['gpnh/weights_update_it2'], rtol=1e-9)
    close(orc.update_gpnh_weights(X, Z0, W0), golden['gpnh/weights_update'], rtol=0, atol=5e-6)


@pytest.mark.parametrize('lam', [0.0, 3.2])
@pytest.mark.parametrize('wtag', ['full', 'w1'])
def test_gpnh_iterate_golden(golden, lam, wtag):
    X, W0, Z0 = golden['gpnh/X'], golden['gpnh/W0'], golden['gpnh/Z0']
    kw = {} if wtag == 'full' else dict(weights_solver_kwargs=dict(max_iterations=1))
    Z, W, cost, n_iter, _, deltas = orc.iterate_gpnh(
        X, Z0.copy(), W0.copy(), lambda_W=lam, tolerance=1e-9, max_iterations=10, **kw)
    tag = 'gpnh/it_l%g_%s' % (lam, wtag)
    gcost, gn = golden[tag + '/stats']
    assert n_iter == int(gn)
    close(cost, gcost, rtol=1e-7)
    close(Z, golden[tag + '/Z'], rtol=0, atol=2e-5)
    close(W, golden[tag + '/W'], rtol=0, atol=2e-5)
    close(deltas, golden[tag + '/deltas'], rtol=1e-4, atol=1e-9)


def test_gpnh_frozen_dictionary(golden):
    X, W0, Z0 = golden['gpnh/X'], golden['gpnh/W0'], golden['gpnh/Z0']
    Z, W, cost, n_iter, _, _ = orc.iterate_gpnh(
        X, Z0.copy(), W0.copy(), update_dictionary=False, tolerance=1e-9, max_iterations=4)
    assert np.array_equal(W, W0)
    close(Z, golden['gpnh/frozenW/Z'], rtol=0, atol=5e-6)
    close([cost, n_iter], golden['gpnh/frozenW/stats'], rtol=1e-7)


@pytest.mark.parametrize('init', ['random', 'furthest_sum'])
def test_gpnh_estimator_flow_golden(golden, init):
    X = golden['gpnh/X']
    k = golden['gpnh/W0'].shape[1]
    rng = np.random.RandomState(0)
    W, Z = orc.init_gpnh(X, k, init, rng)
    Z, W, cost, n_iter, _, deltas = orc.iterate_gpnh(
        X, Z, W, lambda_W=0.5, tolerance=1e-6, max_iterations=30)
    gcost, gn = golden['gpnh/est_%s/stats' % init]
    assert n_iter == int(gn)
    close(cost, gcost, rtol=1e-6)
    close(Z, golden['gpnh/est_%s/Z' % init], rtol=0, atol=1e-4)
    close(W, golden['gpnh/est_%s/W' % init], rtol=0, atol=1e-4)
    # transform = weights-only run with a fresh random start from the same rng
    Xv = golden['gpnh/est_%s/Xv' % init]
    Z0v = orc.right_stochastic_matrix((Xv.shape[0], k), rng)
    Zv, _, cv, _, _, _ = orc.iterate_gpnh(Xv, Z0v, golden['gpnh/est_%s/W' % init],
                                          lambda_W=0.5, update_dictionary=False,
                                          tolerance=1e-6, max_iterations=30)
    close(Zv, golden['gpnh/est_%s/Zv' % init], rtol=0, atol=1e-5)
    close(cv, golden.scalar('gpnh/est_%s/cost_v' % init), rtol=1e-7)


# ---------------------------------------------------------------- k-means
@pytest.mark.parametrize('tag,tol,max_iter', [('conv', 1e-4, 300), ('it2', 1e-4, 2), ('tol0', 0.0, 300)])
def test_kmeans_golden(golden, tag, tol, max_iter):
    X, picks = golden['km/X'], golden['km/picks']
    labels, centres, inertia, n_iter = orc.kmeans_lloyd(X, X[picks], tol=tol, max_iter=max_iter)
    assert labels.dtype == np.int32
    assert np.array_equal(labels, golden['km/%s/labels' % tag])
    close(centres, golden['km/%s/centres' % tag], rtol=1e-10, atol=1e-12)
    close([inertia, n_iter], golden['km/%s/stats' % tag], rtol=1e-10)


def test_kmeans_empty_cluster_golden(golden):
    X = golden['km/X']
    labels, centres, inertia, n_iter = orc.kmeans_lloyd(X, golden['km/empty/init'])
    assert np.array_equal(labels, golden['km/empty/labels'])
    close(centres, golden['km/empty/centres'], rtol=1e-10, atol=1e-12)
    close([inertia, n_iter], golden['km/empty/stats'], rtol=1e-10)


def test_kmeans_furthest_sum_picks(golden):
    X = golden['km/X']
    D = orc.dissimilarity_from_kernel(X.dot(X.T))
    D = np.nan_to_num(D)
    assert np.array_equal(orc.furthest_sum(D, 5, 7, [], 10), golden['km/picks'])
