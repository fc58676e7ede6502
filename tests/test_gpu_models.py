"""GPU parity tests of the drop-in Python API (``-m gpu``): the alternating
loops, their single steps and the estimator classes, against the golden vectors
generated from the reference and against the CPU oracle.

Parity tiers: single steps with pinned iteration counts
agree to 1e-8; short alternating runs agree to 1e-7 in cost and 2e-5 in the
factors (the per-sample QPs stop on a 1e-6 projected-gradient norm, so their
minimisers are only defined to that accuracy); index outputs are exact.
"""

import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
if not torch.cuda.is_available():          # pragma: no cover
    pytest.skip('needs a CUDA device', allow_module_level=True)

from oracle import convex_oracle as orc                                   # noqa: E402
import convex_dim_red as cdr                                                # noqa: E402
from convex_dim_red import archetypal_analysis as aa                        # noqa: E402
from convex_dim_red import gpnh_convex_coding as gp                         # noqa: E402
from convex_dim_red.kmeans import kmeans_lloyd, furthest_sum_centres, KMeans  # noqa: E402


def close(a, b, rtol=1e-9, atol=1e-12):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


@pytest.fixture(autouse=True)
def _quiet():
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        yield


def _aa_inputs(golden):
    X, C0, Z0 = golden['aa/X'], golden['aa/C0'], golden['aa/Z0']
    return X, X.dot(X.T), C0, Z0, np.ones(C0.shape[0])


# ---------------------------------------------------------------- AA: single functions
def test_aa_single_functions_golden(golden):
    X, K, C0, Z0, alpha = _aa_inputs(golden)
    ZtZ, KZ, trK = Z0.T.dot(Z0), K.dot(Z0), K.trace()
    close(aa._kernel_aa_cost(K, Z0, C0, alpha), golden.scalar('aa/kernel_cost'), rtol=1e-11)
    close(aa._aa_dictionary_cost(X, C0, trK, KZ, ZtZ), golden.scalar('aa/dict_cost'), rtol=1e-11)
    close(aa._aa_dictionary_gradient(X, C0, KZ, ZtZ), golden['aa/dict_grad'], rtol=1e-9, atol=1e-11)
    close(aa._kernel_aa_dictionary_cost(K, C0, trK, KZ, ZtZ), golden.scalar('aa/kdict_cost'), rtol=1e-11)
    close(aa._kernel_aa_dictionary_gradient(K, C0, KZ, ZtZ), golden['aa/kdict_grad'], rtol=1e-9, atol=1e-11)
    CK = C0.dot(K)
    close(aa._update_kernel_aa_weights(Z0, alpha, CK, CK.dot(C0.T)),
          golden['aa/weights_update'], rtol=0, atol=5e-6)


@pytest.mark.parametrize('it', [1, 2, 5])
def test_aa_dictionary_update_golden(golden, it):
    X, K, C0, Z0, alpha = _aa_inputs(golden)
    ZtZ, KZ, trK = Z0.T.dot(Z0), K.dot(Z0), K.trace()
    C = aa._update_kernel_aa_dictionary(K, C0, alpha, trK, KZ, ZtZ, max_iterations=it)
    close(C, golden['aa/kdict_update_it%d' % it], rtol=1e-7, atol=1e-10)
    close(C.sum(axis=1), 1.0, atol=1e-12)
    C = aa._update_aa_dictionary(X, C0, alpha, trK, KZ, ZtZ, max_iterations=it)
    close(C, golden['aa/dict_update_it%d' % it], rtol=1e-7, atol=1e-10)


# ---------------------------------------------------------------- AA: alternating loop
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
    fn, data = (aa._iterate_kernel_aa, K) if mode == 'kaa' else (aa._iterate_aa, X)
    Z, C, a, cost, n_iter, avg_t, deltas = fn(data, Z0.copy(), C0.copy(), alpha.copy(),
                                              tolerance=1e-9, max_iterations=12, **kw)
    gcost, gn = golden['aa/%s_%s/stats' % (mode, tag)]
    assert n_iter == int(gn)
    assert len(deltas) == n_iter + 1 and avg_t > 0
    close(cost, gcost, rtol=1e-7)
    close(deltas, golden['aa/%s_%s/deltas' % (mode, tag)], rtol=1e-4, atol=1e-9)
    close(Z, golden['aa/%s_%s/Z' % (mode, tag)], rtol=0, atol=2e-5)
    close(C, golden['aa/%s_%s/C' % (mode, tag)], rtol=0, atol=2e-5)
    close(Z.sum(axis=1), 1.0, atol=1e-12)
    close(C.sum(axis=1), 1.0, atol=1e-12)


def test_aa_graph_and_eager_agree(golden, monkeypatch):
    X, K, C0, Z0, alpha = _aa_inputs(golden)
    kw = dict(tolerance=1e-9, max_iterations=12, dictionary_solver_kwargs=dict(max_iterations=1))
    r1 = aa._iterate_aa(X, Z0.copy(), C0.copy(), alpha.copy(), **kw)
    monkeypatch.setenv('CDR_NO_CUDA_GRAPH', '1')
    r2 = aa._iterate_aa(X, Z0.copy(), C0.copy(), alpha.copy(), **kw)
    assert np.array_equal(r1[0], r2[0]) and np.array_equal(r1[1], r2[1])
    assert r1[3] == r2[3] and r1[4] == r2[4]


def test_aa_frozen_and_delta_golden(golden):
    X, K, C0, Z0, alpha = _aa_inputs(golden)
    Z, C, a, cost, n_iter, _, _ = aa._iterate_kernel_aa(
        K, Z0.copy(), C0.copy(), alpha.copy(), update_dictionary=False,
        tolerance=1e-9, max_iterations=5)
    assert np.array_equal(C, C0)
    close(Z, golden['aa/kaa_frozenC/Z'], rtol=0, atol=5e-6)
    close([cost, n_iter], golden['aa/kaa_frozenC/stats'], rtol=1e-7)
    Z, C, a, cost, n_iter, _, _ = aa._iterate_kernel_aa(
        K, Z0.copy(), C0.copy(), alpha.copy(), delta=0.2, tolerance=1e-9,
        max_iterations=6, dictionary_solver_kwargs=dict(max_iterations=1))
    close(a, golden['aa/kaa_delta/alpha'], rtol=1e-6)
    close(C, golden['aa/kaa_delta/C'], rtol=0, atol=2e-5)
    close(Z, golden['aa/kaa_delta/Z'], rtol=0, atol=2e-5)
    close([cost, n_iter], golden['aa/kaa_delta/stats'], rtol=1e-7)


def test_aa_cost_increase_raises_like_reference(golden):
    """A line search that accepts every full step (gamma << 0) overshoots; the
    monotonicity check must fire at the same stage as in the oracle
    (archetypal_analysis.py:167-174)."""
    X, K, C0, Z0, alpha = _aa_inputs(golden)
    kw = dict(tolerance=0.0, max_iterations=60,
              dictionary_solver_kwargs=dict(max_iterations=1, gamma=-1e12, alpha0=50.0),
              weights_solver_kwargs=dict(max_iterations=1))
    try:
        orc.iterate_kernel_aa(K, Z0.copy(), C0.copy(), alpha.copy(), **kw)
        expected = None
    except RuntimeError as err:
        expected = str(err)
    assert expected is not None, 'test setup no longer provokes a cost increase'
    with pytest.raises(RuntimeError) as info:
        aa._iterate_kernel_aa(K, Z0.copy(), C0.copy(), alpha.copy(), **kw)
    assert str(info.value) == expected


# ---------------------------------------------------------------- AA: estimators
@pytest.mark.parametrize('init', ['random', 'furthest_sum'])
def test_archetypal_analysis_estimator_golden(golden, init):
    X = golden['aa/X']
    k = golden['aa/C0'].shape[0]
    m = cdr.ArchetypalAnalysis(n_components=k, init=init, tolerance=1e-6, max_iterations=40,
                               random_state=0, dictionary_solver_kwargs=dict(max_iterations=1))
    Z = m.fit_transform(X)
    gcost, gn = golden['aa/est_%s/stats' % init]
    assert m.n_iter == int(gn)
    close(m.cost, gcost, rtol=1e-6)
    close(Z, golden['aa/est_%s/Z' % init], rtol=0, atol=1e-4)
    close(m.dictionary, golden['aa/est_%s/C' % init], rtol=0, atol=1e-4)
    close(m.archetypes, golden['aa/est_%s/archetypes' % init], rtol=0, atol=1e-4)
    close(m.cost_deltas, golden['aa/est_%s/deltas' % init], rtol=1e-3, atol=1e-8)
    Zv, cv = m.transform(golden['aa/est_%s/Xv' % init])
    close(Zv, golden['aa/est_%s/Zv' % init], rtol=0, atol=1e-4)
    close(cv, golden.scalar('aa/est_%s/cost_v' % init), rtol=1e-5)
    close(m.inverse_transform(Zv), Zv.dot(m.archetypes))


@pytest.mark.parametrize('init', ['random', 'furthest_sum'])
def test_kernel_aa_estimator_golden(golden, init):
    X = golden['aa/X']
    K = X.dot(X.T)
    k = golden['aa/C0'].shape[0]
    m = cdr.KernelAA(n_components=k, init=init, tolerance=1e-6, max_iterations=40,
                     random_state=0, dictionary_solver_kwargs=dict(max_iterations=2))
    Z = m.fit_transform(K)
    gcost, gn = golden['aa/kest_%s/stats' % init]
    assert m.n_iter == int(gn)
    close(m.cost, gcost, rtol=1e-6)
    close(Z, golden['aa/kest_%s/Z' % init], rtol=0, atol=1e-4)
    close(m.dictionary, golden['aa/kest_%s/C' % init], rtol=0, atol=1e-4)


def test_kernel_aa_triangle_vertices(golden):
    # index-exact pin (reference tests/test_archetypal_analysis.py:496-606 style)
    X = golden['aa/tri/X']
    m = cdr.KernelAA(n_components=3, init='furthest_sum', tolerance=1e-8, max_iterations=200,
                     random_state=0, dictionary_solver_kwargs=dict(max_iterations=20))
    m.fit_transform(X.dot(X.T))
    assert np.array_equal(np.sort(np.argmax(m.dictionary, axis=1)), [5, 27, 32])
    close(m.cost, golden['aa/tri/stats'][0], rtol=0, atol=1e-7)


def test_aa_estimator_validation():
    X = np.random.RandomState(0).uniform(size=(10, 4))
    with pytest.raises(ValueError):
        cdr.ArchetypalAnalysis(n_components=0).fit(X)
    with pytest.raises(ValueError):
        cdr.ArchetypalAnalysis(n_components=2, max_iterations=0).fit(X)
    with pytest.raises(ValueError):
        cdr.ArchetypalAnalysis(n_components=2, tolerance=-1).fit(X)
    with pytest.raises(ValueError):
        cdr.ArchetypalAnalysis(n_components=2, init='nope').fit(X)
    with pytest.raises(ValueError):
        cdr.KernelAA(n_components=2).fit(np.zeros((3, 4)))
    bad = np.full((2, 10), 0.3)
    with pytest.raises(ValueError):
        cdr.ArchetypalAnalysis(n_components=2, init='custom').fit_transform(
            X, dictionary=bad, weights=np.full((10, 2), 0.5), alpha=np.ones(2))


# ---------------------------------------------------------------- GPNH
def test_gpnh_single_functions_golden(golden):
    X, W0, Z0 = golden['gpnh/X'], golden['gpnh/W0'], golden['gpnh/Z0']
    T, d = X.shape
    k = W0.shape[1]
    close(gp._gpnh_regularization(W0), golden.scalar('gpnh/reg'), rtol=1e-12)
    close(gp._gpnh_cost(X, Z0, W0, 0.0), golden.scalar('gpnh/cost0'), rtol=1e-12)
    close(gp._gpnh_cost(X, Z0, W0, 3.2), golden.scalar('gpnh/cost_l'), rtol=1e-12)
    ZtZ, GW = Z0.T.dot(Z0), orc.gpnh_GW(d, k)
    for lam in (0.0, 3.2):
        W = gp._update_gpnh_dictionary(X, Z0, ZtZ, GW, lambda_W=lam)
        assert W.shape == (d, k)
        close(W, golden['gpnh/dict_update_l%g' % lam], rtol=1e-9, atol=1e-11)
    close(gp._update_gpnh_weights(X, Z0, W0, max_iterations=2),
          golden['gpnh/weights_update_it2'], rtol=1e-8, atol=1e-11)
    close(gp._update_gpnh_weights(X, Z0, W0), golden['gpnh/weights_update'], rtol=0, atol=5e-6)


def test_gpnh_exact_factorisation_has_zero_cost():
    # reference tests/test_gpnh_convex_coding.py:18-42
    rs = np.random.RandomState(0)
    Z = orc.right_stochastic_matrix((30, 3), rs)
    W = rs.uniform(size=(7, 3))
    X = Z.dot(W.T)
    assert abs(gp._gpnh_cost(X, Z, W, 0.0)) < 1e-12


@pytest.mark.parametrize('lam', [0.0, 3.2])
@pytest.mark.parametrize('wtag', ['full', 'w1'])
def test_gpnh_iterate_golden(golden, lam, wtag):
    X, W0, Z0 = golden['gpnh/X'], golden['gpnh/W0'], golden['gpnh/Z0']
    kw = {} if wtag == 'full' else dict(weights_solver_kwargs=dict(max_iterations=1))
    Z, W, cost, n_iter, avg_t, deltas = gp._iterate_gpnh_convex_coding(
        X, Z0.copy(), W0.copy(), lambda_W=lam, tolerance=1e-9, max_iterations=10, **kw)
    tag = 'gpnh/it_l%g_%s' % (lam, wtag)
    gcost, gn = golden[tag + '/stats']
    assert n_iter == int(gn)
    close(cost, gcost, rtol=1e-7)
    close(Z, golden[tag + '/Z'], rtol=0, atol=2e-5)
    close(W, golden[tag + '/W'], rtol=0, atol=2e-5)
    close(deltas, golden[tag + '/deltas'], rtol=1e-4, atol=1e-9)
    close(Z.sum(axis=1), 1.0, atol=1e-12)


def test_gpnh_frozen_dictionary_golden(golden):
    X, W0, Z0 = golden['gpnh/X'], golden['gpnh/W0'], golden['gpnh/Z0']
    Z, W, cost, n_iter, _, _ = gp._iterate_gpnh_convex_coding(
        X, Z0.copy(), W0.copy(), update_dictionary=False, tolerance=1e-9, max_iterations=4)
    assert np.array_equal(W, W0)
    close(Z, golden['gpnh/frozenW/Z'], rtol=0, atol=5e-6)
    close([cost, n_iter], golden['gpnh/frozenW/stats'], rtol=1e-7)
    with pytest.raises(TypeError):
        gp._iterate_gpnh_convex_coding(X, Z0.copy(), W0.copy(),
                                       dictionary_solver_kwargs=dict(max_iterations=1))


@pytest.mark.parametrize('init', ['random', 'furthest_sum'])
def test_gpnh_estimator_golden(golden, init):
    X = golden['gpnh/X']
    k = golden['gpnh/W0'].shape[1]
    m = cdr.GPNHConvexCoding(n_components=k, lambda_W=0.5, init=init, tolerance=1e-6,
                             max_iterations=30, random_state=0)
    Z = m.fit_transform(X)
    gcost, gn = golden['gpnh/est_%s/stats' % init]
    assert m.n_iter == int(gn)
    close(m.cost, gcost, rtol=1e-6)
    close(Z, golden['gpnh/est_%s/Z' % init], rtol=0, atol=1e-4)
    close(m.dictionary, golden['gpnh/est_%s/W' % init], rtol=0, atol=1e-4)
    close(m.cost_deltas, golden['gpnh/est_%s/deltas' % init], rtol=1e-3, atol=1e-8)
    Zv, cv = m.transform(golden['gpnh/est_%s/Xv' % init])
    close(Zv, golden['gpnh/est_%s/Zv' % init], rtol=0, atol=1e-4)
    close(cv, golden.scalar('gpnh/est_%s/cost_v' % init), rtol=1e-5)


def test_gpnh_mid_size_vs_oracle():
    """k = 8 on a 400 x 3000 surrogate: 6 alternating iterations against the oracle."""
    from convex_dim_red.datasets import synthetic_field
    X = synthetic_field(400, 3000, seed=3)
    rs = np.random.RandomState(1)
    W0, Z0 = orc.init_gpnh(X, 8, 'random', rs)
    ref = orc.iterate_gpnh(X, Z0.copy(), W0.copy(), lambda_W=0.1, tolerance=1e-12, max_iterations=6,
                           trace_XtX=float(np.sum(X * X)))
    got = gp._iterate_gpnh_convex_coding(X, Z0.copy(), W0.copy(), lambda_W=0.1, tolerance=1e-12,
                                         max_iterations=6)
    assert got[3] == ref[3]
    close(got[2], ref[2], rtol=1e-8)
    close(got[0], ref[0], rtol=0, atol=2e-5)
    close(got[1], ref[1], rtol=0, atol=2e-5)


def test_aa_mid_size_vs_oracle():
    from convex_dim_red.datasets import synthetic_field
    X = synthetic_field(300, 2500, seed=4)
    rs = np.random.RandomState(2)
    C0 = orc.right_stochastic_matrix((8, 300), rs)
    Z0 = orc.right_stochastic_matrix((300, 8), rs)
    kw = dict(tolerance=1e-12, max_iterations=6, dictionary_solver_kwargs=dict(max_iterations=1))
    ref = orc.iterate_aa(X, Z0.copy(), C0.copy(), np.ones(8), trace_XXt=float(np.sum(X * X)), **kw)
    got = aa._iterate_aa(X, Z0.copy(), C0.copy(), np.ones(8), **kw)
    assert got[4] == ref[4]
    close(got[3], ref[3], rtol=1e-8)
    close(got[0], ref[0], rtol=0, atol=2e-5)
    close(got[1], ref[1], rtol=0, atol=2e-5)


# ---------------------------------------------------------------- k-means
@pytest.mark.parametrize('tag,tol,max_iter', [('conv', 1e-4, 300), ('it2', 1e-4, 2), ('tol0', 0.0, 300)])
def test_kmeans_golden(golden, tag, tol, max_iter):
    X, picks = golden['km/X'], golden['km/picks']
    labels, centres, inertia, n_iter = kmeans_lloyd(X, X[picks], tol=tol, max_iter=max_iter)
    assert labels.dtype == np.int32
    assert np.array_equal(labels, golden['km/%s/labels' % tag])
    close(centres, golden['km/%s/centres' % tag], rtol=1e-10, atol=1e-12)
    close([inertia, n_iter], golden['km/%s/stats' % tag], rtol=1e-10)


def test_kmeans_empty_cluster_golden(golden):
    X = golden['km/X']
    labels, centres, inertia, n_iter = kmeans_lloyd(X, golden['km/empty/init'])
    assert np.array_equal(labels, golden['km/empty/labels'])
    close(centres, golden['km/empty/centres'], rtol=1e-10, atol=1e-12)
    close([inertia, n_iter], golden['km/empty/stats'], rtol=1e-10)


def test_kmeans_furthest_sum_init(golden):
    X = golden['km/X']
    picks = furthest_sum_centres(X, 5, 7, 10)
    assert np.array_equal(picks, golden['km/picks'])
    km = KMeans(n_clusters=5, init=X[picks]).fit(X)
    assert np.array_equal(km.labels_, golden['km/conv/labels'])
    assert np.array_equal(km.predict(X), km.labels_)


# ---------------------------------------------------------------- driver logic (restarts)
def test_fit_models_keep_the_best_restart(golden):
    from convex_dim_red import model_selection as ms
    X = golden['aa/X']
    train, val = ms.train_validation_split(X, 0.1)
    assert train.shape[0] == int(np.ceil(0.9 * X.shape[0])) and train.shape[0] + val.shape[0] == X.shape[0]
    best = ms.fit_aa_model(train, n_components=3, n_init=4, tolerance=1e-6, max_iterations=60,
                           random_state=0)
    # replay the restarts by hand from the same shared RNG (bin/run_hadisst_aa.py:154-172)
    rng = np.random.RandomState(0)
    costs = []
    for _ in range(4):
        m = cdr.ArchetypalAnalysis(n_components=3, init='random', tolerance=1e-6, max_iterations=60,
                                   random_state=rng, dictionary_solver_kwargs=dict(max_iterations=1))
        m.fit_transform(train)
        costs.append(m.cost)
    assert best.cost == min(costs)
    res = ms.evaluate_model(best, train, val)
    assert res['training_cost'] == best.cost
    assert res['validation_cost'] > 0 and res['validation_rmse'] > 0 and res['training_rmse'] > 0
    # FurthestSum restarts share one device Gram matrix while the data are resident
    fs = ms.fit_aa_model(train, n_components=3, init='furthest_sum', n_init=3, tolerance=1e-6,
                         max_iterations=40, random_state=2)
    rng = np.random.RandomState(2)
    fs_costs = []
    for _ in range(3):
        m = cdr.ArchetypalAnalysis(n_components=3, init='furthest_sum', tolerance=1e-6,
                                   max_iterations=40, random_state=rng,
                                   dictionary_solver_kwargs=dict(max_iterations=1))
        m.fit_transform(train)
        fs_costs.append(m.cost)
    assert fs.cost == min(fs_costs)
    g = ms.fit_gpnh_model(train, n_components=3, lambda_W=0.1, n_init=3, max_iterations=40,
                          random_state=1)
    assert g.dictionary.shape == (X.shape[1], 3) and g.cost > 0
    km = ms.fit_kmeans_model(golden['km/X'], n_components=5, init='furthest_sum', random_state=0)
    assert km.labels_.shape == (golden['km/X'].shape[0],) and km.inertia_ > 0
    folds = ms.time_series_cross_validate(ms.fit_gpnh_model, train, n_folds=3, n_components=2,
                                          n_init=1, max_iterations=20, random_state=0)
    assert len(folds) == 3 and all(f['test_cost'] > 0 for f in folds)


def test_resident_cache_is_transparent(golden):
    from convex_dim_red import model_selection as ms
    from convex_dim_red import _backend as be
    X = np.ascontiguousarray(golden['gpnh/X'])
    W0, Z0 = golden['gpnh/W0'], golden['gpnh/Z0']
    r1 = gp._iterate_gpnh_convex_coding(X, Z0.copy(), W0.copy(), tolerance=1e-9, max_iterations=5)
    with ms.resident(X):
        assert be.to_device_padded(X) is be.to_device_padded(X)
        r2 = gp._iterate_gpnh_convex_coding(X, Z0.copy(), W0.copy(), tolerance=1e-9, max_iterations=5)
    assert not be._RESIDENT
    assert np.array_equal(r1[0], r2[0]) and r1[2] == r2[2]


def test_aa_gram_formulation_matches_streaming(golden):
    """formulation='gram' (K = X X' once, then passes over K) follows the same trajectory as
    the default streaming formulation up to rounding."""
    X, K, C0, Z0, alpha = _aa_inputs(golden)
    kw = dict(tolerance=1e-9, max_iterations=12, dictionary_solver_kwargs=dict(max_iterations=1))
    a = aa._iterate_aa(X, Z0.copy(), C0.copy(), alpha.copy(), **kw)
    b = aa._iterate_aa(X, Z0.copy(), C0.copy(), alpha.copy(), formulation='gram', **kw)
    assert a[4] == b[4]
    close(b[3], a[3], rtol=1e-8)
    close(b[0], a[0], rtol=0, atol=2e-5)
    close(b[1], a[1], rtol=0, atol=2e-5)
    close(b[3], golden['aa/faa_d1/stats'][0], rtol=1e-7)
    m = cdr.ArchetypalAnalysis(n_components=4, init='random', tolerance=1e-6, max_iterations=40,
                               random_state=0, dictionary_solver_kwargs=dict(max_iterations=1),
                               formulation='gram')
    m.fit_transform(X)
    close(m.cost, golden['aa/est_random/stats'][0], rtol=1e-6)
    with pytest.raises(ValueError):
        aa._iterate_aa(X, Z0.copy(), C0.copy(), alpha.copy(), formulation='nope', **kw)


# ---------------------------------------------------------------- PCA and k-means++ ("next" rows)
def test_pca_matches_sklearn_full_svd():
    from sklearn.decomposition import PCA as SkPCA
    from convex_dim_red import PCA
    from convex_dim_red.datasets import synthetic_field
    X = synthetic_field(90, 700, seed=5) + 3.0          # non-zero column means
    for n in (5, 70):
        ours = PCA(n_components=n)
        scores = ours.fit_transform(X)
        ref = SkPCA(n_components=n, svd_solver='full')
        ref_scores = ref.fit_transform(X)
        close(ours.mean_, ref.mean_, rtol=1e-13)
        close(ours.singular_values_, ref.singular_values_, rtol=1e-9)
        close(ours.explained_variance_, ref.explained_variance_, rtol=1e-9)
        close(ours.explained_variance_ratio_, ref.explained_variance_ratio_, rtol=1e-9)
        close(ours.components_, ref.components_, rtol=0, atol=1e-8)
        close(scores, ref_scores, rtol=0, atol=1e-7 * np.abs(ref_scores).max())
        Xn = synthetic_field(11, 700, seed=6) + 3.0
        close(ours.transform(Xn), ref.transform(Xn), rtol=0, atol=1e-7 * np.abs(ref_scores).max())
        close(ours.inverse_transform(scores), ref.inverse_transform(ref_scores), rtol=0, atol=1e-7)
    with pytest.raises(ValueError):
        PCA(n_components=500).fit(X)


def test_kmeans_plusplus_matches_sklearn(golden):
    from sklearn.cluster import KMeans as SkKMeans, kmeans_plusplus as sk_kmeans_plusplus
    from convex_dim_red import kmeans_plusplus
    X = golden['km/X']
    for seed in (0, 1, 2):
        c_ref, i_ref = sk_kmeans_plusplus(X, 5, random_state=np.random.RandomState(seed))
        c, i = kmeans_plusplus(X, 5, random_state=np.random.RandomState(seed))
        assert np.array_equal(i, i_ref)
        assert np.array_equal(c, c_ref)
    for init in ('k-means++', 'random'):
        ref = SkKMeans(n_clusters=5, init=init, n_init=3, random_state=3, algorithm='lloyd').fit(X.copy())
        km = KMeans(n_clusters=5, init=init, n_init=3, random_state=3).fit(X)
        assert np.array_equal(km.labels_, ref.labels_)
        close(km.inertia_, ref.inertia_, rtol=1e-10)
        close(km.cluster_centers_, ref.cluster_centers_, rtol=1e-10, atol=1e-12)


def test_pca_reduced_sweep_runs_config4_workflow():
    """PCA -> AA / k-means sweep on a small JRA-55-like field (BASELINE.json configs[3])."""
    from convex_dim_red import model_selection as ms
    from convex_dim_red.datasets import synthetic_field
    X = synthetic_field(120, 900, seed=8)
    pca, models = ms.pca_reduced_sweep(X, n_eofs=20, component_counts=(3, 5), method='aa',
                                       n_init=2, tolerance=1e-6, max_iterations=300, random_state=0)
    assert pca.components_.shape == (20, 900) and sorted(models) == [3, 5]
    assert models[5].cost <= models[3].cost + 1e-9         # more archetypes never fit worse here
    for k, m in models.items():
        assert m.weights.shape == (120, k) and np.allclose(m.weights.sum(axis=1), 1, 1e-12)
        assert m.archetypes.shape == (k, 20)
        # archetypes map back to the grid through the EOFs
        assert pca.inverse_transform(m.archetypes).shape == (k, 900)
    _, km = ms.pca_reduced_sweep(X, n_eofs=10, component_counts=(4,), method='kmeans', random_state=1)
    assert km[4].labels_.shape == (120,)


# ---------------------------------------------------------------- degenerate shapes
@pytest.mark.parametrize('T,d,k', [(20, 7, 1), (5, 3, 2), (2, 1, 1), (9, 40, 8), (33, 2, 3)])
def test_tiny_and_degenerate_shapes_match_oracle(T, d, k):
    """One component, fewer samples than a DMMA tile, a single feature: every kernel must
    handle ragged / minimal shapes exactly like the reference algorithm."""
    rs = np.random.RandomState(T * 100 + d * 10 + k)
    X = rs.standard_normal((T, d))
    W0 = 0.5 * rs.standard_normal((d, k))
    Z0 = orc.right_stochastic_matrix((T, k), rs)
    C0 = orc.right_stochastic_matrix((k, T), rs)
    ref = orc.iterate_gpnh(X, Z0.copy(), W0.copy(), lambda_W=0.3, tolerance=1e-12, max_iterations=4)
    got = gp._iterate_gpnh_convex_coding(X, Z0.copy(), W0.copy(), lambda_W=0.3, tolerance=1e-12,
                                         max_iterations=4)
    assert got[3] == ref[3]
    close(got[2], ref[2], rtol=1e-7, atol=1e-12)
    close(got[0], ref[0], rtol=0, atol=2e-5)
    close(got[1], ref[1], rtol=0, atol=2e-5)
    kw = dict(tolerance=1e-12, max_iterations=4, dictionary_solver_kwargs=dict(max_iterations=2))
    ref = orc.iterate_aa(X, Z0.copy(), C0.copy(), np.ones(k), **kw)
    got = aa._iterate_aa(X, Z0.copy(), C0.copy(), np.ones(k), **kw)
    assert got[4] == ref[4]
    close(got[3], ref[3], rtol=1e-7, atol=1e-12)
    close(got[0], ref[0], rtol=0, atol=2e-5)
    close(got[1], ref[1], rtol=0, atol=2e-5)
    ref = orc.iterate_kernel_aa(X.dot(X.T), Z0.copy(), C0.copy(), np.ones(k), **kw)
    got = aa._iterate_kernel_aa(X.dot(X.T), Z0.copy(), C0.copy(), np.ones(k), **kw)
    close(got[3], ref[3], rtol=1e-7, atol=1e-12)
    close(got[1], ref[1], rtol=0, atol=2e-5)
    if T > k:
        labels, centres, inertia, n_iter = kmeans_lloyd(X, X[:k].copy())
        rl, rc, ri, rn = orc.kmeans_lloyd(X, X[:k].copy())
        assert np.array_equal(labels, rl) and n_iter == rn
        close(centres, rc, rtol=1e-10, atol=1e-12)


def test_gap_statistic_follows_the_reference_protocol():
    """gap_statistic (kmeans.py:81-108) against the reference's own code path with
    scikit-learn's KMeans in place (the reference passes the removed `n_jobs` to KMeans, so its
    body is restated here with that argument dropped): same per-trial seeds, same reference
    data, same k-means++ seeding and Lloyd iterations -> the same dispersions."""
    from sklearn.cluster import KMeans as SkKMeans
    from sklearn.utils import check_random_state
    from convex_dim_red import kmeans as km
    rs = np.random.RandomState(3)
    centres = rs.standard_normal((3, 20)) * 4
    X = centres[rs.randint(3, size=120)] + rs.standard_normal((120, 20))
    k, n_trials = 3, 3
    model = km.KMeans(n_clusters=k, init='k-means++', n_init=2, random_state=0).fit(X)
    gap, sk = cdr.gap_statistic(X, model.inertia_, n_components=k, n_trials=n_trials,
                                reference='uniform', n_jobs=1, random_state=5)

    rng = check_random_state(5)
    seeds = []
    for _ in range(n_trials):
        while True:
            seed = rng.randint(np.iinfo(np.int32).max)
            if seed not in seeds:
                seeds.append(seed)
                break
    wk = []
    for seed in seeds:
        r = check_random_state(seed)
        lo = np.broadcast_to(X.min(axis=0), X.shape)
        hi = np.broadcast_to(X.max(axis=0), X.shape)
        data = (hi - lo) * r.uniform(size=X.shape) + lo
        wk.append(SkKMeans(n_clusters=k, n_init=10, random_state=r).fit(data).inertia_)
    ln = np.log(np.array(wk))
    np.testing.assert_allclose(gap, ln.mean() - np.log(model.inertia_), rtol=1e-9)
    np.testing.assert_allclose(sk, np.std(ln) * np.sqrt(1 + 1.0 / n_trials), rtol=1e-6, atol=1e-12)
