"""The drop-in boundary is the reference's Python API: names,
positional order and default values of the public callables, checked without a GPU.
Expected signatures are transcribed from the reference (file:line in each case)."""

import importlib
import inspect

import numpy as np
import pytest

import convex_dim_red as cdr
from convex_dim_red import archetypal_analysis as aa
from convex_dim_red import gpnh_convex_coding as gp
from convex_dim_red import simplex_projection as sp

# the package re-exports the function `spg`, which shadows the module attribute
spg_mod = importlib.import_module('convex_dim_red.spg')


def params(fn):
    return [(n, p.default) for n, p in inspect.signature(fn).parameters.items()
            if n != 'self' and p.kind not in (p.VAR_KEYWORD, p.VAR_POSITIONAL)]


E = inspect.Parameter.empty


def test_estimator_constructors():
    # archetypal_analysis.py:743-745, 999-1001; gpnh_convex_coding.py:476-479
    want_aa = [('n_components', E), ('delta', 0), ('init', None), ('tolerance', 1e-6),
               ('max_iterations', 1000), ('verbose', 0), ('random_state', None)]
    assert params(cdr.ArchetypalAnalysis.__init__) == want_aa
    assert params(cdr.KernelAA.__init__) == want_aa
    assert params(cdr.GPNHConvexCoding.__init__) == [
        ('n_components', E), ('lambda_W', 0), ('init', None), ('tolerance', 1e-6),
        ('max_iterations', 1000), ('verbose', 0), ('random_state', None)]
    for cls in (cdr.ArchetypalAnalysis, cdr.KernelAA, cdr.GPNHConvexCoding):
        sig = inspect.signature(cls.__init__)
        assert any(p.kind == p.VAR_KEYWORD for p in sig.parameters.values())
        m = cls(3, weights_solver_kwargs=dict(max_iterations=5), stopping_criterion='rel_delta_f',
                require_monotonic_cost_decrease=False)
        assert m.weights_solver_kwargs == dict(max_iterations=5)
        assert m.stopping_criterion == 'rel_delta_f' and m.require_monotonic_cost_decrease is False
        assert m.weights is None and m.dictionary is None and m.cost == 0 and m.n_iter == 0
        assert m.avg_time_per_iter == 0 and m.cost_deltas is None


def test_estimator_methods():
    # archetypal_analysis.py:859, 895, 1108, 1151, 1201; gpnh_convex_coding.py:574, 605, 623, 654
    assert params(cdr.ArchetypalAnalysis.fit_transform) == [
        ('data', E), ('dictionary', None), ('weights', None), ('alpha', None)]
    assert params(cdr.KernelAA.fit_transform) == [
        ('data', E), ('dictionary', None), ('weights', None), ('alpha', None)]
    assert params(cdr.GPNHConvexCoding.fit_transform) == [
        ('data', E), ('dictionary', None), ('weights', None)]
    for cls in (cdr.ArchetypalAnalysis, cdr.GPNHConvexCoding):
        assert params(cls.transform) == [('data', E)]
        assert params(cls.inverse_transform) == [('weights', E)]
        assert hasattr(cls, 'fit')
    assert hasattr(cdr.KernelAA, 'fit') and not hasattr(cdr.KernelAA, 'transform')
    assert cdr.ArchetypalAnalysis(2).archetypes is None


def test_solver_and_helper_signatures():
    # spg.py:46-51
    assert params(cdr.spg) == [
        ('f', E), ('df', E), ('x0', E), ('project', None), ('gamma', 1e-4), ('memory', 1),
        ('sigma_one', 0.1), ('sigma_two', 0.9), ('lambda_min', 1e-10), ('alpha0', None),
        ('alpha_min', 1e-5), ('alpha_max', 1e3), ('epsilon_one', 1e-10), ('epsilon_two', 1e-6),
        ('use_infinity_norm', True), ('verbose', 0), ('max_iterations', 10000),
        ('max_feval', 1000000)]
    # spg.py:286-291
    assert params(spg_mod.quad_simplex_spg) == [
        ('A', E), ('b', E), ('x0', E), ('gamma', 1e-4), ('memory', 1), ('sigma_one', 0.1),
        ('sigma_two', 0.9), ('lambda_min', 1e-10), ('alpha0', -1.0), ('alpha_min', 1e-5),
        ('alpha_max', 1e3), ('epsilon_one', 1e-10), ('epsilon_two', 1e-6),
        ('max_iterations', 1000), ('max_feval', 2000)]
    # spg.py:19-20, 36
    assert params(spg_mod.spg_line_search_step_length) == [
        ('current_step_length', E), ('delta', E), ('f_old', E), ('f_new', E),
        ('sigma_one', 0.1), ('sigma_two', 0.9)]
    assert params(spg_mod.spg_line_search_cauchy_step_size) == [
        ('beta', E), ('sksk', E), ('alpha_min', 1e-3), ('alpha_max', 1e3)]
    # furthest_sum.py:130-131
    assert params(cdr.furthest_sum) == [
        ('dissimilarity_matrix', E), ('n_components', E), ('start_index', E),
        ('exclude', None), ('extra_steps', 1)]
    # stochastic_matrices.py:32, 37
    for fn in (cdr.left_stochastic_matrix, cdr.right_stochastic_matrix):
        assert params(fn) == [('shape', E), ('random_state', None)]
    assert [n for n, _ in params(sp.simplex_project_vector)] == ['x']


def test_gap_statistic_signature_is_the_reference_one():
    # kmeans.py:81-82; the drivers call gap_statistic(X, model.inertia_, n_components=k,
    # n_trials=..., reference=..., n_jobs=..., random_state=...) (bin/run_jra55_kmeans.py:128)
    from convex_dim_red import kmeans as km
    assert params(cdr.gap_statistic) == [
        ('X', E), ('Wk', E), ('n_components', E), ('n_trials', 100), ('reference', 'uniform'),
        ('n_jobs', 1), ('random_state', None)]
    assert params(km._calculate_uniform_reference_wk) == [
        ('X', E), ('n_clusters', E), ('n_init', 10), ('n_jobs', None), ('random_state', None)]
    assert params(km._calculate_pca_reference_wk) == [
        ('X', E), ('n_clusters', E), ('n_init', 10), ('n_components', 100), ('n_iter', 10),
        ('n_jobs', None), ('random_state', None)]
    # the driver's call pattern binds
    inspect.signature(cdr.gap_statistic).bind(np.zeros((4, 3)), 1.5, n_components=2, n_trials=3,
                                              reference='pca', n_jobs=2, random_state=0)
    with pytest.raises(ValueError, match='unrecognized reference distribution'):
        km._calculate_reference_wk(np.zeros((4, 3)), 2, reference='box')


def test_private_helpers_used_by_the_reference_tests():
    # tests/test_archetypal_analysis.py:14-18, tests/test_gpnh_convex_coding.py:13-15
    assert params(aa._kernel_aa_cost) == [('K', E), ('weights', E), ('dictionary', E), ('alpha', E)]
    assert params(aa._update_kernel_aa_dictionary) == [
        ('K', E), ('dictionary', E), ('alpha', E), ('trace_K', E), ('KZ', E), ('ZtZ', E)]
    assert params(aa._update_aa_dictionary) == [
        ('X', E), ('dictionary', E), ('alpha', E), ('trace_XXt', E), ('XXtZ', E), ('ZtZ', E)]
    assert params(aa._update_kernel_aa_weights) == [
        ('weights', E), ('alpha', E), ('CK', E), ('CKCt', E)]
    want_iter = [('weights', E), ('dictionary', E), ('alpha', E), ('delta', 0),
                 ('update_weights', True), ('update_dictionary', True),
                 ('update_scale_factors', True), ('tolerance', 1e-6), ('max_iterations', 1000),
                 ('verbose', 0)]
    assert params(aa._iterate_kernel_aa) == [('K', E)] + want_iter
    assert params(aa._iterate_aa) == [('X', E)] + want_iter
    assert params(gp._gpnh_cost) == [('data', E), ('weights', E), ('dictionary', E), ('lambda_W', 0)]
    assert params(gp._update_gpnh_dictionary) == [
        ('X', E), ('weights', E), ('ZtZ', E), ('GW', E), ('lambda_W', 0)]
    assert params(gp._update_gpnh_weights) == [('X', E), ('weights', E), ('dictionary', E)]
    assert params(gp._iterate_gpnh_convex_coding) == [
        ('X', E), ('weights', E), ('dictionary', E), ('lambda_W', 0), ('update_weights', True),
        ('update_dictionary', True), ('tolerance', 1e-6), ('max_iterations', 1000), ('verbose', 0)]
    assert aa.INITIALIZATION_METHODS == (None, 'random', 'furthest_sum',)
    assert gp.INITIALIZATION_METHODS == (None, 'random', 'furthest_sum',)


def test_parameter_validation_happens_before_any_device_work():
    """Bad parameters raise the reference's ValueErrors even on a box without a GPU
    (archetypal_analysis.py:1035-1046, gpnh_convex_coding.py:509-521)."""
    X = np.random.RandomState(0).uniform(size=(6, 4))
    with pytest.raises(ValueError, match='Number of components must be a positive integer'):
        cdr.GPNHConvexCoding(n_components=0).fit(X)
    with pytest.raises(ValueError, match='Maximum number of iterations must be a positive'):
        cdr.GPNHConvexCoding(n_components=2, max_iterations=-3).fit(X)
    with pytest.raises(ValueError, match='Tolerance for stopping criteria must be'):
        cdr.GPNHConvexCoding(n_components=2, tolerance=-1.0).fit(X)
    with pytest.raises(ValueError, match='Expected square kernel matrix'):
        cdr.KernelAA(n_components=2).fit(np.zeros((3, 4)))
    with pytest.raises(ValueError, match='Number of components must be a positive integer'):
        cdr.KernelAA(n_components=-1).fit(np.eye(3))
    with pytest.raises(ValueError, match='Dissimilarity matrix must be square'):
        cdr.furthest_sum(np.zeros((3, 4)), 2, 0)
    assert cdr.furthest_sum(np.zeros((3, 3)), 0, 0) == []


def test_skipped_initialisation_takes_the_same_rng_draws(monkeypatch):
    """Replicated restarts (model_selection.best_of_restarts) skip the fits of other ranks
    by discarding the RNG draws their initialisation would take.  The skip functions must
    advance the RandomState exactly like the real initialisers; FurthestSum's device work is
    stubbed out (it consumes no random numbers)."""
    from convex_dim_red import model_selection as ms
    T, d, k = 23, 31, 4
    X = np.random.RandomState(0).randn(T, d)

    class Shape:
        shape = (T, T)

    monkeypatch.setattr(aa, '_kernel_device', lambda kernel: None)
    monkeypatch.setattr(aa, 'dissimilarity_from_gram_device', lambda K, n: None)
    monkeypatch.setattr(aa, 'furthest_sum_device', lambda *a, **kw: np.arange(k))
    monkeypatch.setattr(gp, 'dissimilarity_from_gram_device', lambda K, n: None)
    monkeypatch.setattr(gp, 'furthest_sum_device', lambda *a, **kw: np.arange(k))
    monkeypatch.setattr(gp.be, 'to_device_padded', lambda a: None)
    monkeypatch.setattr(gp.be, 'gram', lambda *a: None)

    def state(rng):
        st = rng.get_state()
        return st[1].tobytes(), st[2], st[3], st[4]

    for init in ('random', 'furthest_sum', None):
        for delta in (0, 0.3):
            real, skip = np.random.RandomState(3), np.random.RandomState(3)
            aa._initialize_kernel_aa(Shape(), k, init=init, random_state=real)
            aa._initialize_kernel_aa_scale_factors_random(k, delta=delta, random_state=real)
            ms._skip_aa_initialisation(skip, init, T, k, delta)
            assert state(real) == state(skip), ('aa', init, delta)
        real, skip = np.random.RandomState(3), np.random.RandomState(3)
        gp._initialize_gpnh_convex_coding(X, k, init=init, random_state=real)
        ms._skip_gpnh_initialisation(skip, init, T, d, k)
        assert state(real) == state(skip), ('gpnh', init)
    with pytest.raises(ValueError):
        ms._skip_aa_initialisation(np.random.RandomState(0), 'custom', T, k, 0)


def test_best_of_restarts_keeps_the_first_minimum():
    from convex_dim_red.model_selection import best_of_restarts, _serial_winner

    class M:
        def __init__(self, cost, tag):
            self.cost, self.tag = cost, tag

    seq = iter([M(3.0, 'a'), M(1.0, 'b'), M(1.0, 'c'), M(2.0, 'd')])
    assert best_of_restarts(4, lambda: next(seq)).tag == 'b'
    assert _serial_winner([3.0, 1.0, 1.0, 2.0]) == 1
    assert _serial_winner([float('nan'), 1.0]) == 0        # the reference loop never replaces NaN
    assert _serial_winner([2.0, float('nan'), 1.0]) == 2
    assert best_of_restarts(0, lambda: None) is None
