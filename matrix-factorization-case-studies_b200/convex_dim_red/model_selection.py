"""Restart / selection logic of the reference's drivers as a NumPy-in / NumPy-out API.

The reference keeps this logic in its ``bin/`` scripts (``fit_aa_model``
``bin/run_hadisst_aa.py:149-174``, ``fit_gpnh_model`` ``bin/run_hadisst_gpnh.py:149-171``,
``fit_kmeans_model`` ``bin/run_hadisst_kmeans.py:120-137``, the train / validation split
``bin/run_hadisst_aa.py:204-209`` and the ``TimeSeriesSplit`` cross-validation ``:213-244``)
wrapped in xarray / netCDF I/O.  Here the same steps work on plain arrays; the data matrix
is uploaded to the GPU once and shared by all restarts (``resident``).
"""

from copy import deepcopy

import numpy as np
from sklearn.utils import check_random_state

from . import _backend as be
from .archetypal_analysis import ArchetypalAnalysis
from .gpnh_convex_coding import GPNHConvexCoding
from .kmeans import KMeans


def resident(*arrays):
    """Context manager keeping the padded device copies of ``arrays`` alive, so repeated
    fits on the same (unmodified) array skip the host-to-device copy."""
    return be.DeviceCache(arrays)


def train_validation_split(data, validation_frac=0.1):
    """First ceil((1 - validation_frac) n) rows for training, the rest for validation
    (bin/run_hadisst_aa.py:204-209)."""
    n_samples = data.shape[0]
    n_training = int(np.ceil((1 - validation_frac) * n_samples))
    return data[:n_training], data[n_training:]


def root_mean_squared_error(a, b):
    """``sklearn.metrics.mean_squared_error(a, b, squared=False)``: per-column RMSE,
    uniformly averaged (as used at bin/run_hadisst_aa.py:233-244)."""
    return float(np.mean(np.sqrt(np.mean((np.asarray(a) - np.asarray(b)) ** 2, axis=0))))


def fit_aa_model(X, n_components=2, delta=0, init='random', n_init=100,
                 tolerance=1e-6, max_iterations=10000, verbose=False,
                 random_state=None, **kwargs):
    """Run archetypal analysis ``n_init`` times from one shared RNG and keep the fit with
    the lowest cost (bin/run_hadisst_aa.py:149-174)."""
    rng = check_random_state(random_state)
    kwargs.setdefault('dictionary_solver_kwargs', dict(max_iterations=1))
    min_cost = None
    best_model = None
    with resident(X):
        for _ in range(n_init):
            model = ArchetypalAnalysis(
                n_components=n_components, delta=delta, init=init, tolerance=tolerance,
                max_iterations=max_iterations, verbose=verbose, random_state=rng, **kwargs)
            model.fit_transform(X)
            if min_cost is None or model.cost < min_cost:
                best_model = deepcopy(model)
                min_cost = model.cost
    return best_model


def fit_gpnh_model(X, n_components=2, lambda_W=0, init='random', n_init=100,
                   tolerance=1e-6, max_iterations=10000, verbose=False,
                   random_state=None, **kwargs):
    """GPNH counterpart (bin/run_hadisst_gpnh.py:149-171)."""
    rng = check_random_state(random_state)
    min_cost = None
    best_model = None
    with resident(X):
        for _ in range(n_init):
            model = GPNHConvexCoding(
                n_components=n_components, lambda_W=lambda_W, init=init, tolerance=tolerance,
                max_iterations=max_iterations, verbose=verbose, random_state=rng, **kwargs)
            model.fit_transform(X)
            if min_cost is None or model.cost < min_cost:
                best_model = deepcopy(model)
                min_cost = model.cost
    return best_model


def fit_kmeans_model(X, n_components=2, init='furthest_sum', n_init=1, tolerance=1e-4,
                     max_iterations=10000, verbose=False, random_state=None):
    """k-means with restarts, lowest inertia kept (bin/run_hadisst_kmeans.py:120-137)."""
    rng = check_random_state(random_state)
    return KMeans(n_clusters=n_components, init=init, n_init=n_init, tol=tolerance,
                  max_iter=max_iterations, verbose=verbose, random_state=rng).fit(X)


def pca_reduced_sweep(X, n_eofs, component_counts, method='aa', n_init=1, random_state=None,
                      **fit_kwargs):
    """PCA reduction followed by a sweep over the number of components -- the
    ``bin/run_jra55_pca_{aa,gpnh,kmeans}.py`` workflow (config 4 of BASELINE.json): the field
    is reduced to ``n_eofs`` principal components once, then one model per k is fitted on
    the scores (both AA solvers limited to one inner iteration and the ``rel_delta_f``
    stopping rule, as ``bin/run_jra55_pca_aa.py:119-133`` does).
    Returns ``(pca, {k: model})``."""
    from .pca import PCA
    rng = check_random_state(random_state)
    pca = PCA(n_components=n_eofs)
    scores = np.ascontiguousarray(pca.fit_transform(X))
    models = {}
    for k in component_counts:
        if method == 'aa':
            kw = dict(stopping_criterion='rel_delta_f',
                      dictionary_solver_kwargs=dict(max_iterations=1),
                      weights_solver_kwargs=dict(max_iterations=1))
            kw.update(fit_kwargs)
            models[k] = fit_aa_model(scores, n_components=k, n_init=n_init, random_state=rng, **kw)
        elif method == 'gpnh':
            kw = dict(stopping_criterion='rel_delta_f',
                      weights_solver_kwargs=dict(max_iterations=1))
            kw.update(fit_kwargs)
            models[k] = fit_gpnh_model(scores, n_components=k, n_init=n_init, random_state=rng,
                                       **kw)
        elif method == 'kmeans':
            models[k] = fit_kmeans_model(scores, n_components=k, n_init=n_init,
                                         random_state=rng, **fit_kwargs)
        else:
            raise ValueError("method must be 'aa', 'gpnh' or 'kmeans'")
    return pca, models


def evaluate_model(model, training_data, validation_data=None):
    """Training / validation cost and RMSE of a fitted AA or GPNH model
    (bin/run_hadisst_aa.py:226-244, 356-385)."""
    out = {'training_cost': float(model.cost),
           'training_rmse': root_mean_squared_error(
               training_data, model.inverse_transform(model.weights))}
    if validation_data is not None and len(validation_data):
        saved = model.weights
        weights, cost = model.transform(validation_data)
        out['validation_cost'] = float(cost)
        out['validation_rmse'] = root_mean_squared_error(
            validation_data, model.inverse_transform(weights))
        model.weights = saved
    return out


def time_series_cross_validate(fit_model, data, n_folds=10, **fit_kwargs):
    """``TimeSeriesSplit`` cross-validation of ``fit_model`` (bin/run_hadisst_aa.py:213-244).
    Returns the per-fold dictionaries of :func:`evaluate_model`."""
    from sklearn.model_selection import TimeSeriesSplit
    folds = []
    for train, test in TimeSeriesSplit(n_splits=n_folds).split(data):
        model = fit_model(data[train], **fit_kwargs)
        res = evaluate_model(model, data[train], data[test])
        folds.append({'training_cost': res['training_cost'], 'training_rmse': res['training_rmse'],
                      'test_cost': res['validation_cost'], 'test_rmse': res['validation_rmse']})
    return folds
