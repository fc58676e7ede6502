"""Restart / selection logic of the reference's drivers as a NumPy-in / NumPy-out API.

The reference keeps this logic in its ``bin/`` scripts (``fit_aa_model``
``bin/run_hadisst_aa.py:149-174``, ``fit_gpnh_model`` ``bin/run_hadisst_gpnh.py:149-171``,
``fit_kmeans_model`` ``bin/run_hadisst_kmeans.py:120-137``, the train / validation split
``bin/run_hadisst_aa.py:204-209`` and the ``TimeSeriesSplit`` cross-validation ``:213-244``)
wrapped in xarray / netCDF I/O.  Here the same steps work on plain arrays; the data matrix
is uploaded to the GPU once and shared by all restarts (``resident``).

Restarts are independent, so with a process group (``comm=Comm()``, one process per GPU) they
run as replicas: restart i is fitted by rank ``i % world`` and the others only advance the
shared ``RandomState`` by the draws that restart's initialisation makes, which keeps every
initial matrix identical to the single-process sequence ("draw all
initial matrices on the host from the single shared RandomState in the reference's order,
then farm restarts out to GPUs").  The fit with the lowest cost -- the first one in restart
order on ties, as the serial strict ``<`` comparison keeps it -- is broadcast to all ranks.
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


def _serial_winner(costs):
    """Index the serial loop ``if best is None or cost < min_cost`` ends up keeping."""
    winner = None
    for index, cost in enumerate(costs):
        if winner is None or cost < costs[winner]:
            winner = index
    return winner


def best_of_restarts(n_init, fit_one, skip_one=None, comm=None):
    """Fit ``n_init`` restarts and return the model with the lowest ``cost``.

    ``fit_one()`` fits one restart (consuming the shared RNG); ``skip_one()`` advances the
    RNG by exactly the draws ``fit_one`` would make without fitting.  Without a process group
    this is the reference's loop (bin/run_hadisst_aa.py:154-172); with one, restart i runs on
    rank ``i % world`` and every rank returns the same winning model.
    """
    world, rank = (comm.world, comm.rank) if comm is not None and comm.enabled else (1, 0)
    if world > 1 and skip_one is None:
        raise ValueError('replicated restarts need skip_one to keep the RNG sequence')
    mine = {}                 # restart index -> cost, for the restarts fitted here
    best_index, best_model = None, None
    for index in range(n_init):
        if index % world != rank:
            skip_one()
            continue
        model = fit_one()
        mine[index] = model.cost
        if best_index is None or model.cost < mine[best_index]:
            best_model = deepcopy(model)
            best_index = index
    if world == 1:
        return best_model
    costs = [None] * n_init
    for part in comm.allgather_objects(mine):
        for index, cost in part.items():
            costs[index] = cost
    winner = _serial_winner(costs)
    if winner is None:
        return None
    owner = winner % world
    if rank == owner and best_index != winner:
        raise RuntimeError('restart selection diverged between ranks')
    return comm.broadcast_object(best_model if rank == owner else None, owner)


def _skip_aa_initialisation(rng, init, n_samples, n_components, delta):
    """The RNG draws of one ``ArchetypalAnalysis`` initialisation, discarded: dictionary
    (uniform k x T, or one start index for FurthestSum), weights (uniform T x k), scale
    factors when delta != 0 (archetypal_analysis.py:51-164 of the reference)."""
    if init == 'random':
        rng.uniform(size=(n_components, n_samples))
    elif init in (None, 'furthest_sum'):
        rng.randint(n_samples)
    else:
        raise ValueError('replicated restarts support init = random or furthest_sum')
    rng.uniform(size=(n_samples, n_components))
    if delta != 0:
        rng.uniform(low=(1 - delta), high=(1 + delta), size=(n_components,))


def _skip_gpnh_initialisation(rng, init, n_samples, n_features, n_components):
    """GPNH counterpart: normal d x k dictionary (or a start index), then uniform weights
    (gpnh_convex_coding.py:41-143 of the reference)."""
    if init in (None, 'random'):
        rng.randn(n_features, n_components)
    elif init == 'furthest_sum':
        rng.randint(n_samples)
    else:
        raise ValueError('replicated restarts support init = random or furthest_sum')
    rng.uniform(size=(n_samples, n_components))


def fit_aa_model(X, n_components=2, delta=0, init='random', n_init=100,
                 tolerance=1e-6, max_iterations=10000, verbose=False,
                 random_state=None, comm=None, **kwargs):
    """Run archetypal analysis ``n_init`` times from one shared RNG and keep the fit with
    the lowest cost (bin/run_hadisst_aa.py:149-174).  ``comm`` spreads the restarts over
    the ranks of a process group; every rank passes the same X and seed."""
    rng = check_random_state(random_state)
    kwargs.setdefault('dictionary_solver_kwargs', dict(max_iterations=1))
    n_samples = np.shape(X)[0]

    def fit_one():
        model = ArchetypalAnalysis(
            n_components=n_components, delta=delta, init=init, tolerance=tolerance,
            max_iterations=max_iterations, verbose=verbose, random_state=rng, **kwargs)
        model.fit_transform(X)
        return model

    def skip_one():
        _skip_aa_initialisation(rng, init, n_samples, n_components, delta)

    with resident(X):
        return best_of_restarts(n_init, fit_one, skip_one, comm)


def fit_gpnh_model(X, n_components=2, lambda_W=0, init='random', n_init=100,
                   tolerance=1e-6, max_iterations=10000, verbose=False,
                   random_state=None, comm=None, **kwargs):
    """GPNH counterpart (bin/run_hadisst_gpnh.py:149-171)."""
    rng = check_random_state(random_state)
    n_samples, n_features = np.shape(X)

    def fit_one():
        model = GPNHConvexCoding(
            n_components=n_components, lambda_W=lambda_W, init=init, tolerance=tolerance,
            max_iterations=max_iterations, verbose=verbose, random_state=rng, **kwargs)
        model.fit_transform(X)
        return model

    def skip_one():
        _skip_gpnh_initialisation(rng, init, n_samples, n_features, n_components)

    with resident(X):
        return best_of_restarts(n_init, fit_one, skip_one, comm)


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
