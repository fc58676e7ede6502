"""Array-level versions of the pre-processing the reference's drivers do with xarray before
the hot path (``bin/run_hadisst_aa.py:112-146, 196-202``): latitude weighting, flattening of
(time, lat, lon) fields to (samples, features), removal of features with missing values, and
the inverse mapping of feature-space results (archetypes, dictionaries, centres) back onto
the grid.  Host-side NumPy only; netCDF I/O stays with the caller.
"""

import numpy as np


def latitude_weights(latitudes, lat_weights='scos'):
    """'cos', 'scos' (square root of cos, the drivers' default) or 'none'
    (bin/run_hadisst_aa.py:112-124)."""
    lat = np.asarray(latitudes, dtype=np.float64)
    if lat_weights == 'cos':
        return np.cos(np.deg2rad(lat)).clip(0.0, 1.0)
    if lat_weights == 'scos':
        return np.cos(np.deg2rad(lat)).clip(0.0, 1.0) ** 0.5
    if lat_weights == 'none':
        return np.ones_like(lat)
    raise ValueError("Invalid weights descriptor '%r'" % lat_weights)


def weight_and_flatten(field, weights=None):
    """(n_samples, n_lat, n_lon) -> (n_samples, n_lat * n_lon), each latitude row scaled by
    its weight (bin/run_hadisst_aa.py:127-146)."""
    field = np.asarray(field, dtype=np.float64)
    if field.ndim != 3:
        raise ValueError('expected a (time, lat, lon) array, got %d dimensions' % field.ndim)
    if weights is not None:
        weights = np.asarray(weights, dtype=np.float64)
        if weights.shape != (field.shape[1],):
            raise ValueError('expected %d latitude weights, got shape %s'
                             % (field.shape[1], weights.shape))
        field = field * weights[np.newaxis, :, np.newaxis]
    return np.ascontiguousarray(field.reshape(field.shape[0], -1))


def drop_missing_features(flat_data):
    """Remove every column containing a NaN (land / ice cells).  Returns
    ``(valid_data, missing_mask)`` (bin/run_hadisst_aa.py:198-200)."""
    flat_data = np.asarray(flat_data, dtype=np.float64)
    missing = np.any(np.isnan(flat_data), axis=0)
    return np.ascontiguousarray(flat_data[:, np.logical_not(missing)]), missing


def restore_features(values, missing_mask, grid_shape=None, fill_value=np.nan):
    """Scatter (n_rows, n_valid_features) results back to all features (NaN where the
    feature was dropped) and optionally reshape to (n_rows, n_lat, n_lon)
    (bin/run_hadisst_aa.py:330-345)."""
    values = np.atleast_2d(np.asarray(values, dtype=np.float64))
    missing_mask = np.asarray(missing_mask, dtype=bool)
    full = np.full((values.shape[0], missing_mask.size), fill_value)
    full[:, np.logical_not(missing_mask)] = values
    if grid_shape is not None:
        full = full.reshape((values.shape[0],) + tuple(grid_shape))
    return full


def prepare_field(field, latitudes, lat_weights='scos', validation_frac=0.1):
    """Weight, flatten, drop missing features and split into training / validation rows:
    the steps between reading the anomalies and calling ``fit_*_model`` in the drivers.
    Returns ``(training, validation, missing_mask, weights)``."""
    weights = latitude_weights(latitudes, lat_weights)
    valid, missing = drop_missing_features(weight_and_flatten(field, weights))
    n_training = int(np.ceil((1 - validation_frac) * valid.shape[0]))
    return valid[:n_training], valid[n_training:], missing, weights
