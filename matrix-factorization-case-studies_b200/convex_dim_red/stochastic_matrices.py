"""Random stochastic matrices drawn on the host.

Mirrors ``stochastic_matrices.py:15-39`` of the reference.  These stay on the
host on purpose: the reference's drivers share one ``RandomState`` across
restarts (``bin/run_hadisst_aa.py:154-166``), so the MT19937 draw order is part
of the drop-in contract and device-side RNG would break seed parity.
"""

import numpy as np
from sklearn.utils import check_random_state


def _normalised_uniform(shape, random_state, axis):
    draws = check_random_state(random_state).uniform(size=shape)
    if axis not in (0, 1, -1, -2):
        raise ValueError('axis %d is out of bounds for array of dimension %d'
                         % (axis, draws.ndim))
    totals = np.sum(draws, axis=axis)
    if axis in (0, -2):
        return draws / totals[np.newaxis, :]
    return draws / totals[:, np.newaxis]


def left_stochastic_matrix(shape, random_state=None):
    """Uniform random matrix scaled to unit column sums."""
    return _normalised_uniform(shape, random_state, 0)


def right_stochastic_matrix(shape, random_state=None):
    """Uniform random matrix scaled to unit row sums."""
    return _normalised_uniform(shape, random_state, 1)
