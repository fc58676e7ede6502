"""Input checks shared by the estimators (reference ``validation_utils.py``)."""

import numpy as np


def check_unit_axis_sums(a, whom, axis=0):
    """Raise ``ValueError`` unless every sum along ``axis`` is close to one."""
    if not np.isclose(a.sum(axis=axis), 1).all():
        raise ValueError(
            'Array with incorrect axis sums passed to %s. '
            'Expected sums along axis %d to be 1.' % (whom, axis))


def check_array_shape(a, shape, whom):
    """Raise ``ValueError`` unless ``a.shape == shape``."""
    if a.shape != shape:
        raise ValueError(
            'Array with wrong shape passed to %s. '
            'Expected %s, but got %s' % (whom, shape, a.shape))


def check_stochastic_matrix(a, shape, whom, axis=0):
    """Shape check followed by the unit-sum check."""
    check_array_shape(a, shape, whom)
    check_unit_axis_sums(a, whom, axis=axis)
