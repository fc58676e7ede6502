"""Argument checks used by the estimators.

API-compatible with the reference's ``validation_utils.py`` (same function names, same
``ValueError`` texts, which callers and tests may match on); implemented on top of one
private helper.
"""

import numpy as np

_SHAPE_TEXT = 'Array with wrong shape passed to {whom}. Expected {want}, but got {got}'
_SUMS_TEXT = ('Array with incorrect axis sums passed to {whom}. '
              'Expected sums along axis {axis:d} to be 1.')


def _fail(template, **fields):
    raise ValueError(template.format(**fields))


def check_array_shape(a, shape, whom):
    """``a.shape`` must equal ``shape``."""
    got = a.shape
    if got != shape:
        _fail(_SHAPE_TEXT, whom=whom, want=shape, got=got)


def check_unit_axis_sums(a, whom, axis=0):
    """Every sum of ``a`` along ``axis`` must be one (``numpy.isclose`` tolerances)."""
    totals = np.add.reduce(a, axis=axis)
    if not bool(np.isclose(totals, 1).all()):
        _fail(_SUMS_TEXT, whom=whom, axis=axis)


def check_stochastic_matrix(a, shape, whom, axis=0):
    """Right shape first, then unit sums along ``axis``."""
    check_array_shape(a, shape, whom)
    check_unit_axis_sums(a, whom, axis=axis)
