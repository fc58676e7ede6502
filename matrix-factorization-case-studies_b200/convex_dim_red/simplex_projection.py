"""Projections onto the probability simplex (reference ``simplex_projection.py``).

Same names and array semantics as the reference (NumPy in, new NumPy array out,
optional ``out=`` like the reference's gufuncs); the work is done by
``cdr_simplex_project_rows`` / ``_columns`` on the GPU.  ``*_device`` variants
take and return device tensors for callers that keep data resident.
"""

import numpy as np

from . import _backend as be


def _project(a, columns, out=None):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim != 2:
        raise ValueError('expected a 2-D array, got %d-D' % a.ndim)
    m, n = a.shape
    result = out if out is not None else np.empty_like(a)
    if m == 0 or n == 0:
        return result
    dev = be.to_device(a)
    fn = be.library().cdr_simplex_project_columns if columns else \
        be.library().cdr_simplex_project_rows
    be.check(fn(dev.data_ptr(), dev.data_ptr(), m, n, n, n, None, be.stream_ptr()),
             'cdr_simplex_project')
    result[...] = dev.cpu().numpy()
    return result


def simplex_project_rows(A, out=None):
    """Project rows of matrix onto standard simplex (simplex_projection.py:40-47)."""
    return _project(A, False, out)


def simplex_project_columns(A, out=None):
    """Project columns of matrix onto standard simplex (simplex_projection.py:30-37)."""
    return _project(A, True, out)


def simplex_project_vector(x):
    """Project vector onto standard simplex (simplex_projection.py:13-27)."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim != 1:
        raise ValueError('expected a 1-D array')
    if x.size == 0:
        return x.copy()
    return _project(x[np.newaxis, :], False)[0]


def simplex_project_rows_device(t, out=None):
    """In-place capable device version: ``t`` is a (m, n) CUDA fp64 tensor."""
    out = t if out is None else out
    be.check(be.library().cdr_simplex_project_rows(
        t.data_ptr(), out.data_ptr(), t.shape[0], t.shape[1], t.stride(0), out.stride(0),
        None, be.stream_ptr()), 'cdr_simplex_project_rows')
    return out
