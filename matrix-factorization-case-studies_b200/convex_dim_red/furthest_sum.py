"""FurthestSum initialisation (reference ``furthest_sum.py``)."""

import numpy as np

from . import _backend as be


def _validate(n_samples, n_components, start_index, exclude):
    n_excluded = len(exclude)
    if start_index >= n_samples:
        raise ValueError('Start index %r is out of bounds (n_samples = %d)' %
                         (start_index, n_samples))
    for index in exclude:
        if index == start_index:
            raise ValueError('Start index %r is excluded' % start_index)
    if n_excluded < n_samples and n_components > n_samples - n_excluded:
        raise ValueError(
            'Too few point available to select requested number of components '
            '(n_components=%d, n_samples=%d, n_excluded=%d)' %
            (n_components, n_samples, n_excluded))


def furthest_sum_device(D, n_samples, n_components, start_index, exclude=None, extra_steps=1):
    """FurthestSum on a device-resident (n_samples, ld) dissimilarity matrix.

    Returns the selected indices as an int64 NumPy array (furthest_sum.py:23-127).
    """
    torch = be.require_cuda()
    exclude = [] if exclude is None else [int(i) for i in exclude]
    _validate(n_samples, n_components, start_index, exclude)
    lib = be.library()
    extra_steps = max(int(extra_steps), 0)
    nbytes = lib.cdr_furthest_sum_workspace_bytes(n_samples, n_components, extra_steps)
    ws = torch.zeros(nbytes // 8 + 1, dtype=torch.float64, device='cuda')
    sel = torch.zeros(n_components, dtype=torch.int64, device='cuda')
    excl = torch.tensor(exclude, dtype=torch.int64, device='cuda') if exclude else None
    be.check(lib.cdr_furthest_sum(
        D.data_ptr(), D.stride(0), n_samples, n_components, int(start_index), be.ptr(excl),
        len(exclude), extra_steps, sel.data_ptr(), ws.data_ptr(), ws.numel() * 8,
        be.stream_ptr()), 'cdr_furthest_sum')
    return sel.cpu().numpy()


def furthest_sum(dissimilarity_matrix, n_components, start_index,
                 exclude=None, extra_steps=1):
    """Run FurthestSum initialization on given dissimilarity matrix.

    Same signature, validation errors and int64 result as the reference
    (furthest_sum.py:130-170).
    """
    dissimilarity_matrix = np.asarray(dissimilarity_matrix)
    if dissimilarity_matrix.shape[0] != dissimilarity_matrix.shape[1]:
        raise ValueError(
            'Dissimilarity matrix must be square, but got shape %r' %
            list(dissimilarity_matrix.shape))
    if n_components == 0:
        return []
    n_samples = dissimilarity_matrix.shape[0]
    exclude = [] if exclude is None else list(exclude)
    _validate(n_samples, n_components, start_index, exclude)
    D = be.to_device(dissimilarity_matrix)
    return furthest_sum_device(D, n_samples, n_components, start_index, exclude, extra_steps)


def dissimilarity_from_gram_device(K, n_samples):
    """sqrt(K_jj - 2 K_ij + K_ii) on the device (archetypal_analysis.py:96-100)."""
    torch = be.require_cuda()
    D = torch.zeros((n_samples, be.round_up(n_samples)), dtype=torch.float64, device='cuda')
    be.check(be.library().cdr_dissimilarity_from_gram(
        K.data_ptr(), K.stride(0), n_samples, D.data_ptr(), D.stride(0), be.stream_ptr()),
        'cdr_dissimilarity_from_gram')
    return D
