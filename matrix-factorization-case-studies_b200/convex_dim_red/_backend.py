"""ctypes binding of ``libcdr_b200.so`` (the C ABI in ``include/cdr_b200.h``).

PyTorch is used for device memory, streams and (in ``_dist``) collectives only;
every numerical kernel on the hot path lives in the shared library.  There is
no CPU fallback: without the library or without a CUDA device the calls raise.
"""

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CDR_LIBRARY: another build of the same library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get('CDR_LIBRARY') or os.path.join(_HERE, 'libcdr_b200.so')

LD_ALIGN = 32
MAX_COMPONENTS = 64
MAX_MEMORY = 8
GRAM_BATCH = 4


class BackendError(RuntimeError):
    """Raised when the CUDA library is missing or a launch fails."""


class SpgParams(ctypes.Structure):
    """Mirror of ``cdr_spg_params``."""

    _fields_ = [('gamma', ctypes.c_double), ('sigma_one', ctypes.c_double),
                ('sigma_two', ctypes.c_double), ('lambda_min', ctypes.c_double),
                ('alpha0', ctypes.c_double), ('alpha_min', ctypes.c_double),
                ('alpha_max', ctypes.c_double), ('epsilon_one', ctypes.c_double),
                ('epsilon_two', ctypes.c_double), ('memory', ctypes.c_int),
                ('max_iterations', ctypes.c_int), ('max_feval', ctypes.c_int),
                ('use_infinity_norm', ctypes.c_int)]


class LoopState(ctypes.Structure):
    """Mirror of ``cdr_loop_state`` (device-resident control block)."""

    _fields_ = [('done', ctypes.c_int), ('error_stage', ctypes.c_int),
                ('n_iter', ctypes.c_int), ('converged', ctypes.c_int),
                ('max_iterations', ctypes.c_int), ('stopping_rule', ctypes.c_int),
                ('require_monotone', ctypes.c_int), ('spg_iter', ctypes.c_int),
                ('spg_feval', ctypes.c_int), ('spg_active', ctypes.c_int),
                ('spg_alpha_set', ctypes.c_int), ('spg_warnings', ctypes.c_int),
                ('tolerance', ctypes.c_double), ('trace_data', ctypes.c_double),
                ('cost', ctypes.c_double), ('old_cost', ctypes.c_double),
                ('f_old', ctypes.c_double), ('f_new', ctypes.c_double),
                ('lam', ctypes.c_double), ('alpha', ctypes.c_double),
                ('delta', ctypes.c_double), ('dd', ctypes.c_double),
                ('a0', ctypes.c_double), ('a1', ctypes.c_double),
                ('beta', ctypes.c_double), ('res2', ctypes.c_double),
                ('resinf', ctypes.c_double), ('penalty', ctypes.c_double),
                ('f_mem', ctypes.c_double * MAX_MEMORY), ('tickets', ctypes.c_uint * 4)]


class SmallGramDesc(ctypes.Structure):
    """Mirror of ``cdr_small_gram_desc``."""

    _fields_ = [('A', ctypes.c_void_p), ('B', ctypes.c_void_p), ('out', ctypes.c_void_p),
                ('sAi', ctypes.c_long), ('sAn', ctypes.c_long),
                ('sBj', ctypes.c_long), ('sBn', ctypes.c_long),
                ('ka', ctypes.c_int), ('kb', ctypes.c_int), ('n', ctypes.c_int),
                ('mode', ctypes.c_int), ('scale', ctypes.c_double)]


class AaBuffers(ctypes.Structure):
    """Mirror of ``cdr_aa_buffers``."""

    _fields_ = [('C', ctypes.c_void_p), ('G', ctypes.c_void_p), ('D', ctypes.c_void_p),
                ('CK', ctypes.c_void_p), ('DK', ctypes.c_void_p), ('KZt', ctypes.c_void_p),
                ('alpha', ctypes.c_void_p), ('ZtZ', ctypes.c_void_p),
                ('CKCt', ctypes.c_void_p), ('CKZ', ctypes.c_void_p),
                ('G01', ctypes.c_void_p), ('G11', ctypes.c_void_p),
                ('row_scratch', ctypes.c_void_p), ('state', ctypes.c_void_p),
                ('cost_deltas', ctypes.c_void_p),
                ('k', ctypes.c_int), ('T', ctypes.c_int), ('ldt', ctypes.c_long),
                ('grad_scale', ctypes.c_double), ('cost_scale', ctypes.c_double)]


class GpnhProblem(ctypes.Structure):
    """Mirror of ``cdr_gpnh_problem``."""

    _fields_ = [('X', ctypes.c_void_p), ('ldx', ctypes.c_long),
                ('T', ctypes.c_int), ('d', ctypes.c_int), ('k', ctypes.c_int),
                ('T_total', ctypes.c_int), ('lambda_W', ctypes.c_double),
                ('Z', ctypes.c_void_p), ('WT', ctypes.c_void_p), ('XWt', ctypes.c_void_p),
                ('ldt', ctypes.c_long), ('ZtZ', ctypes.c_void_p), ('XWtZ', ctypes.c_void_p),
                ('WtW', ctypes.c_void_p), ('REG', ctypes.c_void_p), ('P', ctypes.c_void_p),
                ('state', ctypes.c_void_p), ('cost_deltas', ctypes.c_void_p),
                ('weights_params', SpgParams), ('workspace', ctypes.c_void_p),
                ('workspace_bytes', ctypes.c_size_t),
                ('peers', ctypes.c_void_p), ('T_min', ctypes.c_int)]


class KmeansState(ctypes.Structure):
    """Mirror of ``cdr_kmeans_state``."""

    _fields_ = [('done', ctypes.c_int), ('n_iter', ctypes.c_int), ('max_iter', ctypes.c_int),
                ('strict', ctypes.c_int), ('needs_relocation', ctypes.c_int),
                ('changed', ctypes.c_int), ('ticket', ctypes.c_uint), ('reserved_', ctypes.c_int),
                ('tol_abs', ctypes.c_double), ('shift_total', ctypes.c_double)]


class KmeansProblem(ctypes.Structure):
    """Mirror of ``cdr_kmeans_problem``."""

    _fields_ = [('X', ctypes.c_void_p), ('ldx', ctypes.c_long),
                ('T', ctypes.c_int), ('d', ctypes.c_int), ('k', ctypes.c_int),
                ('centres', ctypes.c_void_p), ('labels', ctypes.c_void_p),
                ('onehot', ctypes.c_void_p), ('ldt', ctypes.c_long), ('sums', ctypes.c_void_p),
                ('cnorm', ctypes.c_void_p), ('shift', ctypes.c_void_p), ('counts', ctypes.c_void_p),
                ('state', ctypes.c_void_p), ('workspace', ctypes.c_void_p),
                ('workspace_bytes', ctypes.c_size_t)]


class AaProblem(ctypes.Structure):
    """Mirror of ``cdr_aa_problem``."""

    _fields_ = [('X', ctypes.c_void_p), ('ldx', ctypes.c_long),
                ('T', ctypes.c_int), ('d', ctypes.c_int), ('buf', AaBuffers),
                ('Z', ctypes.c_void_p), ('tmp_kd', ctypes.c_void_p),
                ('dictionary_params', SpgParams), ('weights_params', SpgParams),
                ('workspace', ctypes.c_void_p), ('workspace_bytes', ctypes.c_size_t),
                ('row0', ctypes.c_int), ('T_min', ctypes.c_int), ('peers', ctypes.c_void_p)]


MAX_PEERS = 8
PEER_HEADER_BYTES = 1 << 20


class PeerGroupStruct(ctypes.Structure):
    """Mirror of ``cdr_peer_group``."""

    _fields_ = [('world', ctypes.c_int), ('rank', ctypes.c_int),
                ('region', ctypes.c_void_p * MAX_PEERS), ('region_bytes', ctypes.c_size_t),
                ('inbox_offset', ctypes.c_size_t), ('inbox_slot_bytes', ctypes.c_size_t)]


# name -> (restype, argtypes); used both to bind and by the CPU-side symbol test
_vp, _i, _l, _d, _sz = (ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_double,
                        ctypes.c_size_t)
SIGNATURES = {
    'cdr_version': (ctypes.c_char_p, []),
    'cdr_device_check': (_i, []),
    'cdr_launch_count': (ctypes.c_ulonglong, []),
    'cdr_debug_stream_plan': (_i, [_i, _i, _i, _i, ctypes.POINTER(ctypes.c_int)]),
    'cdr_debug_dmma_probe': (_i, [_vp, _i, _i, _vp]),
    'cdr_debug_timing_report': (_i, [_i]),
    'cdr_simplex_project_rows': (_i, [_vp, _vp, _i, _i, _l, _l, _vp, _vp]),
    'cdr_simplex_project_columns': (_i, [_vp, _vp, _i, _i, _l, _l, _vp, _vp]),
    'cdr_quad_simplex_spg_batched': (_i, [_vp, _vp, _vp, _l, _l, _vp, _i, _i,
                                          ctypes.POINTER(SpgParams), _vp, _vp, _vp, _vp]),
    'cdr_reduce_samples_workspace_bytes': (_sz, [_i, _i, _i]),
    'cdr_reduce_samples': (_i, [_vp, _l, _l, _vp, _l, _i, _i, _i, _vp, _vp, _l, _vp, _sz,
                                _vp, _vp]),
    'cdr_reduce_features_workspace_bytes': (_sz, [_i, _i, _i]),
    'cdr_reduce_features': (_i, [_vp, _l, _vp, _l, _i, _i, _i, _vp, _l, _vp, _sz, _vp, _vp]),
    'cdr_gram_workspace_bytes': (_sz, [_i, _i]),
    'cdr_gram': (_i, [_vp, _l, _i, _i, _vp, _l, _vp, _sz, _vp]),
    'cdr_syrk_workspace_bytes': (_sz, [_i, _i]),
    'cdr_syrk': (_i, [_vp, _l, _i, _i, _vp, _l, _i, _i, _vp, _sz, _vp]),
    'cdr_frobenius_workspace_bytes': (_sz, []),
    'cdr_frobenius_sq': (_i, [_vp, _l, _i, _i, _vp, _vp, _sz, _vp]),
    'cdr_sum_vector': (_i, [_vp, _i, _vp, _vp]),
    'cdr_small_gram_workspace_bytes': (_sz, []),
    'cdr_small_gram': (_i, [ctypes.POINTER(SmallGramDesc), _i, _vp, _sz, _vp, _vp]),
    'cdr_sym_pinv_workspace_bytes': (_sz, [_i]),
    'cdr_gpnh_solve_matrix': (_i, [_vp, _i, _i, _i, _d, _vp, _vp, _sz, _vp, _vp]),
    'cdr_sym_pinv': (_i, [_vp, _i, _vp, _vp, _vp]),
    'cdr_residual_sq': (_i, [_vp, _l, _i, _i, _vp, _i, _vp, _l, _vp, _vp, _vp]),
    'cdr_aa_gradient': (_i, [ctypes.POINTER(AaBuffers), _vp]),
    'cdr_aa_dictionary_cost': (_i, [ctypes.POINTER(AaBuffers), _d, _vp, _vp]),
    'cdr_loop_begin': (_i, [_vp, _vp]),
    'cdr_aa_spg_begin': (_i, [ctypes.POINTER(AaBuffers), ctypes.POINTER(SpgParams), _vp]),
    'cdr_aa_spg_direction': (_i, [ctypes.POINTER(AaBuffers), ctypes.POINTER(SpgParams), _vp]),
    'cdr_aa_spg_linesearch': (_i, [ctypes.POINTER(AaBuffers), ctypes.POINTER(SpgParams), _vp]),
    'cdr_aa_spg_update': (_i, [ctypes.POINTER(AaBuffers), ctypes.POINTER(SpgParams), _i, _vp]),
    'cdr_aa_scale_factors_step': (_i, [ctypes.POINTER(AaBuffers), ctypes.POINTER(SpgParams), _d, _vp]),
    'cdr_aa_cost_check': (_i, [ctypes.POINTER(AaBuffers), _i, _i, _vp]),
    'cdr_gpnh_cost_check': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _d, _i, _i, _vp]),
    'cdr_gpnh_workspace_bytes': (_sz, [_i, _i, _i]),
    'cdr_gpnh_prepare_enqueue': (_i, [ctypes.POINTER(GpnhProblem), _vp]),
    'cdr_gpnh_iterate_enqueue': (_i, [ctypes.POINTER(GpnhProblem), _vp]),
    'cdr_gpnh_fused_applicable': (_i, [_i, _i, _i]),
    'cdr_aa_workspace_bytes': (_sz, [_i, _i, _i]),
    'cdr_aa_prepare_enqueue': (_i, [ctypes.POINTER(AaProblem), _vp]),
    'cdr_aa_iterate_enqueue': (_i, [ctypes.POINTER(AaProblem), _vp]),
    'cdr_aa_fused_applicable': (_i, [_i, _i, _i, _i]),
    'cdr_dissimilarity_from_gram': (_i, [_vp, _l, _i, _vp, _l, _vp]),
    'cdr_furthest_sum_workspace_bytes': (_sz, [_i, _i, _i]),
    'cdr_furthest_sum': (_i, [_vp, _l, _i, _i, _i, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    'cdr_kmeans_labels': (_i, [_vp, _l, _vp, _i, _i, _vp, _vp, _l, _vp, _vp, _vp]),
    'cdr_kmeans_sqdist': (_i, [_vp, _l, _i, _i, _vp, _l, _vp, _vp, _vp]),
    'cdr_kmeans_update': (_i, [_vp, _l, _vp, _vp, _l, _i, _i, _vp, _vp]),
    'cdr_kmeans_workspace_bytes': (_sz, [_i, _i, _i]),
    'cdr_kmeans_fused_applicable': (_i, [_i, _i, _i]),
    'cdr_kmeans_prepare_enqueue': (_i, [ctypes.POINTER(KmeansProblem), _vp]),
    'cdr_kmeans_iterate_enqueue': (_i, [ctypes.POINTER(KmeansProblem), _vp]),
    'cdr_row_sqnorms': (_i, [_vp, _l, _i, _i, _vp, _vp]),
    'cdr_column_moments': (_i, [_vp, _l, _i, _i, _vp, _vp, _vp]),
    'cdr_center_columns': (_i, [_vp, _l, _i, _i, _vp, _d, _vp]),
    'cdr_peer_region_alloc': (_i, [_sz, ctypes.POINTER(ctypes.c_void_p)]),
    'cdr_peer_region_free': (_i, [_vp]),
    'cdr_peer_export': (_i, [_vp, ctypes.c_char_p]),
    'cdr_peer_import': (_i, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
    'cdr_peer_release': (_i, [_vp]),
    'cdr_peer_error': (_i, [ctypes.POINTER(PeerGroupStruct), ctypes.POINTER(ctypes.c_int), _vp]),
    'cdr_peer_allreduce': (_i, [ctypes.POINTER(PeerGroupStruct), _sz, _sz, _vp, _vp]),
    'cdr_peer_allgather_columns': (_i, [ctypes.POINTER(PeerGroupStruct), _vp, _l, _sz, _l, _i,
                                        _i, _i, _i, _vp, _vp]),
    'cdr_reduce_samples_allreduce': (_i, [ctypes.POINTER(PeerGroupStruct), _vp, _l, _l, _vp, _l,
                                          _i, _i, _i, _i, _vp, _sz, _l, _vp, _vp]),
}

_LIB = None


def library():
    """Load the shared library (once).  Fails loudly when it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise BackendError(
                'libcdr_b200.so not found at %s: build it with '
                '`python -c "import __graft_entry__ as g; g.build()"` '
                '(there is no CPU fallback)' % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)           # AttributeError if the symbol is missing
            fn.restype = restype
            fn.argtypes = argtypes
        _LIB = lib
    return _LIB


_ERRORS = {-1: 'invalid argument', -2: 'unsupported configuration', -3: 'workspace too small',
           -4: 'shape not covered by the fused kernel'}
ERR_NOT_APPLICABLE = -4


def check(rc, what):
    if rc != 0:
        msg = _ERRORS.get(rc, 'CUDA error %d' % rc)
        raise BackendError('%s failed: %s' % (what, msg))


def torch_mod():
    import torch
    return torch


def require_cuda():
    torch = torch_mod()
    if not torch.cuda.is_available():
        raise BackendError('convex_dim_red (B200 build) needs a CUDA device; '
                           'there is no CPU fallback')
    return torch


def stream_ptr():
    return torch_mod().cuda.current_stream().cuda_stream


def ptr(t):
    """Device address of a tensor (ints pass through, None stays NULL)."""
    if t is None or isinstance(t, int):
        return t
    return t.data_ptr()


def round_up(n, m=LD_ALIGN):
    return (n + m - 1) // m * m


def zeros(*shape, dtype=None):
    torch = require_cuda()
    return torch.zeros(*shape, dtype=dtype or torch.float64, device='cuda')


_RESIDENT = {}
_RESIDENT_DERIVED = {}      # (key of a resident array, name) -> device tensor derived from it


def _cache_key(a):
    return (a.__array_interface__['data'][0], a.shape, a.strides, a.dtype.str)


class DeviceCache:
    """Keeps padded device copies of the given host arrays for the duration of a ``with``
    block; ``to_device_padded`` returns the cached copy for the same memory.  The caller
    promises not to modify the arrays (or the device copies) inside the block."""

    def __init__(self, arrays):
        self.arrays = [a for a in arrays if isinstance(a, np.ndarray) and a.ndim == 2]
        self.keys = []

    def __enter__(self):
        for a in self.arrays:
            if a.dtype == np.float64 and a.flags.c_contiguous:
                key = _cache_key(a)
                if key not in _RESIDENT:
                    _RESIDENT[key] = _upload_padded(a)
                    self.keys.append(key)
        return self

    def __exit__(self, *exc):
        for key in self.keys:
            _RESIDENT.pop(key, None)
            for dkey in [k for k in _RESIDENT_DERIVED if k[0] == key]:
                _RESIDENT_DERIVED.pop(dkey, None)


def resident_derived(a, name, builder):
    """``builder()`` for a host array, memoised while the array is resident (e.g. its Gram
    matrix, shared by the FurthestSum initialisations of all restarts)."""
    if isinstance(a, np.ndarray) and _RESIDENT:
        key = _cache_key(a)
        if key in _RESIDENT:
            dkey = (key, name)
            if dkey not in _RESIDENT_DERIVED:
                _RESIDENT_DERIVED[dkey] = builder()
            return _RESIDENT_DERIVED[dkey]
    return builder()


def _upload_padded(a):
    torch = require_cuda()
    rows, cols = a.shape
    buf = torch.zeros((rows, round_up(cols)), dtype=torch.float64, device='cuda')
    buf[:, :cols].copy_(torch.from_numpy(a))
    return buf


def to_device_padded(a, dtype=np.float64):
    """Copy a 2-D host array into a zero-padded (rows, round_up(cols)) device buffer."""
    a = np.ascontiguousarray(a, dtype=dtype)
    if _RESIDENT:
        hit = _RESIDENT.get(_cache_key(a))
        if hit is not None:
            return hit
    return _upload_padded(a)


def to_device(a, dtype=np.float64):
    torch = require_cuda()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).cuda()


def to_host(t, rows=None, cols=None):
    """Copy (a leading block of) a device matrix back as a C-contiguous array."""
    if rows is not None:
        t = t[:rows, :cols]
    return np.ascontiguousarray(t.cpu().numpy())


def make_spg_params(kwargs=None, **defaults):
    """Solver options with the reference's defaults (spg.py:46-51, :287-291)."""
    vals = dict(gamma=1e-4, memory=1, sigma_one=0.1, sigma_two=0.9, lambda_min=1e-10,
                alpha0=-1.0, alpha_min=1e-5, alpha_max=1e3, epsilon_one=1e-10,
                epsilon_two=1e-6, max_iterations=1000, max_feval=2000,
                use_infinity_norm=1)
    vals.update(defaults)
    if kwargs:
        for key, value in kwargs.items():
            if key in vals:
                vals[key] = value
    if vals['alpha0'] is None:
        vals['alpha0'] = -1.0
    vals['use_infinity_norm'] = int(bool(vals['use_infinity_norm']))
    p = SpgParams()
    for key, value in vals.items():
        setattr(p, key, value)
    if not 1 <= p.memory <= MAX_MEMORY:
        raise ValueError('memory must be in [1, %d] in the B200 build; got %r'
                         % (MAX_MEMORY, vals['memory']))
    return p


class Workspace:
    """Device scratch shared by the streaming passes and the small products."""

    def __init__(self, T, d, k):
        lib = library()
        torch = require_cuda()
        self.nbytes_samples = lib.cdr_reduce_samples_workspace_bytes(T, d, k)
        self.nbytes_features = lib.cdr_reduce_features_workspace_bytes(T, d, k)
        self.nbytes_gram = lib.cdr_small_gram_workspace_bytes()
        n = max(self.nbytes_samples, self.nbytes_features, 8)
        # scratch is fully written before it is read: no need to clear it
        self.stream_ws = torch.empty(n // 8 + 1, dtype=torch.float64, device='cuda')
        self.gram_ws = torch.empty(self.nbytes_gram // 8 + 1, dtype=torch.float64, device='cuda')

    def ensure(self, T, d, k):
        lib = library()
        torch = torch_mod()
        n = max(lib.cdr_reduce_samples_workspace_bytes(T, d, k),
                lib.cdr_reduce_features_workspace_bytes(T, d, k), 8)
        if n > self.stream_ws.numel() * 8:
            self.stream_ws = torch.zeros(n // 8 + 1, dtype=torch.float64, device='cuda')


# ---------------------------------------------------------------------------
# thin wrappers (device tensors in, device tensors out, asynchronous)
# ---------------------------------------------------------------------------

def reduce_samples(L, sLi, sLt, X, T, d, k, out, ws, E=None, flags=None):
    """out (k x ld) = E (L X); see cdr_reduce_samples."""
    check(library().cdr_reduce_samples(
        ptr(L), sLi, sLt, ptr(X), X.stride(0), T, d, k, ptr(E), ptr(out), out.stride(0),
        ptr(ws.stream_ws), ws.stream_ws.numel() * 8, ptr(flags), stream_ptr()),
        'cdr_reduce_samples')


def reduce_features(M, X, T, d, k, out, ws, flags=None):
    """out (k x T) = M X'; see cdr_reduce_features."""
    check(library().cdr_reduce_features(
        ptr(M), M.stride(0), ptr(X), X.stride(0), T, d, k, ptr(out), out.stride(0),
        ptr(ws.stream_ws), ws.stream_ws.numel() * 8, ptr(flags), stream_ptr()),
        'cdr_reduce_features')


def small_gram(descs, ws, flags=None):
    """descs: list of (A, sAi, sAn, ka, B, sBj, sBn, kb, n, out, scale, mode)."""
    arr = (SmallGramDesc * len(descs))()
    for i, (A, sAi, sAn, ka, B, sBj, sBn, kb, n, out, scale, mode) in enumerate(descs):
        arr[i] = SmallGramDesc(ptr(A), ptr(B), ptr(out), sAi, sAn, sBj, sBn, ka, kb, n, mode,
                               scale)
    check(library().cdr_small_gram(arr, len(descs), ptr(ws.gram_ws), ws.gram_ws.numel() * 8,
                                   ptr(flags), stream_ptr()), 'cdr_small_gram')


def quad_simplex_spg_batched(A, alpha, B, sb_t, sb_c, Z, T, k, params, n_iter=None,
                             n_feval=None, flags=None):
    check(library().cdr_quad_simplex_spg_batched(
        ptr(A), ptr(alpha), ptr(B), sb_t, sb_c, ptr(Z), T, k, ctypes.byref(params),
        ptr(n_iter), ptr(n_feval), ptr(flags), stream_ptr()), 'cdr_quad_simplex_spg_batched')


def frobenius_sq(X, T, d):
    torch = require_cuda()
    out = torch.zeros(1, dtype=torch.float64, device='cuda')
    ws = torch.zeros(library().cdr_frobenius_workspace_bytes() // 8 + 1, dtype=torch.float64,
                     device='cuda')
    check(library().cdr_frobenius_sq(ptr(X), X.stride(0), T, d, ptr(out), ptr(ws),
                                     ws.numel() * 8, stream_ptr()), 'cdr_frobenius_sq')
    return out


def gram(X, T, d, comm=None):
    """K = X X' as a padded (T, round_up(T)) device matrix (SYRK: upper-triangle tiles,
    mirrored; csrc/syrk.cu).

    With a process group every rank holds the same X (all T rows): the upper-triangle tiles
    are dealt round-robin to the ranks and summed into place by one all-reduce (zeros
    elsewhere, so the sum is exact and K is bit-identical to the single-GPU result).
    """
    torch = require_cuda()
    lib = library()
    K = torch.zeros((T, round_up(T)), dtype=torch.float64, device='cuda')
    nbytes = lib.cdr_syrk_workspace_bytes(T, d)
    ws = torch.empty(nbytes // 8 + 1, dtype=torch.float64, device='cuda')
    sharded = comm is not None and comm.enabled
    index, count = (comm.rank, comm.world) if sharded else (0, 1)
    check(lib.cdr_syrk(ptr(X), X.stride(0), T, d, ptr(K), K.stride(0), index, count, ptr(ws),
                       ws.numel() * 8, stream_ptr()), 'cdr_syrk')
    if sharded:
        comm.allreduce_sum(K)
    return K


def gram_slabs(X, T, d):
    """K = X X' by repeated 64-row feature reductions (cdr_gram; the round-1 path, kept as a
    cross-check of the SYRK kernel)."""
    torch = require_cuda()
    lib = library()
    K = torch.zeros((T, round_up(T)), dtype=torch.float64, device='cuda')
    nbytes = lib.cdr_gram_workspace_bytes(T, d)
    ws = torch.zeros(nbytes // 8 + 1, dtype=torch.float64, device='cuda')
    check(lib.cdr_gram(ptr(X), X.stride(0), T, d, ptr(K), K.stride(0), ptr(ws),
                       ws.numel() * 8, stream_ptr()), 'cdr_gram')
    return K


_DMMA_PEAK = [None]


def dmma_peak_tflops():
    """Measured fp64 tensor-pipe peak of this GPU (DMMA.8x8x4 on registers, best of 5 launches,
    CUDA events): the roofline denominator of the tensor-bound shapes."""
    if _DMMA_PEAK[0] is None:
        torch = require_cuda()
        lib = library()
        blocks, iters = 148 * 4, 20000
        out = torch.empty(blocks * 256, dtype=torch.float64, device='cuda')
        best = None
        for _ in range(6):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            check(lib.cdr_debug_dmma_probe(ptr(out), blocks, iters, stream_ptr()),
                  'cdr_debug_dmma_probe')
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None or ms < best else best
        flops = 2.0 * 256 * 8 * iters * 8 * blocks
        _DMMA_PEAK[0] = flops / (best * 1e-3) / 1e12
    return _DMMA_PEAK[0]


_TRACE_T0 = [None]


def trace(msg):
    """``CDR_TRACE_TIMING=1``: synchronise and print the wall time since the last call."""
    if os.environ.get('CDR_TRACE_TIMING', '0') != '1':
        return
    import sys
    import time
    torch_mod().cuda.synchronize()
    now = time.perf_counter()
    if _TRACE_T0[0] is not None:
        sys.stderr.write('[cdr trace] %-32s %8.2f ms\n' % (msg, (now - _TRACE_T0[0]) * 1e3))
    _TRACE_T0[0] = now


def graphs_disabled():
    """``CDR_NO_CUDA_GRAPH=1`` launches every iteration eagerly (debugging aid)."""
    return os.environ.get('CDR_NO_CUDA_GRAPH', '0') == '1'


def graph_after(c_loop):
    """Outer iterations launched eagerly before the iteration is captured into a CUDA graph.

    Capturing and instantiating costs 2.5-3 ms.  An engine on the whole-iteration C entry
    points launches 3-9 kernels of ~0.1 ms each per outer iteration, so the host keeps far
    ahead of the device without a graph and a short fit (a 20-iteration call is ~9 ms of device
    time) should not pay for one; in a long fit the replayed graph's smaller launch gaps win
    the cost back after ~150 iterations.  The general kernel sequences (12-25 short launches
    per iteration) are captured right after the first iteration.  ``CDR_GRAPH_AFTER=n``
    overrides."""
    env = os.environ.get('CDR_GRAPH_AFTER')
    if env is not None:
        return max(1, int(env))
    return 32 if c_loop else 1


def graph_collectives():
    """Whether NCCL collectives are captured into the iteration graph (``CDR_GRAPH_NCCL``,
    default on; set to 0 to launch sharded iterations eagerly)."""
    return os.environ.get('CDR_GRAPH_NCCL', '1') == '1'


def capture_graph(fn):
    """Capture the kernels ``fn`` enqueues on the current stream into a CUDA graph.

    Uses ``capture_begin`` / ``capture_end`` on a side stream directly: the
    ``torch.cuda.graph`` context manager also runs ``gc.collect()`` and
    ``torch.cuda.empty_cache()``, which costs tens of milliseconds (cudaFree of every cached
    block) inside what is otherwise a 2 ms operation.
    """
    torch = torch_mod()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        graph.capture_begin()
        try:
            fn()
        finally:
            graph.capture_end()
    torch.cuda.current_stream().wait_stream(side)
    return graph


class DeviceState:
    """Host handle of a device-resident ``cdr_loop_state``."""

    def __init__(self, tolerance, max_iterations, stopping_criterion, require_monotone):
        torch = require_cuda()
        if stopping_criterion not in ('abs_delta_f', 'rel_delta_f'):
            raise ValueError("unsupported stopping criterion '%s'" % stopping_criterion)
        init = LoopState()
        init.tolerance = float(tolerance)
        init.max_iterations = int(max_iterations)
        init.stopping_rule = 0 if stopping_criterion == 'abs_delta_f' else 1
        init.require_monotone = int(bool(require_monotone))
        raw = np.frombuffer(bytes(init), dtype=np.uint8).copy()
        self.buf = torch.from_numpy(raw).cuda()
        self.max_iterations = int(max_iterations)
        self.cost_deltas = torch.zeros(max(1, int(max_iterations)), dtype=torch.float64,
                                       device='cuda')

    @property
    def ptr(self):
        return self.buf.data_ptr()

    def read(self):
        """Synchronising read of the control block."""
        raw = self.buf.cpu().numpy().tobytes()
        return LoopState.from_buffer_copy(raw)

    def write_field(self, name, value):
        st = self.read()
        setattr(st, name, value)
        raw = np.frombuffer(bytes(st), dtype=np.uint8).copy()
        self.buf.copy_(torch_mod().from_numpy(raw))
