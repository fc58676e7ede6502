"""Archetypal analysis (reference ``archetypal_analysis.py``).

Same public surface as the reference: ``ArchetypalAnalysis``, ``KernelAA`` and
the private helpers its tests import (``_iterate_kernel_aa``, ``_kernel_aa_cost``,
``_update_kernel_aa_dictionary``, ``_update_kernel_aa_weights``, ...), NumPy in
and out.  The alternating loop runs on the GPU (``_AaEngine``): the dictionary
SPG, its line search, the per-sample QPs and the convergence tests are device
kernels; one outer iteration is replayed from a CUDA graph when the inner
iteration count is static.
"""

import ctypes
import numbers
import time
import warnings

import numpy as np
from sklearn.utils import check_array, check_random_state

from . import _backend as be
from ._dist import Comm, shard_bounds, shard_sizes
from .furthest_sum import dissimilarity_from_gram_device, furthest_sum_device
from .spg import spg
from .stochastic_matrices import right_stochastic_matrix
from .validation_utils import check_array_shape, check_stochastic_matrix

INTEGER_TYPES = (numbers.Integral, np.integer)

INITIALIZATION_METHODS = (None, 'random', 'furthest_sum',)

_STAGES = {1: 'scale factors', 2: 'dictionary', 3: 'weights'}

# inner SPG iterations up to which the whole outer iteration is captured in one graph
_MAX_UNROLLED_SPG = 8
# csrc/aa_steps.cu kRowMaxT: a dictionary row (length T) is staged in shared memory
MAX_SAMPLES = 26000


def _check_init_weights(weights, shape, whom):
    weights = check_array(weights)
    check_stochastic_matrix(weights, shape, whom, axis=1)


def _check_init_dictionary(dictionary, shape, whom):
    dictionary = check_array(dictionary)
    check_stochastic_matrix(dictionary, shape, whom, axis=1)


def _check_init_scale_factors(alpha, delta, shape, whom):
    check_array_shape(alpha, shape, whom)
    if np.any(np.logical_or(alpha < 1 - delta, alpha > 1 + delta)):
        raise ValueError('Initial scale factors infeasible in %s' % (whom))


def _dictionary_params(kwargs):
    """Options of the dictionary SPG with the defaults of spg() (spg.py:46-51)."""
    return be.make_spg_params(kwargs, max_iterations=10000, max_feval=1000000)


class _AaEngine:
    """Device-resident state of one AA fit.

    ``mode='kernel'``: ``data`` is the T x T kernel matrix (archetypal_analysis.py:399-531).
    ``mode='feature'``: ``data`` is the T x d data matrix (:534-670); products with
    K = X X' are two streaming passes over X and K is never formed.
    """

    def __init__(self, data, weights, dictionary, alpha, mode, delta=0, tolerance=1e-6,
                 max_iterations=1000, stopping_criterion='abs_delta_f',
                 require_monotonic_cost_decrease=True, weights_solver_kwargs=None,
                 dictionary_solver_kwargs=None, scale_factors_solver_kwargs=None,
                 update_weights=True, update_dictionary=True, update_scale_factors=True,
                 trace_data=None, data_device=None, comm=None, grad_scale=None):
        torch = be.require_cuda()
        self.mode = mode
        # sample-sharded fit: `data` / `weights` hold this rank's rows; the dictionary and
        # every k x T quantity are replicated (T = total number of samples)
        self.comm = comm if comm is not None else Comm(enabled=False)
        if self.comm.enabled and mode != 'feature':
            raise NotImplementedError('sample sharding is implemented for feature-space AA')
        self.Tl = weights.shape[0]
        self.k = weights.shape[1]
        self.T = np.asarray(dictionary).shape[1]
        self.sizes = shard_sizes(self.T, self.comm.world)
        self.lo = shard_bounds(self.T, self.comm.world, self.comm.rank)[0]
        if self.sizes[self.comm.rank] != self.Tl:
            raise ValueError('local weights have %d rows, expected %d of %d'
                             % (self.Tl, self.sizes[self.comm.rank], self.T))
        T, k = self.T, self.k
        if k > be.MAX_COMPONENTS:
            raise ValueError('n_components > %d is not supported by the B200 build'
                             % be.MAX_COMPONENTS)
        if T > MAX_SAMPLES:
            raise ValueError('archetypal analysis of more than %d samples (total, also when '
                             'sharded) is not supported by the B200 build: the dictionary row '
                             'kernels stage one k x T row in shared memory; got %d'
                             % (MAX_SAMPLES, T))
        self.d = data.shape[1]
        self.delta = delta
        self.update_weights = update_weights
        self.update_dictionary = update_dictionary
        self.update_scale_factors = update_scale_factors
        self.w_params = be.make_spg_params(weights_solver_kwargs)
        self.d_params = _dictionary_params(dictionary_solver_kwargs)
        self.s_params = _dictionary_params(dict(scale_factors_solver_kwargs or {}))
        self.lib = be.library()

        self.X = data_device if data_device is not None else be.to_device_padded(data)
        self.ldx = self.X.stride(0)
        self.ldt = be.round_up(T)
        ldt = self.ldt
        self.C = be.to_device_padded(dictionary)
        self.Z = be.to_device(weights)
        self.alpha = be.to_device(np.asarray(alpha, dtype=np.float64))
        self.G = be.zeros(k, ldt)
        self.D = be.zeros(k, ldt)
        # buffers that are summed / gathered over ranks live in the symmetric peer region when
        # the peer-memory collectives are on (CDR_PEER_COLLECTIVES=1); plain tensors otherwise
        self.peer = None
        if mode == 'feature':
            self.peer = self.comm.setup_peer(
                [(k, ldt), (k, ldt), (k, ldt), (k, k), (k, self.ldx)], (k, self.ldx))
        self.CK = self.comm.zeros(k, ldt)
        self.DK = self.comm.zeros(k, ldt)
        self.KZt = self.comm.zeros(k, ldt)
        self.ZtZ = self.comm.zeros(k, k)
        self.CKCt = be.zeros(k, k)
        self.CKZ = be.zeros(k, k)
        self.G01 = be.zeros(k, k)
        self.G11 = be.zeros(k, k)
        self.row_scratch = be.zeros(8 * k)
        if mode == 'feature':
            self.tmp_kd = self.comm.zeros(k, self.ldx)
        else:
            self.Zt = be.zeros(k, ldt)
        if self.comm.enabled:
            self.loc_kt = be.zeros(k, be.round_up(self.Tl))
            self.gather_scratch = be.zeros(self.comm.world + 1, k, max(self.sizes))
        self.ws = be.Workspace(max(T, self.Tl), self.d, k)
        self.state = be.DeviceState(tolerance, max_iterations, stopping_criterion,
                                    require_monotonic_cost_decrease)
        if trace_data is None:
            if mode == 'feature':
                # archetypal_analysis.py:552 forms X X' only for its trace
                tr = be.frobenius_sq(self.X, self.Tl, self.d)
                self.comm.allreduce_sum(tr)
                trace_data = float(tr.item())
            else:
                trace_data = float(np.asarray(data).trace())
        self.trace_data = float(trace_data)
        self.state.write_field('trace_data', self.trace_data)
        # archetypal_analysis.py:265,277 divide the dictionary cost by k; the gradient is
        # divided by T in feature space (:297) and by k in kernel space (:288)
        if grad_scale is None:
            grad_scale = 1.0 / T if mode == 'feature' else 1.0 / k
        self.buf = be.AaBuffers(
            self.C.data_ptr(), self.G.data_ptr(), self.D.data_ptr(), self.CK.data_ptr(),
            self.DK.data_ptr(), self.KZt.data_ptr(), self.alpha.data_ptr(),
            self.ZtZ.data_ptr(), self.CKCt.data_ptr(), self.CKZ.data_ptr(),
            self.G01.data_ptr(), self.G11.data_ptr(), self.row_scratch.data_ptr(),
            self.state.ptr, self.state.cost_deltas.data_ptr(), k, T, ldt, grad_scale, 1.0 / k)
        self._first_dictionary_update = True
        # feature-space fits with a short inner SPG loop and no scale-factor update run behind
        # the C entry points cdr_aa_prepare_enqueue / cdr_aa_iterate_enqueue (eight kernels per
        # outer iteration at streaming shapes with one inner iteration): always on a single
        # GPU, and for a sample-sharded fit when the peer-memory collectives are on and the
        # eight-kernel path covers the shape
        eligible = (mode == 'feature' and update_weights and update_dictionary and
                    not (update_scale_factors and delta != 0) and
                    1 <= self.d_params.max_iterations <= _MAX_UNROLLED_SPG)
        self.c_sharded = (eligible and self.comm.enabled and self.peer is not None and
                          bool(self.lib.cdr_aa_fused_applicable(
                              min(self.sizes), self.d, k, self.d_params.max_iterations)))
        self.c_loop = (eligible and not self.comm.enabled) or self.c_sharded
        if self.c_loop:
            nbytes = self.lib.cdr_aa_workspace_bytes(self.Tl, self.d, k)
            self.c_ws = torch.empty(nbytes // 8 + 1, dtype=torch.float64, device='cuda')
            self.problem = be.AaProblem(
                self.X.data_ptr(), self.ldx, self.Tl, self.d, self.buf, self.Z.data_ptr(),
                self.tmp_kd.data_ptr(), self.d_params, self.w_params, self.c_ws.data_ptr(),
                self.c_ws.numel() * 8, self.lo, min(self.sizes),
                ctypes.addressof(self.peer.struct) if self.c_sharded else None)
            self.fused = bool(self.lib.cdr_aa_fused_applicable(
                min(self.sizes), self.d, k, self.d_params.max_iterations))

    # -- products with K ----------------------------------------------------
    def _samples_sum(self, L, sLi, sLt, flags):
        """tmp_kd = sum over ranks of L_r X_r (k x d): one kernel over peer memory where the
        strip kernel applies, else the local pass followed by an all-reduce."""
        Tl, d, k = self.Tl, self.d, self.k
        if self.peer is not None and self.peer.reduce_samples_allreduce(
                L, sLi, sLt, self.X, Tl, min(self.sizes), d, k, self.tmp_kd, flags=flags):
            return
        be.reduce_samples(L, sLi, sLt, self.X, Tl, d, k, self.tmp_kd, self.ws, flags=flags)
        self.comm.allreduce_sum(self.tmp_kd)

    def _features_to_columns(self, out, flags):
        """(tmp_kd) X' for this rank's samples, placed in (all-gathered into) `out`."""
        Tl, d, k = self.Tl, self.d, self.k
        if not self.comm.enabled:
            be.reduce_features(self.tmp_kd, self.X, Tl, d, k, out, self.ws, flags)
            return
        be.reduce_features(self.tmp_kd, self.X, Tl, d, k, self.loc_kt, self.ws, flags)
        self.comm.allgather_columns(self.loc_kt, out, self.sizes, self.gather_scratch)

    def apply_left(self, L, out, flags):
        """out = L K for a k x T matrix L (dictionary.dot(K) or (L X) X')."""
        Tl, k = self.Tl, self.k
        if self.mode == 'feature':
            Lloc = L[:, self.lo:] if self.lo else L
            self._samples_sum(Lloc, L.stride(0), 1, flags)
            self._features_to_columns(out, flags)
        else:
            be.reduce_samples(L, L.stride(0), 1, self.X, Tl, Tl, k, out, self.ws, flags=flags)

    def apply_right(self, flags):
        """KZt = (K Z)' (K.dot(weights) or X (X' Z))."""
        Tl, k = self.Tl, self.k
        if self.mode == 'feature':
            self._samples_sum(self.Z, 1, k, flags)
            self._features_to_columns(self.KZt, flags)
        else:
            self.Zt[:, :Tl].copy_(self.Z.t())
            be.reduce_features(self.Zt, self.X, Tl, Tl, k, self.KZt, self.ws, flags)

    # -- small products -----------------------------------------------------
    def _desc(self, A, B, out):
        return (A, A.stride(0), 1, self.k, B, B.stride(0), 1, self.k, self.T, out, 1.0, 0)

    def _desc_ZtZ(self):
        k = self.k
        return (self.Z, 1, k, k, self.Z, 1, k, k, self.Tl, self.ZtZ, 1.0, 0)

    def _cost_check(self, stage, end):
        be.check(self.lib.cdr_aa_cost_check(ctypes.byref(self.buf), stage, int(end),
                                            be.stream_ptr()), 'cdr_aa_cost_check')

    # -- pieces of the loop -------------------------------------------------
    def precompute(self, need_right=True):
        """archetypal_analysis.py:406-417 / :541-556."""
        fl = self.state.ptr
        self.apply_left(self.C, self.CK, fl)
        if need_right:
            self.apply_right(fl)
        be.small_gram([self._desc_ZtZ(), self._desc(self.CK, self.C, self.CKCt),
                       self._desc(self.C, self.KZt, self.CKZ)], self.ws, fl)
        self.comm.allreduce_sum(self.ZtZ)

    def initial_cost(self):
        if self.c_loop and not self.c_sharded:
            be.check(self.lib.cdr_aa_prepare_enqueue(ctypes.byref(self.problem),
                                                     be.stream_ptr()), 'cdr_aa_prepare_enqueue')
            self._first_dictionary_update = False
            return
        self.precompute()
        self._cost_check(0, False)
        if self.c_sharded:
            # what cdr_aa_prepare_enqueue does next on one GPU: the projection of the start
            # (spg.py:146-148) and old_cost
            self._project_start()
            be.check(self.lib.cdr_loop_begin(self.state.ptr, be.stream_ptr()), 'cdr_loop_begin')

    def _project_start(self):
        """spg() starts from project(x0) (spg.py:146-148).  A custom start may only be
        feasible to np.isclose accuracy, so C K is rebuilt for the projected iterate once;
        later iterates are feasible by construction."""
        fl = self.state.ptr
        be.check(self.lib.cdr_simplex_project_rows(
            self.C.data_ptr(), self.C.data_ptr(), self.k, self.T, self.ldt, self.ldt, fl,
            be.stream_ptr()), 'cdr_simplex_project_rows')
        self.apply_left(self.C, self.CK, fl)
        be.small_gram([self._desc(self.CK, self.C, self.CKCt)], self.ws, fl)
        self._first_dictionary_update = False

    def spg_iteration(self, last):
        """One iteration of spg.py:165-281 on the dictionary."""
        buf, p, s = ctypes.byref(self.buf), ctypes.byref(self.d_params), be.stream_ptr
        fl = self.state.ptr
        be.check(self.lib.cdr_aa_spg_direction(buf, p, s()), 'cdr_aa_spg_direction')
        self.apply_left(self.D, self.DK, fl)
        be.small_gram([self._desc(self.CK, self.D, self.G01),
                       self._desc(self.DK, self.D, self.G11)], self.ws, fl)
        be.check(self.lib.cdr_aa_spg_linesearch(buf, p, s()), 'cdr_aa_spg_linesearch')
        if not last:
            be.check(self.lib.cdr_aa_spg_update(buf, p, 1, s()), 'cdr_aa_spg_update')

    def dictionary_step(self, stage=2, end=False):
        """_update_kernel_aa_dictionary / _update_aa_dictionary and the recomputes
        that follow it (archetypal_analysis.py:304-341, 474-487, 611-630)."""
        fl = self.state.ptr
        buf, p = ctypes.byref(self.buf), ctypes.byref(self.d_params)
        if self._first_dictionary_update:
            self._project_start()
        be.check(self.lib.cdr_aa_spg_begin(buf, p, be.stream_ptr()), 'cdr_aa_spg_begin')
        max_it = self.d_params.max_iterations
        if max_it <= _MAX_UNROLLED_SPG:
            for n in range(max_it):
                self.spg_iteration(last=(n == max_it - 1))
        else:
            n = 0
            while n < max_it:
                burst = min(16, max_it - n)
                for i in range(burst):
                    self.spg_iteration(last=(n + i == max_it - 1))
                n += burst
                st = self.state.read()
                if st.done or not st.spg_active:
                    break
        be.small_gram([self._desc(self.CK, self.C, self.CKCt),
                       self._desc(self.C, self.KZt, self.CKZ)], self.ws, fl)
        if stage is not None:
            self._cost_check(stage, end)

    def weights_step(self, stage=3, end=True):
        """_update_kernel_aa_weights and the recomputes that follow it
        (archetypal_analysis.py:369-396, 489-503, 632-652)."""
        fl = self.state.ptr
        CKloc = self.CK[:, self.lo:] if self.lo else self.CK
        be.quad_simplex_spg_batched(self.CKCt, self.alpha, CKloc, 1, self.ldt, self.Z,
                                    self.Tl, self.k, self.w_params, flags=fl)
        self.apply_right(fl)
        be.small_gram([self._desc_ZtZ(), self._desc(self.C, self.KZt, self.CKZ)], self.ws, fl)
        self.comm.allreduce_sum(self.ZtZ)
        if stage is not None:
            self._cost_check(stage, end)

    def scale_factors_step(self):
        """_update_kernel_aa_scale_factors (archetypal_analysis.py:243-258): the generic spg()
        on the k-vector alpha over the box [1 - delta, 1 + delta], from three k x k
        statistics -- one warp on the device (csrc/aa_steps.cu), no host round trip."""
        be.check(self.lib.cdr_aa_scale_factors_step(
            ctypes.byref(self.buf), ctypes.byref(self.s_params), float(self.delta),
            be.stream_ptr()), 'cdr_aa_scale_factors_step')
        self._cost_check(1, False)

    def iteration(self):
        if self.c_loop:
            be.check(self.lib.cdr_aa_iterate_enqueue(ctypes.byref(self.problem),
                                                     be.stream_ptr()), 'cdr_aa_iterate_enqueue')
            return
        be.check(self.lib.cdr_loop_begin(self.state.ptr, be.stream_ptr()), 'cdr_loop_begin')
        if self.update_scale_factors and self.delta != 0:
            self.scale_factors_step()
        if self.update_dictionary:
            self.dictionary_step(end=not self.update_weights)
        if self.update_weights:
            self.weights_step()

    def graph_capturable(self):
        if self.update_dictionary and self.d_params.max_iterations > _MAX_UNROLLED_SPG:
            return False
        return True

    # -- driver -------------------------------------------------------------
    def run(self, verbose=0, use_graph=None, label='AA'):
        torch = be.torch_mod()
        if use_graph is None:
            use_graph = not be.graphs_disabled()
        use_graph = (use_graph and self.graph_capturable() and not verbose and
                     (not self.comm.enabled or be.graph_collectives()))
        be.trace('aa: engine ready')
        self.initial_cost()
        be.trace('aa: initial cost')
        max_it = self.state.max_iterations
        if verbose:
            print("*** {}: n_components = {:d} ***".format(label, self.k))
            print('{:<12s} | {:<13s} | {:<13s}'.format('Iteration', 'Cost', 'Cost delta'))
            print(80 * '-')
        start = time.perf_counter()
        self.iteration()
        launched = 1
        st = self.state.read()
        be.trace('aa: first iteration')
        if verbose:
            print('{:12d} | {: 12.6e} | {: 12.6e}'.format(st.n_iter, st.cost, st.cost - st.old_cost))
        chunk = 1
        graph = None
        graph_after = be.graph_after(self.c_loop)
        while not st.done and launched < max_it:
            if use_graph and graph is None and launched >= graph_after:
                graph = be.capture_graph(self.iteration)
                be.trace('aa: graph capture')
            n = min(chunk, max_it - launched)
            for _ in range(n):
                if graph is not None:
                    graph.replay()
                else:
                    self.iteration()
            launched += n
            st = self.state.read()
            if verbose:
                print('{:12d} | {: 12.6e} | {: 12.6e}'.format(
                    st.n_iter, st.cost, st.cost - st.old_cost))
            elif chunk < 32 and (graph is not None or self.c_loop):
                # (past `done` the kernels of an iteration return at once, so a chunk may
                # overshoot; the host-paced general sequence is not launched ahead)
                chunk *= 2
        torch.cuda.synchronize()
        elapsed = time.perf_counter() - start
        st = self.state.read()
        if self.peer is not None:
            self.peer.check()          # a wait inside a peer collective timed out
        be.trace('aa: remaining iterations')
        if st.error_stage:
            raise RuntimeError('factorization cost increased after {} update'.format(
                _STAGES[st.error_stage]))
        if st.spg_warnings & 1:
            warnings.warn('step size below tolerance in SPG line search', UserWarning)
        if st.spg_warnings & 2:
            warnings.warn('maximum number of function evaluations exceeded in SPG', UserWarning)
        if st.spg_warnings & 4:
            warnings.warn('maximum number of iterations exceeded in SPG', UserWarning)
        self.cost = st.cost
        self.n_iter = st.n_iter - 1            # 0-based loop index, like the reference
        self.avg_time_per_iter = elapsed / max(st.n_iter, 1)
        self.cost_deltas = self.state.cost_deltas[:st.n_iter].cpu().numpy().tolist()
        return self

    def weights(self):
        return be.to_host(self.Z)

    def dictionary(self):
        return be.to_host(self.C, self.k, self.T)

    def scale_factors(self):
        return self.alpha.cpu().numpy()


# ---------------------------------------------------------------------------
# initialisation (archetypal_analysis.py:51-164)
# ---------------------------------------------------------------------------

def _initialize_kernel_aa_dictionary_random(kernel, n_components, random_state=None):
    rng = check_random_state(random_state)
    return right_stochastic_matrix((n_components, kernel.shape[0]), random_state=rng)


def _initialize_kernel_aa_weights_random(kernel, n_components, random_state=None):
    rng = check_random_state(random_state)
    return right_stochastic_matrix((kernel.shape[0], n_components), random_state=rng)


def _initialize_kernel_aa_scale_factors_random(n_components, delta=0, random_state=None):
    rng = check_random_state(random_state)
    if delta != 0:
        return rng.uniform(low=(1 - delta), high=(1 + delta), size=(n_components,))
    return np.ones(n_components)


class _LazyKernel:
    """Shape/dtype stand-in for K = X X' that builds the device Gram matrix only when
    an initialisation actually reads it (archetypal_analysis.py:1032 always forms K,
    but init='random' uses nothing but its shape)."""

    def __init__(self, data, data_device=None, comm=None, n_samples=None):
        self._data = data
        self._data_device = data_device
        self._comm = comm if comm is not None and comm.enabled else None
        n_samples = data.shape[0] if n_samples is None else n_samples
        self.shape = (n_samples, n_samples)
        self.dtype = np.dtype(np.float64)
        self._K = None

    def device(self):
        if self._K is None:
            def build():
                Xd = self._data_device if self._data_device is not None else \
                    be.to_device_padded(self._data)
                if self._comm is not None:
                    # sample-sharded fit: gather the rows once, deal the slabs of K to the ranks
                    sizes = shard_sizes(self.shape[0], self._comm.world)
                    Xd = self._comm.allgather_row_blocks(Xd, sizes)
                    return be.gram(Xd, self.shape[0], self._data.shape[1], self._comm)
                return be.gram(Xd, self._data.shape[0], self._data.shape[1])
            # shared by all restarts of fit_aa_model while the data are resident
            self._K = be.resident_derived(self._data, 'gram', build)
        return self._K


def _kernel_device(kernel):
    if isinstance(kernel, _LazyKernel):
        return kernel.device()
    return be.to_device_padded(np.asarray(kernel, dtype=np.float64))


def _initialize_kernel_aa_dictionary_furthest_sum(
        kernel, n_components, start_index=None, n_extra_steps=10,
        exclude=None, random_state=None):
    rng = check_random_state(random_state)
    n_samples = kernel.shape[0]
    if start_index is None:
        start_index = rng.randint(n_samples)
    if exclude is None:
        exclude = np.array([], dtype='i8')
    K = _kernel_device(kernel)
    D = dissimilarity_from_gram_device(K, n_samples)
    selected = furthest_sum_device(D, n_samples, n_components, start_index, exclude,
                                   n_extra_steps)
    dictionary = np.zeros((n_components, n_samples), dtype=np.float64)
    for i in range(n_components):
        dictionary[i, selected[i]] = 1
    return dictionary


def _initialize_kernel_aa_dictionary(kernel, n_components, init='furthest_sum',
                                     random_state=None, **kwargs):
    if init is None:
        init = 'furthest_sum'
    if init == 'furthest_sum':
        return _initialize_kernel_aa_dictionary_furthest_sum(
            kernel, n_components, start_index=kwargs.get('start_index', None),
            n_extra_steps=kwargs.get('n_extra_steps', 10),
            exclude=kwargs.get('exclude', None), random_state=random_state)
    if init == 'random':
        return _initialize_kernel_aa_dictionary_random(
            kernel, n_components, random_state=random_state)
    raise ValueError('Invalid init parameter: got %r instead of one of %r' %
                     (init, INITIALIZATION_METHODS))


def _initialize_kernel_aa_weights(kernel, n_components, init='furthest_sum',
                                  random_state=None, **kwargs):
    if init is None:
        init = 'furthest_sum'
    if init in ('furthest_sum', 'random'):
        return _initialize_kernel_aa_weights_random(
            kernel, n_components, random_state=random_state)
    raise ValueError('Invalid init parameter: got %r instead of one of %r' %
                     (init, INITIALIZATION_METHODS))


def _initialize_kernel_aa(kernel, n_components, init='furthest_sum',
                          random_state=None, **kwargs):
    """Dictionary first, then weights (the RNG draw order of archetypal_analysis.py:151-164)."""
    if init is None:
        init = 'furthest_sum'
    rng = check_random_state(random_state)
    dictionary = _initialize_kernel_aa_dictionary(
        kernel, n_components, init=init, random_state=rng, **kwargs)
    weights = _initialize_kernel_aa_weights(
        kernel, n_components, init=init, random_state=rng)
    return dictionary, weights


# ---------------------------------------------------------------------------
# function-level API (what the reference's tests import)
# ---------------------------------------------------------------------------

def _engine(data, weights, dictionary, alpha, mode, **kw):
    return _AaEngine(np.asarray(data, dtype=np.float64), np.asarray(weights, dtype=np.float64),
                     np.asarray(dictionary, dtype=np.float64), alpha, mode, **kw)


def _kernel_aa_cost(K, weights, dictionary, alpha):
    """Evaluate kernel AA cost function (archetypal_analysis.py:200-217)."""
    eng = _engine(K, weights, dictionary, alpha, 'kernel')
    eng.initial_cost()
    return eng.state.read().cost


def _dictionary_problem(data, dictionary, mode, trace, KZD, DZtZD):
    """Engine set up for f / df of the dictionary sub-problem with the scaling already
    folded into KZD = K Z D and DZtZD (alpha = 1)."""
    dictionary = np.asarray(dictionary, dtype=np.float64)
    k, T = dictionary.shape
    eng = _engine(data, np.zeros((T, k)), dictionary, np.ones(k), mode, trace_data=trace)
    eng.KZt[:, :T].copy_(be.to_device(np.ascontiguousarray(np.asarray(KZD).T)))
    eng.ZtZ.copy_(be.to_device(DZtZD))
    eng.apply_left(eng.C, eng.CK, None)
    be.small_gram([eng._desc(eng.CK, eng.C, eng.CKCt), eng._desc(eng.C, eng.KZt, eng.CKZ)],
                  eng.ws)
    return eng


def _dictionary_cost(data, dictionary, mode, trace, KZD, DZtZD):
    eng = _dictionary_problem(data, dictionary, mode, trace, KZD, DZtZD)
    out = be.zeros(1)
    be.check(eng.lib.cdr_aa_dictionary_cost(ctypes.byref(eng.buf), float(trace), out.data_ptr(),
                                            be.stream_ptr()), 'cdr_aa_dictionary_cost')
    return float(out.item())


def _dictionary_gradient(data, dictionary, mode, KZD, DZtZD):
    eng = _dictionary_problem(data, dictionary, mode, 0.0, KZD, DZtZD)
    be.check(eng.lib.cdr_aa_gradient(ctypes.byref(eng.buf), be.stream_ptr()), 'cdr_aa_gradient')
    return be.to_host(eng.G, eng.k, eng.T)


def _aa_dictionary_cost(X, dictionary, trace_XXt, XXtZD, DZtZD):
    """archetypal_analysis.py:261-270 (divided by k)."""
    return _dictionary_cost(X, dictionary, 'feature', trace_XXt, XXtZD, DZtZD)


def _kernel_aa_dictionary_cost(K, dictionary, trace_K, KZD, DZtZD):
    """archetypal_analysis.py:273-281 (divided by k)."""
    return _dictionary_cost(K, dictionary, 'kernel', trace_K, KZD, DZtZD)


def _aa_dictionary_gradient(X, dictionary, XXtZD, DZtZD):
    """archetypal_analysis.py:293-301 (divided by T)."""
    return _dictionary_gradient(X, dictionary, 'feature', XXtZD, DZtZD)


def _kernel_aa_dictionary_gradient(K, dictionary, KZD, DZtZD):
    """archetypal_analysis.py:284-290 (divided by k)."""
    return _dictionary_gradient(K, dictionary, 'kernel', KZD, DZtZD)


def _update_dictionary(data, dictionary, alpha, trace, KZ, ZtZ, mode, kwargs):
    dictionary = np.asarray(dictionary, dtype=np.float64)
    k, T = dictionary.shape
    eng = _engine(data, np.zeros((T, k)), dictionary, alpha, mode, trace_data=trace,
                  dictionary_solver_kwargs=kwargs, max_iterations=1)
    eng.KZt[:, :T].copy_(be.to_device(np.ascontiguousarray(np.asarray(KZ).T)))
    eng.ZtZ.copy_(be.to_device(ZtZ))
    eng.apply_left(eng.C, eng.CK, None)
    be.small_gram([eng._desc(eng.CK, eng.C, eng.CKCt)], eng.ws)
    eng.dictionary_step(stage=None)
    st = eng.state.read()
    if st.spg_warnings & 1:
        warnings.warn('step size below tolerance in SPG line search', UserWarning)
    if st.spg_warnings & 2:
        warnings.warn('maximum number of function evaluations exceeded in SPG', UserWarning)
    return eng.dictionary()


def _update_kernel_aa_dictionary(K, dictionary, alpha, trace_K, KZ, ZtZ, **kwargs):
    """Update dictionary for kernel AA (archetypal_analysis.py:304-321)."""
    return _update_dictionary(K, dictionary, alpha, trace_K, KZ, ZtZ, 'kernel', kwargs)


def _update_aa_dictionary(X, dictionary, alpha, trace_XXt, XXtZ, ZtZ, **kwargs):
    """Update dictionary for AA (archetypal_analysis.py:324-341)."""
    return _update_dictionary(X, dictionary, alpha, trace_XXt, XXtZ, ZtZ, 'feature', kwargs)


def _solve_weights(CKCt, alpha, CK, weights, params):
    """Batched per-sample QPs (archetypal_analysis.py:344-366); CK is k x n."""
    CK = np.ascontiguousarray(CK, dtype=np.float64)
    k, n = CK.shape
    dA = be.to_device(CKCt)
    dB = be.to_device(CK)
    dZ = be.to_device(weights)
    dalpha = None if alpha is None else be.to_device(np.asarray(alpha, dtype=np.float64))
    be.quad_simplex_spg_batched(dA, dalpha, dB, 1, n, dZ, n, k, params)
    return be.to_host(dZ)


def _update_kernel_aa_weights(weights, alpha, CK, CKCt, **solver_kwargs):
    """Update weights for kernel AA (archetypal_analysis.py:369-396)."""
    return _solve_weights(CKCt, alpha, CK, weights, be.make_spg_params(solver_kwargs))


class _ShapeOnly:
    """Stand-in for a host matrix that only lives on the device."""

    def __init__(self, shape):
        self.shape = tuple(shape)


def _iterate(data, weights, dictionary, alpha, mode, delta, update_weights,
             update_dictionary, update_scale_factors, tolerance, max_iterations, verbose,
             kwargs):
    be.trace('aa: enter _iterate')
    grad_scale = None
    if mode == 'feature' and kwargs.get('formulation', 'stream') == 'gram':
        # Opt-in Gram-space formulation of the *feature-space* problem: K = X X' is built once
        # on the device (T x T, L2-resident at HadISST shape) and every later product is a pass
        # over K instead of two passes over X.  Same cost / gradient scaling as _iterate_aa
        # (gradient / T); only the rounding of (C X) X' vs C (X X') differs.
        if kwargs.get('comm') is not None and kwargs['comm'].enabled:
            raise NotImplementedError("formulation='gram' is single-GPU")
        data = np.asarray(data, dtype=np.float64)
        T, d = data.shape
        Xd = kwargs.get('data_device')
        if Xd is None:
            Xd = be.to_device_padded(data)
        trace = kwargs.get('trace_data')
        if trace is None:
            trace = float(be.frobenius_sq(Xd, T, d).item())
        kwargs = dict(kwargs, data_device=be.gram(Xd, T, d), trace_data=trace)
        data, mode, grad_scale = _ShapeOnly((T, T)), 'kernel', 1.0 / T
    elif kwargs.get('formulation', 'stream') not in ('stream', 'gram'):
        raise ValueError("formulation must be 'stream' or 'gram'")
    eng = _AaEngine(
        data if isinstance(data, _ShapeOnly) else np.asarray(data, dtype=np.float64),
        np.asarray(weights, dtype=np.float64), np.asarray(dictionary, dtype=np.float64), alpha,
        mode, grad_scale=grad_scale, delta=delta, tolerance=tolerance,
        max_iterations=max_iterations,
        stopping_criterion=kwargs.get('stopping_criterion', 'abs_delta_f'),
        require_monotonic_cost_decrease=kwargs.get('require_monotonic_cost_decrease', True),
        weights_solver_kwargs=kwargs.get('weights_solver_kwargs', {}),
        dictionary_solver_kwargs=kwargs.get('dictionary_solver_kwargs', {}),
        scale_factors_solver_kwargs=kwargs.get('scale_factors_solver_kwargs', {}),
        update_weights=update_weights, update_dictionary=update_dictionary,
        update_scale_factors=update_scale_factors, trace_data=kwargs.get('trace_data'),
        data_device=kwargs.get('data_device'), comm=kwargs.get('comm'))
    eng.run(verbose=verbose, label='Kernel AA' if mode == 'kernel' else 'AA')
    new_weights = eng.weights() if update_weights else weights
    new_dictionary = eng.dictionary() if update_dictionary else dictionary
    new_alpha = eng.scale_factors() if (update_scale_factors and delta != 0) else alpha
    be.trace('aa: results to host')
    return (new_weights, new_dictionary, new_alpha, eng.cost, eng.n_iter,
            eng.avg_time_per_iter, eng.cost_deltas)


def _iterate_kernel_aa(K, weights, dictionary, alpha, delta=0,
                       update_weights=True, update_dictionary=True,
                       update_scale_factors=True, tolerance=1e-6,
                       max_iterations=1000, verbose=0, **kwargs):
    """Iteratively update kernel AA parameters until convergence is reached
    (archetypal_analysis.py:399-531).  Returns ``(weights, dictionary, alpha, cost,
    n_iter, avg_time_per_iter, cost_deltas)``."""
    return _iterate(K, weights, dictionary, alpha, 'kernel', delta, update_weights,
                    update_dictionary, update_scale_factors, tolerance, max_iterations,
                    verbose, kwargs)


def _iterate_aa(X, weights, dictionary, alpha, delta=0,
                update_weights=True, update_dictionary=True,
                update_scale_factors=True, tolerance=1e-6,
                max_iterations=1000, verbose=0, **kwargs):
    """Iteratively update AA parameters until convergence is reached
    (archetypal_analysis.py:534-670)."""
    return _iterate(X, weights, dictionary, alpha, 'feature', delta, update_weights,
                    update_dictionary, update_scale_factors, tolerance, max_iterations,
                    verbose, kwargs)


# ---------------------------------------------------------------------------
# estimators
# ---------------------------------------------------------------------------

class _AaBase():
    """Constructor and parameter handling shared by KernelAA and ArchetypalAnalysis
    (archetypal_analysis.py:743-771, 999-1024)."""

    def __init__(self, n_components, delta=0, init=None,
                 tolerance=1e-6, max_iterations=1000, verbose=0,
                 random_state=None, **kwargs):
        self.n_components = n_components
        self.delta = delta
        self.init = init
        self.tolerance = tolerance
        self.max_iterations = max_iterations
        self.verbose = verbose
        self.random_state = check_random_state(random_state)
        self.require_monotonic_cost_decrease = kwargs.get(
            'require_monotonic_cost_decrease', True)
        self.stopping_criterion = kwargs.get('stopping_criterion', 'abs_delta_f')
        # 'stream' (default: products with X, as the reference) or 'gram' (ArchetypalAnalysis
        # only: one Gram matrix up front, then passes over K)
        self.formulation = kwargs.get('formulation', 'stream')
        self.weights = None
        self.dictionary = None
        self.alpha = None
        self.cost = 0
        self.n_iter = 0
        self.avg_time_per_iter = 0
        self.cost_deltas = None
        self.weights_solver_kwargs = kwargs.get('weights_solver_kwargs', {})
        self.dictionary_solver_kwargs = kwargs.get('dictionary_solver_kwargs', {})
        self.scale_factors_solver_kwargs = kwargs.get('scale_factors_solver_kwargs', {})

    def _check_params(self, default_components):
        if self.n_components is None:
            self.n_components = default_components
        if not isinstance(self.n_components, INTEGER_TYPES) or self.n_components <= 0:
            raise ValueError('Number of components must be a positive integer;'
                             ' got (n_components=%r)' % self.n_components)
        if not isinstance(self.max_iterations, INTEGER_TYPES) or self.max_iterations <= 0:
            raise ValueError('Maximum number of iterations must be a positive '
                             'integer; got (max_iterations=%r)' % self.max_iterations)
        if not isinstance(self.tolerance, numbers.Number) or self.tolerance < 0:
            raise ValueError('Tolerance for stopping criteria must be '
                             'positive; got (tolerance=%r)' % self.tolerance)

    def _initial_factors(self, kernel, n_samples, whom, dictionary, weights, alpha,
                         update_dictionary, update_weights, kwargs):
        k = self.n_components
        if self.init == 'custom':
            _check_init_weights(weights, (n_samples, k), '%s (input weights)' % whom)
            _check_init_dictionary(dictionary, (k, n_samples), '%s (input dictionary)' % whom)
            _check_init_scale_factors(alpha, self.delta, (k,), '%s (input scale factors)' % whom)
        elif not update_dictionary and update_weights:
            _check_init_dictionary(dictionary, (k, n_samples), '%s (input dictionary)' % whom)
            weights = _initialize_kernel_aa_weights(
                kernel, k, init=self.init, random_state=self.random_state, **kwargs)
        elif update_dictionary and not update_weights:
            _check_init_weights(weights, (n_samples, k), '%s (input weights)' % whom)
            dictionary = _initialize_kernel_aa_dictionary(
                kernel, k, init=self.init, random_state=self.random_state, **kwargs)
        else:
            dictionary, weights = _initialize_kernel_aa(
                kernel, k, init=self.init, random_state=self.random_state, **kwargs)
        if alpha is None:
            alpha = _initialize_kernel_aa_scale_factors_random(
                k, delta=self.delta, random_state=self.random_state)
        else:
            _check_init_scale_factors(alpha, self.delta, (k,),
                                      '%s (input scale factors)' % whom)
        self.weights = np.array(weights, dtype=np.float64)
        self.dictionary = np.array(dictionary, dtype=np.float64)
        self.alpha = np.array(alpha, dtype=np.float64)

    def _run(self, iterate, data, update_weights, update_dictionary, update_scale_factors,
             **extra):
        (self.weights, self.dictionary, self.alpha, cost, n_iter, avg_time_per_iter,
         cost_deltas) = iterate(
            data, self.weights, self.dictionary, self.alpha, delta=self.delta,
            update_weights=update_weights, update_dictionary=update_dictionary,
            update_scale_factors=update_scale_factors, tolerance=self.tolerance,
            max_iterations=self.max_iterations, verbose=self.verbose,
            require_monotonic_cost_decrease=self.require_monotonic_cost_decrease,
            stopping_criterion=self.stopping_criterion,
            weights_solver_kwargs=self.weights_solver_kwargs,
            dictionary_solver_kwargs=self.dictionary_solver_kwargs,
            scale_factors_solver_kwargs=self.scale_factors_solver_kwargs, **extra)
        if n_iter == self.max_iterations and self.tolerance > 0:
            warnings.warn('Maximum number of iterations %d reached.' %
                          self.max_iterations, UserWarning)
        return cost, n_iter, avg_time_per_iter, cost_deltas


class KernelAA(_AaBase):
    """Kernel archetypal analysis: drop-in for archetypal_analysis.py:673-910."""

    def _kernel_aa(self, kernel, dictionary=None, weights=None, alpha=None,
                   update_dictionary=True, update_weights=True,
                   update_scale_factors=True, **kwargs):
        """Perform kernel archetypal analysis (archetypal_analysis.py:773-857)."""
        kernel = np.asarray(kernel)
        n_samples = kernel.shape[0]
        if kernel.shape[1] != n_samples:
            raise ValueError('Expected square kernel matrix in %s. '
                             'Got shape %s' % ('kernel_aa', kernel.shape))
        self._check_params(n_samples)
        self._initial_factors(kernel, n_samples, '_kernel_aa', dictionary, weights, alpha,
                              update_dictionary, update_weights, kwargs)
        return self._run(_iterate_kernel_aa, kernel, update_weights, update_dictionary,
                         update_scale_factors)

    def fit_transform(self, data, dictionary=None, weights=None, alpha=None, **kwargs):
        """Perform kernel archetypal analysis and return transformed data."""
        cost_, n_iter_, avg_time_per_iter_, cost_deltas_ = self._kernel_aa(
            data, dictionary=dictionary, weights=weights, alpha=alpha, **kwargs)
        self.cost = cost_
        self.n_iter = n_iter_
        self.avg_time_per_iter = avg_time_per_iter_
        self.cost_deltas = cost_deltas_
        return self.weights

    def fit(self, kernel, **kwargs):
        """Perform kernel archetypal analysis on given kernel."""
        self.fit_transform(kernel, **kwargs)
        return self


class ArchetypalAnalysis(_AaBase):
    """Standard archetypal analysis: drop-in for archetypal_analysis.py:913-1215."""

    def __init__(self, n_components, delta=0, init=None,
                 tolerance=1e-6, max_iterations=1000, verbose=0,
                 random_state=None, **kwargs):
        super().__init__(n_components, delta=delta, init=init, tolerance=tolerance,
                         max_iterations=max_iterations, verbose=verbose,
                         random_state=random_state, **kwargs)
        self.archetypes = None

    def _aa(self, data, dictionary=None, weights=None, alpha=None,
            update_dictionary=True, update_weights=True,
            update_scale_factors=True, **kwargs):
        """Perform archetypal analysis (archetypal_analysis.py:1026-1106).

        ``comm=Comm()`` runs the sample-sharded fit: ``data`` is this rank's balanced row
        block, initial factors passed in are the full-size ones, the random draws are the
        single-process ones (same seed on every rank); ``self.weights`` ends up holding the
        weights of all rows on every rank and the dictionary is replicated."""
        kwargs = dict(kwargs)
        comm = kwargs.pop('comm', None)
        sharded = comm is not None and comm.enabled
        data = np.asarray(data)
        n_samples = data.shape[0]
        row0 = 0
        if sharded:
            row0, n_samples = comm.local_rows(n_samples)
            comm.sync_random_state(self.random_state)      # rank 0 is authoritative
        self._check_params(data.shape[1])
        data64 = np.ascontiguousarray(data, dtype=np.float64)
        Xd = be.to_device_padded(data64)
        # the reference always forms kernel = data.dot(data.T) (:1032); here it is only
        # built (on the device) if the initialisation reads it
        kernel = _LazyKernel(data64, Xd, comm, n_samples)
        self._initial_factors(kernel, n_samples, '_aa', dictionary, weights, alpha,
                              update_dictionary, update_weights, kwargs)
        del kernel
        if not sharded:
            return self._run(_iterate_aa, data64, update_weights, update_dictionary,
                             update_scale_factors, data_device=Xd,
                             formulation=self.formulation)
        self.weights = np.ascontiguousarray(self.weights[row0:row0 + data.shape[0]])
        result = self._run(_iterate_aa, data64, update_weights, update_dictionary,
                           update_scale_factors, data_device=Xd,
                           formulation=self.formulation, comm=comm)
        self.weights = comm.allgather_rows(self.weights)
        return result

    def fit_transform(self, data, dictionary=None, weights=None, alpha=None, **kwargs):
        """Perform archetypal analysis and return transformed data
        (archetypal_analysis.py:1108-1149)."""
        cost_, n_iter_, avg_time_per_iter_, cost_deltas_ = self._aa(
            data, dictionary=dictionary, weights=weights, alpha=alpha, **kwargs)
        self.cost = cost_
        if self.delta != 0:
            self.dictionary = np.dot(np.diag(self.alpha), self.dictionary)
        self.archetypes = self._archetypes(np.asarray(data, dtype=np.float64),
                                           kwargs.get('comm'))
        self.n_iter = n_iter_
        self.avg_time_per_iter = avg_time_per_iter_
        self.cost_deltas = cost_deltas_
        return self.weights

    def _archetypes(self, data, comm=None):
        """archetypes = dictionary.dot(data) (archetypal_analysis.py:1144); with row-sharded
        data each rank contracts its own columns of the dictionary and the partial k x d
        products are summed over ranks."""
        T, d = data.shape
        k = self.dictionary.shape[0]
        sharded = comm is not None and comm.enabled
        dictionary = self.dictionary
        if sharded:
            row0 = comm.local_rows(T)[0]
            dictionary = np.ascontiguousarray(dictionary[:, row0:row0 + T])
        Xd = be.to_device_padded(data)
        C = be.to_device_padded(dictionary)
        out = be.zeros(k, Xd.stride(0))
        ws = be.Workspace(T, d, k)
        be.reduce_samples(C, C.stride(0), 1, Xd, T, d, k, out, ws)
        if sharded:
            comm.allreduce_sum(out)
        return be.to_host(out, k, d)

    def fit(self, data, **kwargs):
        """Perform archetypal analysis on given data."""
        self.fit_transform(data, **kwargs)
        return self

    def transform(self, data):
        """Transform the data according to the fitted factorization
        (archetypal_analysis.py:1151-1199)."""
        data = np.ascontiguousarray(data, dtype=np.float64)
        n_samples, n_features = data.shape
        k = self.n_components
        # max_iterations of the QPs is the *outer* limit here, as in the reference (:1193)
        kw = dict(self.weights_solver_kwargs)
        kw['max_iterations'] = self.max_iterations
        params = be.make_spg_params(kw)

        A = be.to_device_padded(self.archetypes)
        Xd = be.to_device_padded(data)
        lda = A.stride(0)
        ldt = be.round_up(n_samples)
        CKCt = be.zeros(k, k)
        CK = be.zeros(k, ldt)
        ws = be.Workspace(n_samples, n_features, k)
        be.small_gram([(A, lda, 1, k, A, lda, 1, k, n_features, CKCt, 1.0, 0)], ws)
        be.reduce_features(A, Xd, n_samples, n_features, k, CK, ws)

        initial_weights = right_stochastic_matrix(
            (n_samples, k), random_state=self.random_state)
        Z = be.to_device(initial_weights)
        be.quad_simplex_spg_batched(CKCt, None, CK, 1, ldt, Z, n_samples, k, params)
        self.weights = be.to_host(Z)

        out = be.zeros(1)
        part = be.zeros(n_samples)
        be.check(be.library().cdr_residual_sq(
            Xd.data_ptr(), Xd.stride(0), n_samples, n_features, Z.data_ptr(), k, A.data_ptr(),
            lda, out.data_ptr(), part.data_ptr(), be.stream_ptr()), 'cdr_residual_sq')
        cost = 0.5 * float(out.item()) / n_samples
        return self.weights, cost

    def inverse_transform(self, weights):
        """Transform data back into its original space."""
        return weights.dot(self.archetypes)
