"""GPNH-regularised convex coding (reference ``gpnh_convex_coding.py``).

Same public surface as the reference -- ``GPNHConvexCoding`` plus the private
helpers its tests import (``_gpnh_cost``, ``_iterate_gpnh_convex_coding``,
``_update_gpnh_dictionary``, ``_update_gpnh_weights``) -- with NumPy arrays in
and out.  The alternating loop itself runs on the GPU: the data matrix is
uploaded once, all factors and k x k statistics stay resident, the convergence
and monotonicity tests run on the device, and one outer iteration is replayed
from a CUDA graph.
"""

import ctypes
import numbers
import time
import warnings

import numpy as np
from sklearn.utils import check_array, check_random_state

from . import _backend as be
from ._dist import Comm
from .furthest_sum import dissimilarity_from_gram_device, furthest_sum_device
from .stochastic_matrices import right_stochastic_matrix
from .validation_utils import check_array_shape, check_unit_axis_sums

INTEGER_TYPES = (numbers.Integral, np.integer)

INITIALIZATION_METHODS = (None, 'random', 'furthest_sum',)

_STAGES = {1: 'scale factors', 2: 'dictionary', 3: 'weights'}


def _check_init_weights(weights, shape, whom):
    weights = check_array(weights)
    check_array_shape(weights, shape, whom)
    check_unit_axis_sums(weights, whom, axis=1)


def _check_init_dictionary(dictionary, shape, whom):
    dictionary = check_array(dictionary)
    check_array_shape(dictionary, shape, whom)


class _GpnhEngine:
    """Device-resident state of one GPNH fit (gpnh_convex_coding.py:282-402)."""

    def __init__(self, X, weights, dictionary, lambda_W=0.0, tolerance=1e-6,
                 max_iterations=1000, stopping_criterion='abs_delta_f',
                 require_monotonic_cost_decrease=True, weights_solver_kwargs=None,
                 update_dictionary=True, update_weights=True, trace_XtX=None,
                 X_device=None, comm=None):
        torch = be.require_cuda()
        # sample-sharded fit: X / weights hold this rank's rows, the dictionary is replicated
        self.comm = comm if comm is not None else Comm(enabled=False)
        self.T, self.d = X.shape
        self.k = weights.shape[1]
        if self.k > be.MAX_COMPONENTS:
            raise ValueError('n_components > %d is not supported by the B200 build'
                             % be.MAX_COMPONENTS)
        T, d, k = self.T, self.d, self.k
        self.lambda_W = float(lambda_W)
        self.update_dictionary = update_dictionary
        self.update_weights = update_weights
        self.params = be.make_spg_params(weights_solver_kwargs)
        self.X = X_device if X_device is not None else be.to_device_padded(X)
        self.ldx = self.X.stride(0)
        self.ldt = be.round_up(T)
        self.Z = be.to_device(weights)
        # buffers that are summed over ranks live in the symmetric peer region when the
        # peer-memory collectives are on (CDR_PEER_COLLECTIVES=1); plain tensors otherwise
        self.peer = self.comm.setup_peer([(k, self.ldx), (3, k, k)], (k, self.ldx))
        self.WT = self.comm.zeros(k, self.ldx)
        # uploaded as it is (d x k) and transposed on the device: the strided host copy of the
        # 44 000 x 8 dictionary costs ~1 ms, a tenth of the upload of X
        W_host = np.ascontiguousarray(np.asarray(dictionary, dtype=np.float64))
        self.WT[:, :d].copy_(torch.from_numpy(W_host).cuda().t())
        self.XWt = be.zeros(k, self.ldt)
        # the statistics that reduce over samples share one buffer (one all-reduce):
        # Z'Z, (X W)'Z and -- sharded fits only -- the (X W)'Z of the dictionary sub-step
        self.stats = self.comm.zeros(3, k, k)
        self.ZtZ = self.stats[0]
        self.XWtZ = self.stats[1]
        self.XWtZ_dict = self.stats[2]
        self.ZtZ_prev = be.zeros(k, k)
        self.WtW = be.zeros(k, k)
        self.REG = be.zeros(k, k)
        self.P = be.zeros(k, k)
        self.ws = be.Workspace(T, d, k)
        self.state = be.DeviceState(tolerance, max_iterations, stopping_criterion,
                                    require_monotonic_cost_decrease)
        if trace_XtX is None:
            # gpnh_convex_coding.py:302 forms the d x d product only for its trace
            tr = be.frobenius_sq(self.X, T, d)
            self.comm.allreduce_sum(tr)
            trace_XtX = float(tr.item())
        self.state.write_field('trace_data', float(trace_XtX))
        if self.comm.enabled:
            n_tot = torch.tensor([T], dtype=torch.int64, device='cuda')
            self.comm.allreduce_sum(n_tot)
            self.T_total = int(n_tot.item())
            t_min = torch.tensor([-T], dtype=torch.int64, device='cuda')
            self.comm.allreduce_max(t_min)
            self.T_min = -int(t_min.item())  # smallest local T: keeps kernel choices identical
        else:
            self.T_total = self.T_min = T
        self.lib = be.library()
        # full iterations run behind the C entry points cdr_gpnh_prepare_enqueue /
        # cdr_gpnh_iterate_enqueue (three kernels per iteration at streaming shapes, k <= 16):
        # always on a single GPU, and for a sample-sharded fit when the peer-memory collectives
        # are on and the three-kernel path covers the shape
        full = update_dictionary and update_weights
        self.c_sharded = (full and self.comm.enabled and self.peer is not None and
                          bool(self.lib.cdr_gpnh_fused_applicable(self.T_min, d, k)))
        self.c_loop = (full and not self.comm.enabled) or self.c_sharded
        if self.c_loop:
            nbytes = self.lib.cdr_gpnh_workspace_bytes(T, d, k)
            self.c_ws = torch.empty(nbytes // 8 + 1, dtype=torch.float64, device='cuda')
            self.problem = be.GpnhProblem(
                self.X.data_ptr(), self.ldx, T, d, k, self.T_total, self.lambda_W,
                self.Z.data_ptr(), self.WT.data_ptr(), self.XWt.data_ptr(), self.ldt,
                self.ZtZ.data_ptr(), self.XWtZ.data_ptr(), self.WtW.data_ptr(),
                self.REG.data_ptr(), self.P.data_ptr(), self.state.ptr,
                self.state.cost_deltas.data_ptr(), self.params, self.c_ws.data_ptr(),
                self.c_ws.numel() * 8,
                ctypes.addressof(self.peer.struct) if self.c_sharded else None, self.T_min)
            self.fused = bool(self.lib.cdr_gpnh_fused_applicable(self.T_min, d, k))

    # -- small products -----------------------------------------------------
    def _desc_ZtZ(self):
        k, T = self.k, self.T
        return (self.Z, 1, k, k, self.Z, 1, k, k, T, self.ZtZ, 1.0, 0)

    def _desc_WtW(self):
        k, d = self.k, self.d
        return (self.WT, self.ldx, 1, k, self.WT, self.ldx, 1, k, d, self.WtW, 1.0, 0)

    def _desc_XWtZ(self):
        k, T = self.k, self.T
        return (self.XWt, self.ldt, 1, k, self.Z, 1, k, k, T, self.XWtZ, 1.0, 0)

    def _desc_REG(self):
        k, d = self.k, self.d
        return (self.WT, self.ldx, 1, k, self.WT, self.ldx, 1, k, d, self.REG, 1.0, 1)

    def _cost_check(self, stage, end, with_reg, XWtZ=None, ZtZ=None):
        XWtZ = self.XWtZ if XWtZ is None else XWtZ
        ZtZ = self.ZtZ if ZtZ is None else ZtZ
        be.check(self.lib.cdr_gpnh_cost_check(
            self.state.ptr, self.state.cost_deltas.data_ptr(), XWtZ.data_ptr(),
            ZtZ.data_ptr(), self.WtW.data_ptr(),
            self.REG.data_ptr() if with_reg else None, self.k, self.T_total, self.d,
            self.lambda_W, stage, int(end), be.stream_ptr()), 'cdr_gpnh_cost_check')

    # -- pieces of the loop -------------------------------------------------
    def initial_cost(self):
        """gpnh_convex_coding.py:292-314."""
        if self.c_loop and not self.c_sharded:
            be.check(self.lib.cdr_gpnh_prepare_enqueue(ctypes.byref(self.problem),
                                                       be.stream_ptr()),
                     'cdr_gpnh_prepare_enqueue')
            return
        flags = self.state.ptr
        be.reduce_features(self.WT, self.X, self.T, self.d, self.k, self.XWt, self.ws, flags)
        descs = [self._desc_ZtZ(), self._desc_WtW(), self._desc_XWtZ()]
        if self.lambda_W != 0:
            descs.append(self._desc_REG())
        be.small_gram(descs, self.ws, flags)
        self.comm.allreduce_sum(self.stats[:2])
        self._cost_check(0, False, self.lambda_W != 0)
        if self.c_sharded:
            # start of the first iteration, as cdr_gpnh_prepare_enqueue does on one GPU
            be.check(self.lib.cdr_loop_begin(self.state.ptr, be.stream_ptr()), 'cdr_loop_begin')
            be.check(self.lib.cdr_gpnh_solve_matrix(
                self.ZtZ.data_ptr(), self.k, self.T_total, self.d, self.lambda_W,
                self.P.data_ptr(), None, 0, flags, be.stream_ptr()), 'cdr_gpnh_solve_matrix')

    def dictionary_step(self, stage=2, end=False, flags=True):
        """gpnh_convex_coding.py:348-369 (one pass Z'X with the k x k solve folded
        into the epilogue, one pass X W)."""
        fl = self.state.ptr if flags else None
        T, d, k = self.T, self.d, self.k
        be.check(self.lib.cdr_gpnh_solve_matrix(
            self.ZtZ.data_ptr(), k, self.T_total, d, self.lambda_W, self.P.data_ptr(), None, 0,
            fl, be.stream_ptr()), 'cdr_gpnh_solve_matrix')
        # W' = P sum_g Z_g' X_g: one kernel over peer memory where the strip kernel applies,
        # else the local pass followed by an all-reduce
        if self.peer is None or not self.peer.reduce_samples_allreduce(
                self.Z, 1, k, self.X, T, self.T_min, d, k, self.WT, E=self.P, flags=fl):
            be.reduce_samples(self.Z, 1, k, self.X, T, d, k, self.WT, self.ws, E=self.P, flags=fl)
            self.comm.allreduce_sum(self.WT)
        be.reduce_features(self.WT, self.X, T, d, k, self.XWt, self.ws, fl)
        descs = [self._desc_WtW(), self._desc_XWtZ()]
        if self.lambda_W != 0:
            descs.append(self._desc_REG())
        be.small_gram(descs, self.ws, fl)
        if self._defer_dictionary_check(stage, end):
            # sharded full iteration: the k x k all-reduce latency (~30 us on 8 GPUs) is paid
            # once per iteration -- this sub-step's (X W)'Z partial rides along with the
            # statistics of the weights sub-step and both cost checks run after it
            self.XWtZ_dict.copy_(self.XWtZ)
            self.ZtZ_prev.copy_(self.ZtZ)
            return
        self.comm.allreduce_sum(self.XWtZ)
        if stage is not None:
            self._cost_check(stage, end, self.lambda_W != 0)

    def _defer_dictionary_check(self, stage, end):
        return self.comm.enabled and stage == 2 and not end and self.update_weights

    def weights_step(self, stage=3, end=True, flags=True):
        """gpnh_convex_coding.py:371-384."""
        fl = self.state.ptr if flags else None
        be.quad_simplex_spg_batched(self.WtW, None, self.XWt, 1, self.ldt, self.Z, self.T,
                                    self.k, self.params, flags=fl)
        be.small_gram([self._desc_ZtZ(), self._desc_XWtZ()], self.ws, fl)
        if self.comm.enabled and self.update_dictionary and stage == 3:
            self.comm.allreduce_sum(self.stats)
            # deferred check of the dictionary sub-step (its own (X W)'Z, the old Z'Z) ...
            self._cost_check(2, False, self.lambda_W != 0, XWtZ=self.XWtZ_dict,
                             ZtZ=self.ZtZ_prev)
        else:
            self.comm.allreduce_sum(self.stats[:2])
        if stage is not None:
            self._cost_check(stage, end, False)

    def iteration(self):
        if self.c_loop:
            be.check(self.lib.cdr_gpnh_iterate_enqueue(ctypes.byref(self.problem),
                                                       be.stream_ptr()),
                     'cdr_gpnh_iterate_enqueue')
            return
        be.check(self.lib.cdr_loop_begin(self.state.ptr, be.stream_ptr()), 'cdr_loop_begin')
        if self.update_dictionary:
            self.dictionary_step(end=not self.update_weights)
        if self.update_weights:
            self.weights_step()

    # -- driver -------------------------------------------------------------
    def run(self, verbose=0, use_graph=None):
        torch = be.torch_mod()
        if use_graph is None:
            use_graph = not be.graphs_disabled() and (not self.comm.enabled or be.graph_collectives())
        be.trace('gpnh: engine ready')
        self.initial_cost()
        be.trace('gpnh: initial cost')
        max_it = self.state.max_iterations
        start = time.perf_counter()
        launched = 0
        st = None
        if verbose:
            print("*** GPNH convex coding: n_components = {:d} ***".format(self.k))
            print('{:<12s} | {:<13s} | {:<13s}'.format('Iteration', 'Cost', 'Cost delta'))
            print(100 * '-')
        # first iteration eagerly (also warms up every kernel before capture)
        self.iteration()
        launched += 1
        st = self.state.read()
        chunk = 1
        graph = None
        graph_after = be.graph_after(self.c_loop)
        while not st.done and launched < max_it:
            if use_graph and graph is None and not verbose and launched >= graph_after:
                be.trace('gpnh: first iteration')
                graph = be.capture_graph(self.iteration)
                be.trace('gpnh: graph capture')
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
            elif chunk < 32:
                chunk *= 2
        torch.cuda.synchronize()
        elapsed = time.perf_counter() - start
        st = self.state.read()
        if self.peer is not None:
            self.peer.check()          # a wait inside a peer collective timed out
        be.trace('gpnh: remaining iterations')
        if st.error_stage:
            raise RuntimeError('factorization cost increased after {} update'.format(
                _STAGES[st.error_stage]))
        done_iters = max(st.n_iter, 1)
        self.cost = st.cost
        self.n_iter = st.n_iter - 1            # the reference returns the 0-based loop index
        self.avg_time_per_iter = elapsed / done_iters
        self.cost_deltas = self.state.cost_deltas[:st.n_iter].cpu().numpy().tolist()
        return self

    def weights(self):
        return be.to_host(self.Z)

    def dictionary(self):
        # d x k view of the k x d solution, like `sol.T` at gpnh_convex_coding.py:226
        return be.to_host(self.WT, self.k, self.d).T


def _gpnh_regularization(dictionary):
    """Evaluate GPNH regularization term (gpnh_convex_coding.py:179-196)."""
    dictionary = np.asarray(dictionary, dtype=np.float64)
    n_features, n_components = dictionary.shape
    if n_components == 1:
        return 0.0
    WT = be.to_device_padded(np.ascontiguousarray(dictionary.T))
    REG = be.zeros(n_components, n_components)
    ws = be.Workspace(1, n_features, n_components)
    ld = WT.stride(0)
    be.small_gram([(WT, ld, 1, n_components, WT, ld, 1, n_components, n_features, REG, 1.0, 1)],
                  ws)
    pairs = REG.cpu().numpy()
    phi = 0.0
    for i in range(n_components):
        for j in range(i + 1, n_components):
            phi += pairs[i, j]
    return 2.0 / (n_components * n_features * (n_components - 1.0)) * phi


def _gpnh_cost(data, weights, dictionary, lambda_W=0):
    """Evaluate GPNH convex coding cost function (gpnh_convex_coding.py:199-210)."""
    torch = be.require_cuda()
    data = np.asarray(data, dtype=np.float64)
    n_samples, n_features = data.shape
    k = weights.shape[1]
    X = be.to_device_padded(data)
    Z = be.to_device(weights)
    WT = be.to_device_padded(np.ascontiguousarray(np.asarray(dictionary).T))
    out = be.zeros(1)
    part = be.zeros(n_samples)
    be.check(be.library().cdr_residual_sq(
        X.data_ptr(), X.stride(0), n_samples, n_features, Z.data_ptr(), k, WT.data_ptr(),
        WT.stride(0), out.data_ptr(), part.data_ptr(), be.stream_ptr()), 'cdr_residual_sq')
    cost = 0.5 * float(out.item()) / n_samples
    if lambda_W != 0:
        cost += lambda_W * _gpnh_regularization(dictionary)
    return cost


def _update_gpnh_dictionary(X, weights, ZtZ, GW, lambda_W=0):
    """Update dictionary for GPNH regularized convex coding
    (gpnh_convex_coding.py:213-226)."""
    X = np.asarray(X, dtype=np.float64)
    n_samples, n_features = X.shape
    k = weights.shape[1]
    # pinv(ZtZ/n + lambda GW) / n == pinv(ZtZ + n lambda GW)
    lhs = be.to_device(np.asarray(ZtZ, dtype=np.float64) +
                       n_samples * lambda_W * np.asarray(GW, dtype=np.float64))
    P = be.zeros(k, k)
    be.check(be.library().cdr_sym_pinv(lhs.data_ptr(), k, P.data_ptr(), None, be.stream_ptr()),
             'cdr_sym_pinv')
    Xd = be.to_device_padded(X)
    Z = be.to_device(weights)
    WT = be.zeros(k, Xd.stride(0))
    ws = be.Workspace(n_samples, n_features, k)
    be.reduce_samples(Z, 1, k, Xd, n_samples, n_features, k, WT, ws, E=P)
    return be.to_host(WT, k, n_features).T


def _update_gpnh_weights(X, weights, dictionary, **solver_kwargs):
    """Update weights for GPNH regularized convex coding
    (gpnh_convex_coding.py:254-279)."""
    X = np.asarray(X, dtype=np.float64)
    n_samples, n_features = X.shape
    k = weights.shape[1]
    params = be.make_spg_params(solver_kwargs)
    Xd = be.to_device_padded(X)
    Z = be.to_device(weights)
    WT = be.to_device_padded(np.ascontiguousarray(np.asarray(dictionary).T))
    ldt = be.round_up(n_samples)
    XWt = be.zeros(k, ldt)
    WtW = be.zeros(k, k)
    ws = be.Workspace(n_samples, n_features, k)
    be.reduce_features(WT, Xd, n_samples, n_features, k, XWt, ws)
    ld = WT.stride(0)
    be.small_gram([(WT, ld, 1, k, WT, ld, 1, k, n_features, WtW, 1.0, 0)], ws)
    be.quad_simplex_spg_batched(WtW, None, XWt, 1, ldt, Z, n_samples, k, params)
    return be.to_host(Z)


def _iterate_gpnh_convex_coding(X, weights, dictionary, lambda_W=0,
                                update_weights=True, update_dictionary=True,
                                tolerance=1e-6, max_iterations=1000, verbose=0,
                                **kwargs):
    """Iteratively update weights and dictionary until convergence is reached.

    Returns ``(weights, dictionary, cost, n_iter, avg_time_per_iter, cost_deltas)``
    exactly like gpnh_convex_coding.py:282-402.  ``trace_XtX`` may be passed to
    skip the initial ||X||_F^2 reduction.
    """
    if kwargs.get('dictionary_solver_kwargs', {}):
        # _update_gpnh_dictionary takes no solver options (gpnh_convex_coding.py:213, :348-350)
        raise TypeError("_update_gpnh_dictionary() got an unexpected keyword argument '%s'"
                        % next(iter(kwargs['dictionary_solver_kwargs'])))
    X = np.asarray(X, dtype=np.float64)
    be.trace('gpnh: enter _iterate')
    engine = _GpnhEngine(
        X, weights, dictionary, lambda_W=lambda_W, tolerance=tolerance,
        max_iterations=max_iterations,
        stopping_criterion=kwargs.get('stopping_criterion', 'abs_delta_f'),
        require_monotonic_cost_decrease=kwargs.get('require_monotonic_cost_decrease', True),
        weights_solver_kwargs=kwargs.get('weights_solver_kwargs', {}),
        update_dictionary=update_dictionary, update_weights=update_weights,
        trace_XtX=kwargs.get('trace_XtX'), X_device=kwargs.get('X_device'),
        comm=kwargs.get('comm'))
    engine.run(verbose=verbose)
    new_weights = engine.weights() if update_weights else weights
    new_dictionary = engine.dictionary() if update_dictionary else dictionary
    be.trace('gpnh: results to host')
    return (new_weights, new_dictionary, engine.cost, engine.n_iter,
            engine.avg_time_per_iter, engine.cost_deltas)


def _initialize_gpnh_convex_coding_dictionary(data, n_components, init='random',
                                              random_state=None, **kwargs):
    """gpnh_convex_coding.py:41-81, 93-115."""
    if init is None:
        init = 'random'
    rng = check_random_state(random_state)
    n_samples, n_features = data.shape
    comm = kwargs.get('comm', None)
    sharded = comm is not None and comm.enabled     # data = this rank's rows
    if sharded:
        n_samples = comm.local_rows(n_samples)[1]
    if init == 'random':
        if sharded:
            mean_abs = comm.sum_scalars([np.abs(data).sum()])[0] / (n_samples * n_features)
        else:
            mean_abs = np.abs(data).mean()
        avg = np.sqrt(mean_abs / n_components)
        return avg * rng.randn(n_features, n_components)
    if init == 'furthest_sum':
        start_index = kwargs.get('start_index', None)
        n_extra_steps = kwargs.get('n_extra_steps', 10)
        exclude = kwargs.get('exclude', None)
        if start_index is None:
            start_index = rng.randint(n_samples)
        if exclude is None:
            exclude = np.array([], dtype='i8')
        if sharded:
            from .kmeans import furthest_sum_centres, gather_rows
            selected = furthest_sum_centres(data, n_components, start_index, n_extra_steps,
                                            exclude, comm=comm)
            return np.ascontiguousarray(gather_rows(data, selected, comm).T)
        Xd = be.to_device_padded(data)
        K = be.gram(Xd, n_samples, n_features)
        D = dissimilarity_from_gram_device(K, n_samples)
        selected = furthest_sum_device(D, n_samples, n_components, start_index, exclude,
                                       n_extra_steps)
        dictionary = np.zeros((n_features, n_components), dtype=np.float64)
        for i in range(n_components):
            dictionary[:, i] = data[selected[i]]
        return dictionary
    raise ValueError('Invalid init parameter: got %r instead of one of %r' %
                     (init, INITIALIZATION_METHODS))


def _initialize_gpnh_convex_coding_weights(data, n_components, init='random',
                                           random_state=None, n_samples=None):
    """gpnh_convex_coding.py:84-90, 118-129 (``n_samples``: total rows of a sharded fit)."""
    if init is None:
        init = 'random'
    if init in ('furthest_sum', 'random'):
        rng = check_random_state(random_state)
        if n_samples is None:
            n_samples = data.shape[0]
        return right_stochastic_matrix((n_samples, n_components), random_state=rng)
    raise ValueError('Invalid init parameter: got %r instead of one of %r' %
                     (init, INITIALIZATION_METHODS))


def _initialize_gpnh_convex_coding(data, n_components, init='random',
                                   random_state=None, **kwargs):
    """Dictionary first, then weights: the RNG draw order of gpnh_convex_coding.py:132-143."""
    rng = check_random_state(random_state)
    dictionary = _initialize_gpnh_convex_coding_dictionary(
        data, n_components, init=init, random_state=rng, **kwargs)
    weights = _initialize_gpnh_convex_coding_weights(
        data, n_components, init=init, random_state=rng, n_samples=kwargs.get('n_samples'))
    return dictionary, weights


class GPNHConvexCoding():
    """Convex encoding of data with GPNH regularization.

    Drop-in for the reference class (gpnh_convex_coding.py:405-668): same
    constructor, ``fit`` / ``fit_transform`` / ``transform`` /
    ``inverse_transform`` and the attributes ``weights``, ``dictionary``,
    ``cost``, ``n_iter``, ``avg_time_per_iter``, ``cost_deltas``.
    """

    def __init__(self, n_components, lambda_W=0, init=None, tolerance=1e-6,
                 max_iterations=1000, verbose=0, random_state=None, **kwargs):
        self.n_components = n_components
        self.lambda_W = lambda_W
        self.init = init
        self.tolerance = tolerance
        self.max_iterations = max_iterations
        self.verbose = verbose
        self.random_state = check_random_state(random_state)
        self.require_monotonic_cost_decrease = kwargs.get(
            'require_monotonic_cost_decrease', True)
        self.stopping_criterion = kwargs.get('stopping_criterion', 'abs_delta_f')
        self.weights = None
        self.dictionary = None
        self.cost = 0
        self.n_iter = 0
        self.avg_time_per_iter = 0
        self.cost_deltas = None
        self.weights_solver_kwargs = kwargs.get('weights_solver_kwargs', {})
        self.dictionary_solver_kwargs = kwargs.get('dictionary_solver_kwargs', {})

    def _check_params(self, n_features):
        if self.n_components is None:
            self.n_components = n_features
        if not isinstance(self.n_components, INTEGER_TYPES) or self.n_components <= 0:
            raise ValueError('Number of components must be a positive integer;'
                             ' got (n_components=%r)' % self.n_components)
        if not isinstance(self.max_iterations, INTEGER_TYPES) or self.max_iterations <= 0:
            raise ValueError('Maximum number of iterations must be a positive '
                             'integer; got (max_iterations=%r)' % self.max_iterations)
        if not isinstance(self.tolerance, numbers.Number) or self.tolerance < 0:
            raise ValueError('Tolerance for stopping criteria must be '
                             'positive; got (tolerance=%r)' % self.tolerance)

    def _gpnh_convex_coding(self, data, dictionary=None, weights=None,
                            update_dictionary=True, update_weights=True, **kwargs):
        """Calculate GPNH-regularized convex coding of dataset
        (gpnh_convex_coding.py:501-572).

        ``comm=Comm()`` runs the sample-sharded fit: ``data`` is this rank's balanced row
        block, initial factors passed in are the full-size ones, the random draws are the
        single-process ones (same seed on every rank) and ``self.weights`` ends up holding
        the weights of all rows on every rank."""
        n_samples, n_features = data.shape
        comm = kwargs.get('comm', None)
        sharded = comm is not None and comm.enabled
        row0 = 0
        if sharded:
            row0, n_samples = comm.local_rows(n_samples)
            comm.sync_random_state(self.random_state)      # rank 0 is authoritative
            kwargs = dict(kwargs, n_samples=n_samples)
        self._check_params(n_features)
        k = self.n_components
        if self.init == 'custom':
            _check_init_weights(weights, (n_samples, k), '_gpnh_convex_coding (input weights)')
            _check_init_dictionary(dictionary, (n_features, k),
                                   '_gpnh_convex_coding (input dictionary)')
        elif not update_dictionary and update_weights:
            _check_init_dictionary(dictionary, (n_features, k),
                                   '_gpnh_convex_coding (input dictionary)')
            weights = _initialize_gpnh_convex_coding_weights(
                data, k, init=self.init, random_state=self.random_state, n_samples=n_samples)
        elif update_dictionary and not update_weights:
            _check_init_weights(weights, (n_samples, k), '_gpnh_convex_coding (input weights)')
            dictionary = _initialize_gpnh_convex_coding_dictionary(
                data, k, init=self.init, random_state=self.random_state, **kwargs)
        else:
            dictionary, weights = _initialize_gpnh_convex_coding(
                data, k, init=self.init, random_state=self.random_state, **kwargs)

        self.weights = np.array(weights, dtype=np.float64)
        self.dictionary = np.array(dictionary, dtype=np.float64)
        extra = {}
        if sharded:
            self.weights = np.ascontiguousarray(self.weights[row0:row0 + data.shape[0]])
            extra['comm'] = comm

        self.weights, self.dictionary, cost, n_iter, avg_time_per_iter, cost_deltas = \
            _iterate_gpnh_convex_coding(
                data, self.weights, self.dictionary, lambda_W=self.lambda_W,
                update_dictionary=update_dictionary, update_weights=update_weights,
                tolerance=self.tolerance, max_iterations=self.max_iterations,
                verbose=self.verbose,
                require_monotonic_cost_decrease=self.require_monotonic_cost_decrease,
                stopping_criterion=self.stopping_criterion,
                weights_solver_kwargs=self.weights_solver_kwargs,
                dictionary_solver_kwargs=self.dictionary_solver_kwargs, **extra)
        if sharded:
            self.weights = comm.allgather_rows(self.weights)

        if n_iter == self.max_iterations and self.tolerance > 0:
            warnings.warn('Maximum number of iterations %d reached.' %
                          self.max_iterations, UserWarning)
        return cost, n_iter, avg_time_per_iter, cost_deltas

    def fit_transform(self, data, dictionary=None, weights=None, **kwargs):
        """Fit convex coding and return transformed data
        (gpnh_convex_coding.py:574-603)."""
        data = np.asarray(data)
        cost_, n_iter_, avg_time_per_iter_, cost_deltas_ = self._gpnh_convex_coding(
            data, dictionary=dictionary, weights=weights, **kwargs)
        self.cost = cost_
        self.n_iter = n_iter_
        self.avg_time_per_iter = avg_time_per_iter_
        self.cost_deltas = cost_deltas_
        return self.weights

    def fit(self, data, **kwargs):
        """Fit convex coding to data."""
        self.fit_transform(data, **kwargs)
        return self

    def transform(self, data):
        """Transform the data according to the fitted factorization
        (gpnh_convex_coding.py:623-652): weights-only run with a fresh random start."""
        data = np.asarray(data)
        cost_ = self._gpnh_convex_coding(
            data=data, dictionary=self.dictionary,
            update_dictionary=False, update_weights=True)[0]
        return self.weights, cost_

    def inverse_transform(self, weights):
        """Transform data back into its original space."""
        return weights.dot(self.dictionary.T)
