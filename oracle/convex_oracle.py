"""CPU oracle for the convex_dim_red hot path -- TEST INFRASTRUCTURE ONLY.

This module is a NumPy (+ small C library, ``oracle/cdr_oracle.c``) restatement
of the alternating-update path of azedarach/matrix-factorization-case-studies.
It exists so the CUDA path can be checked against the reference's algorithm on
a GPU box where ``/root/reference`` is not present.

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs.  The product package
(``matrix-factorization-case-studies_b200/convex_dim_red``) never imports it
and has no CPU fallback.

Parity status: **pinned**.  ``tests/test_oracle_golden.py`` compares every
function below with fixtures produced by importing the real reference in the
build container (``tests/golden/make_golden.py``) and with the known-answer
vectors of the reference's own tests.  k-means restates scikit-learn 1.9.0
(third-party, not vendored by the reference; call sites
``bin/run_hadisst_kmeans.py:128-131``) and is pinned against fixtures generated
by that sklearn version with an explicit ``init`` array.

All citations ``file:line`` are relative to ``/root/reference/src/convex_dim_red``.
"""

import ctypes
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class SpgParams(ctypes.Structure):
    """Mirror of ``orc_spg_params`` (defaults: spg.py:287-291)."""

    _fields_ = [
        ('gamma', ctypes.c_double), ('memory', ctypes.c_int),
        ('sigma_one', ctypes.c_double), ('sigma_two', ctypes.c_double),
        ('lambda_min', ctypes.c_double), ('alpha0', ctypes.c_double),
        ('alpha_min', ctypes.c_double), ('alpha_max', ctypes.c_double),
        ('epsilon_one', ctypes.c_double), ('epsilon_two', ctypes.c_double),
        ('max_iterations', ctypes.c_int), ('max_feval', ctypes.c_int)]


QP_DEFAULTS = dict(gamma=1e-4, memory=1, sigma_one=0.1, sigma_two=0.9,
                   lambda_min=1e-10, alpha0=-1.0, alpha_min=1e-5,
                   alpha_max=1e3, epsilon_one=1e-10, epsilon_two=1e-6,
                   max_iterations=1000, max_feval=2000)


def qp_params(**kw):
    vals = dict(QP_DEFAULTS)
    vals.update({k: v for k, v in kw.items() if k in vals})
    return SpgParams(**vals)


def build_clib(force=False):
    """Compile oracle/cdr_oracle.c (building the checker is not using it)."""
    so = os.path.join(_HERE, 'libcdr_oracle.so')
    src = os.path.join(_HERE, 'cdr_oracle.c')
    if force or not os.path.exists(so) or \
            os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return so


def clib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_clib())
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ---------------------------------------------------------------------------
# simplex projection (simplex_projection.py)
# ---------------------------------------------------------------------------

def simplex_project_vector_py(x):
    """Pure-Python restatement of simplex_projection.py:13-27 (small cases)."""
    x = np.asarray(x, dtype=np.float64)
    srt = np.sort(x)
    n = srt.size
    t_hat = 0.0
    for i in range(n - 2, -2, -1):
        m = n - 1 - i
        s = 0.0
        for v in srt[n - m:]:
            s += v
        t_hat = (s - 1) / m
        if t_hat >= srt[i]:
            break
    return np.fmax(x - t_hat, 0)


def simplex_project_vector(x):
    x = _c64(x)
    out = np.empty_like(x)
    work = np.empty_like(x)
    clib().orc_simplex_project_vector(_dp(x), _dp(out), ctypes.c_int(x.size), _dp(work))
    return out


def simplex_project_rows(a):
    """simplex_projection.py:40-47."""
    a = _c64(a)
    out = np.empty_like(a)
    clib().orc_simplex_project_rows(_dp(a), _dp(out), ctypes.c_int(a.shape[0]),
                                    ctypes.c_int(a.shape[1]))
    return out


def simplex_project_columns(a):
    """simplex_projection.py:30-37."""
    a = _c64(a)
    out = np.empty_like(a)
    clib().orc_simplex_project_columns(_dp(a), _dp(out), ctypes.c_int(a.shape[0]),
                                       ctypes.c_int(a.shape[1]))
    return out


# ---------------------------------------------------------------------------
# SPG (spg.py)
# ---------------------------------------------------------------------------

def line_search_step_length(lam, delta, f_old, f_new, sigma_one=0.1, sigma_two=0.9):
    """spg.py:19-33."""
    cand = -0.5 * lam ** 2 * delta / (f_new - f_old - lam * delta)
    if sigma_one <= cand <= sigma_two * lam:
        return cand
    return 0.5 * lam


def cauchy_step_size(beta, sksk, alpha_min=1e-3, alpha_max=1e3):
    """spg.py:36-43."""
    if beta <= 0:
        return alpha_max
    return min(alpha_max, max(alpha_min, sksk / beta))


def spg(f, df, x0, project=None, gamma=1e-4, memory=1, sigma_one=0.1,
        sigma_two=0.9, lambda_min=1e-10, alpha0=None, alpha_min=1e-5,
        alpha_max=1e3, epsilon_one=1e-10, epsilon_two=1e-6,
        use_infinity_norm=True, max_iterations=10000, max_feval=1000000,
        trace=None):
    """Generic non-monotone SPG, spg.py:46-283 (warnings omitted).

    ``trace`` (optional list) receives one dict per iteration with the scalar
    decisions (lam, alpha, f, n_feval), used by the parity tests.
    """
    multi = not np.isscalar(x0)
    x = x0.copy() if multi else x0
    if project is not None:
        x = project(x)
    alpha = alpha0
    f_mem = np.zeros(memory)                      # zeros, not NaN: spg.py:153
    f_old = f(x)
    n_feval = 1
    n_iter = -1
    converged = False
    for n_iter in range(max_iterations):
        x_old = x.copy() if multi else x
        gk = df(x)
        if alpha is None:
            if project is None:
                alpha = 1.0 / np.max(np.abs(gk))
            else:
                a_inv = np.max(np.abs(project(x - gk) - x))
                alpha = 1.0 / a_inv if abs(a_inv) > 1e-12 else 1.0
        dk = -alpha * gk
        if project is not None:
            dk = project(x + dk)
            dk -= x
        f_mem = np.roll(f_mem, 1)
        f_mem[0] = f_old
        f_max = None
        for prev in f_mem:
            if f_max is None or prev >= f_max:
                f_max = prev
        delta = np.sum(dk * gk)
        lam = 1
        x = x_old + dk
        f_new = f(x)
        n_feval += 1
        while f_new > f_max + gamma * lam * delta:
            lam = line_search_step_length(lam, delta, f_old, f_new,
                                          sigma_one, sigma_two)
            x = x_old + lam * dk
            f_new = f(x)
            n_feval += 1
            if abs(lam) < lambda_min:
                break
        yk = gk.copy() if multi else gk
        gk = df(x)
        yk = gk - yk
        sksk = lam ** 2 * np.sum(dk * dk)
        betak = lam * np.sum(dk * yk)
        alpha_used = alpha
        alpha = cauchy_step_size(betak, sksk, alpha_min, alpha_max)
        f_old = f(x)
        n_feval += 1
        res = -gk if project is None else project(x - gk) - x
        res_norm = np.sum(res ** 2) ** 0.5
        converged = res_norm < epsilon_two
        if use_infinity_norm:
            converged = converged or np.max(np.abs(res)) < epsilon_one
        if trace is not None:
            trace.append(dict(lam=float(lam), alpha=float(alpha_used),
                              alpha_next=float(alpha), f=float(f_old),
                              n_feval=n_feval, delta=float(delta),
                              res_norm=float(res_norm)))
        if converged:
            break
        if n_feval > max_feval:
            break
    return x, f_old, n_iter, n_feval


def quad_simplex_spg_py(A, b, x0, **kw):
    """Pure-Python restatement of spg.py:286-398 (small cases only)."""
    p = dict(QP_DEFAULTS)
    p.update(kw)
    x = simplex_project_vector_py(x0)
    f_mem = np.full(p['memory'], np.nan)
    Ax = A.dot(x)
    f_old = 0.5 * x.dot(Ax) + x.dot(b)
    n_feval = 1
    alpha = None
    for n_iter in range(p['max_iterations']):
        x_old = x.copy()
        gk = Ax + b
        if n_iter == 0:
            if p['alpha_min'] <= p['alpha0'] <= p['alpha_max']:
                alpha = p['alpha0']
            else:
                a_inv = np.max(np.abs(simplex_project_vector_py(x - gk) - x))
                if abs(a_inv) < 1e-12:
                    a_inv = 1.0
                alpha = min(max(p['alpha_min'], 1.0 / a_inv), p['alpha_max'])
        dk = simplex_project_vector_py(x - alpha * gk) - x
        f_mem = np.roll(f_mem, 1)
        f_mem[0] = f_old
        f_max = np.nanmax(f_mem)
        delta = dk.dot(gk)
        lam = 1
        x = x_old + dk
        Ax = A.dot(x)
        f_new = 0.5 * x.dot(Ax) + x.dot(b)
        n_feval += 1
        while f_new > f_max + p['gamma'] * lam * delta:
            lam = line_search_step_length(lam, delta, f_old, f_new,
                                          p['sigma_one'], p['sigma_two'])
            x = x_old + lam * dk
            Ax = A.dot(x)
            f_new = 0.5 * x.dot(Ax) + x.dot(b)
            n_feval += 1
            if abs(lam) < p['lambda_min']:
                break
        yk = Ax + b - gk
        gk = yk + gk
        sksk = lam ** 2 * dk.dot(dk)
        betak = lam * dk.dot(yk)
        alpha = cauchy_step_size(betak, sksk, p['alpha_min'], p['alpha_max'])
        f_old = 0.5 * x.dot(Ax) + x.dot(b)
        n_feval += 1
        res = simplex_project_vector_py(x - gk) - x
        if np.sum(res ** 2) ** 0.5 < p['epsilon_two'] or \
                np.max(np.abs(res)) < p['epsilon_one']:
            break
        if n_feval > p['max_feval']:
            break
    return x


def quad_simplex_spg(A, b, x0, **kw):
    """C-backed spg.py:286-398."""
    A = _c64(A)
    b = _c64(b)
    x0 = _c64(x0)
    k = x0.size
    p = qp_params(**kw)
    out = np.empty(k)
    ws = np.empty(10 * k + p.memory)
    clib().orc_quad_simplex_spg(_dp(A), _dp(b), _dp(x0), _dp(out), ctypes.c_int(k),
                                ctypes.byref(p), None, None, _dp(ws))
    return out


def weights_update(A, B, Z0, b_is_k_by_T, return_counts=False, **kw):
    """One QP per sample with linear term -B[:, t] (AA) or -B[t] (GPNH).

    archetypal_analysis.py:344-366, gpnh_convex_coding.py:229-251.
    """
    A = _c64(A)
    B = _c64(B)
    Z0 = _c64(Z0)
    T, k = Z0.shape
    p = qp_params(**kw)
    Z = np.empty_like(Z0)
    n_iter = np.zeros(T, dtype=np.int32)
    n_feval = np.zeros(T, dtype=np.int32)
    if b_is_k_by_T:
        sb_t, sb_c = 1, B.shape[1]
    else:
        sb_t, sb_c = B.shape[1], 1
    clib().orc_weights_update(
        _dp(A), _dp(B), ctypes.c_long(sb_t), ctypes.c_long(sb_c), _dp(Z0), _dp(Z),
        ctypes.c_int(T), ctypes.c_int(k), ctypes.byref(p),
        n_iter.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
        n_feval.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    if return_counts:
        return Z, n_iter, n_feval
    return Z


# ---------------------------------------------------------------------------
# stochastic matrices (stochastic_matrices.py:15-39)
# ---------------------------------------------------------------------------

def _rng(random_state):
    if random_state is None:
        return np.random.mtrand._rand
    if isinstance(random_state, (int, np.integer)):
        return np.random.RandomState(random_state)
    return random_state


def right_stochastic_matrix(shape, random_state=None):
    m = _rng(random_state).uniform(size=shape)
    return m / np.sum(m, axis=1)[:, np.newaxis]


def left_stochastic_matrix(shape, random_state=None):
    m = _rng(random_state).uniform(size=shape)
    return m / np.sum(m, axis=0)[np.newaxis, :]


# ---------------------------------------------------------------------------
# furthest sum (furthest_sum.py:23-127)
# ---------------------------------------------------------------------------

def furthest_sum(D, n_components, start_index, exclude=None, extra_steps=1):
    """Greedy furthest-sum selection; list order and stable sort preserved."""
    if D.shape[0] != D.shape[1]:
        raise ValueError('Dissimilarity matrix must be square')
    if n_components == 0:
        return []
    exclude = [] if exclude is None else list(exclude)
    n = D.shape[0]
    if start_index >= n:
        raise ValueError('Start index out of bounds')
    if start_index in exclude:
        raise ValueError('Start index is excluded')
    if len(exclude) < n and n_components > n - len(exclude):
        raise ValueError('Too few points available')
    selected = np.full((n_components,), start_index)
    banned = set(exclude) | {start_index}
    queue = [[i, D[i, start_index]] for i in range(n) if i not in banned]

    def take_furthest():
        queue.sort(key=lambda item: item[1])       # stable, persistent order
        return queue.pop(-1)[0]

    def add_row(idx):
        for item in queue:
            item[1] += D[idx, item[0]]

    for i in range(1, n_components):
        selected[i] = take_furthest()
        add_row(selected[i])
    for step in range(max(extra_steps, 0)):
        slot = step % n_components
        old = selected[slot]
        for item in queue:
            item[1] -= D[item[0], old]
        acc = 0
        for idx in selected:
            if idx != old:
                acc += D[old, idx]
        queue.append([old, acc])
        selected[slot] = take_furthest()
        add_row(selected[slot])
    return selected


def dissimilarity_from_kernel(K):
    """archetypal_analysis.py:96-100 / gpnh_convex_coding.py:67-71."""
    n = K.shape[0]
    kd = np.diag(K)
    return np.sqrt(np.tile(kd, (n, 1)) - 2 * K + np.tile(kd[:, np.newaxis], (1, n)))


# ---------------------------------------------------------------------------
# archetypal analysis (archetypal_analysis.py)
# ---------------------------------------------------------------------------

def cost_increased(old, new, tolerance):
    """archetypal_analysis.py:167-174 (predicate only)."""
    return (new > old) and (abs(new - old) > tolerance)


def make_stopping_rule(name):
    """archetypal_analysis.py:177-197."""
    if name == 'abs_delta_f':
        return lambda old, new, tol: abs(new - old) < tol
    if name == 'rel_delta_f':
        return lambda old, new, tol: abs((new - old) / max(abs(new), abs(old))) < tol
    raise ValueError("unsupported stopping criterion '%s'" % name)


def kernel_aa_cost(K, Z, C, alpha):
    """archetypal_analysis.py:200-217."""
    da = np.diag(alpha)
    CK = C.dot(K)
    CKCt = CK.dot(C.T)
    CKZ = CK.dot(Z)
    ZtZ = Z.T.dot(Z)
    return 0.5 * (np.trace(K) - 2 * np.trace(da.dot(CKZ)) +
                  np.trace((da.dot(ZtZ.dot(da))).dot(CKCt))) / K.shape[0]


def scale_factors_objective(alpha, trace_K, CKZ, ZtZ, CKCt):
    """archetypal_analysis.py:220-229."""
    a2 = np.outer(alpha, alpha)
    return 0.5 * (trace_K - 2 * alpha.dot(np.diag(CKZ)) +
                  np.sum(a2 * ZtZ * CKCt)) / CKZ.shape[1]


def scale_factors_gradient(alpha, CKZ, ZtZ, CKCt):
    """archetypal_analysis.py:232-240."""
    return np.diag(ZtZ.dot(np.diag(alpha).dot(CKCt)) - CKZ) / CKZ.shape[1]


def update_scale_factors(alpha, trace_K, CKZ, ZtZ, CKCt, delta, **kw):
    """archetypal_analysis.py:243-258."""
    return spg(lambda a: scale_factors_objective(a, trace_K, CKZ, ZtZ, CKCt),
               lambda a: scale_factors_gradient(a, CKZ, ZtZ, CKCt), alpha,
               project=lambda a: np.fmin(np.fmax(1.0 - delta, a), 1.0 + delta),
               **kw)[0]


def aa_dictionary_cost(X, C, trace_XXt, XXtZD, DZtZD):
    """archetypal_analysis.py:261-270 -- note the division by k (quirk 3.5.1)."""
    CX = C.dot(X)
    return 0.5 * (trace_XXt - 2 * np.trace(C.dot(XXtZD)) +
                  np.trace(DZtZD.dot(CX.dot(CX.T)))) / C.shape[0]


def aa_dictionary_gradient(X, C, XXtZD, DZtZD):
    """archetypal_analysis.py:293-301 -- divided by T."""
    return (DZtZD.dot(C.dot(X).dot(X.T)) - XXtZD.T) / C.shape[1]


def kernel_aa_dictionary_cost(K, C, trace_K, KZD, DZtZD):
    """archetypal_analysis.py:273-281 -- divided by k."""
    return 0.5 * (trace_K - 2 * np.trace(C.dot(KZD)) +
                  np.trace(DZtZD.dot(C.dot(K.dot(C.T))))) / C.shape[0]


def kernel_aa_dictionary_gradient(K, C, KZD, DZtZD):
    """archetypal_analysis.py:284-290 -- divided by k."""
    return (DZtZD.dot(C.dot(K)) - KZD.T) / C.shape[0]


def update_kernel_aa_dictionary(K, C, alpha, trace_K, KZ, ZtZ, trace=None, **kw):
    """archetypal_analysis.py:304-321."""
    da = np.diag(alpha)
    KZD = KZ.dot(da)
    DZtZD = da.dot(ZtZ.dot(da))
    return spg(lambda c: kernel_aa_dictionary_cost(K, c, trace_K, KZD, DZtZD),
               lambda c: kernel_aa_dictionary_gradient(K, c, KZD, DZtZD), C,
               project=simplex_project_rows, trace=trace, **kw)[0]


def update_aa_dictionary(X, C, alpha, trace_XXt, XXtZ, ZtZ, trace=None, **kw):
    """archetypal_analysis.py:324-341."""
    da = np.diag(alpha)
    XXtZD = XXtZ.dot(da)
    DZtZD = da.dot(ZtZ.dot(da))
    return spg(lambda c: aa_dictionary_cost(X, c, trace_XXt, XXtZD, DZtZD),
               lambda c: aa_dictionary_gradient(X, c, XXtZD, DZtZD), C,
               project=simplex_project_rows, trace=trace, **kw)[0]


def update_kernel_aa_weights(Z, alpha, CK, CKCt, **kw):
    """archetypal_analysis.py:369-396."""
    da = np.diag(alpha)
    return weights_update(da.dot(CKCt.dot(da)), da.dot(CK), Z, True, **kw)


def _aa_loop(apply_left, apply_right, trace_data, Z, C, alpha, delta,
             update_dict, n_samples, update_weights, update_dictionary,
             update_scale, tolerance, max_iterations, kwargs):
    """Shared body of archetypal_analysis.py:399-531 and :534-670.

    apply_left(C) -> (state, C K-like k x T, C K C' k x k); apply_right(Z) ->
    K Z-like T x k.  The two reference loops differ only in these products and
    in the dictionary cost/gradient scaling handled by ``update_dict``.
    """
    da = np.diag(alpha)
    ZtZ = Z.T.dot(Z)
    CK, CKCt = apply_left(C)
    KZ = apply_right(Z)
    CKZ = C.dot(KZ)

    def cost_now():
        return 0.5 * (trace_data - 2 * da.dot(CKZ).trace() +
                      (da.dot(ZtZ.dot(da))).dot(CKCt).trace()) / n_samples

    new_cost = cost_now()
    require = kwargs.get('require_monotonic_cost_decrease', True)
    stop = make_stopping_rule(kwargs.get('stopping_criterion', 'abs_delta_f'))
    d_kw = kwargs.get('dictionary_solver_kwargs', {})
    w_kw = kwargs.get('weights_solver_kwargs', {})
    s_kw = kwargs.get('scale_factors_solver_kwargs', {})
    iter_times, cost_deltas = [], []
    n_iter = -1
    for n_iter in range(max_iterations):
        t0 = time.perf_counter()
        old_cost = new_cost
        if update_scale and delta != 0:
            alpha = update_scale_factors(alpha, trace_data, CKZ, ZtZ, CKCt, delta, **s_kw)
            da = np.diag(alpha)
            new_cost = cost_now()
            if require and cost_increased(old_cost, new_cost, tolerance):
                raise RuntimeError('factorization cost increased after scale factors update')
        if update_dictionary:
            C = update_dict(C, alpha, KZ, ZtZ, d_kw)
            CK, CKCt = apply_left(C)
            CKZ = C.dot(KZ)
            new_cost = cost_now()
            if require and cost_increased(old_cost, new_cost, tolerance):
                raise RuntimeError('factorization cost increased after dictionary update')
        if update_weights:
            Z = update_kernel_aa_weights(Z, alpha, CK, CKCt, **w_kw)
            ZtZ = Z.T.dot(Z)
            KZ = apply_right(Z)
            CKZ = C.dot(KZ)
            new_cost = cost_now()
            if require and cost_increased(old_cost, new_cost, tolerance):
                raise RuntimeError('factorization cost increased after weights update')
        iter_times.append(time.perf_counter() - t0)
        cost_deltas.append(new_cost - old_cost)
        if stop(old_cost, new_cost, tolerance):
            break
    if kwargs.get('iter_times_out') is not None:
        kwargs['iter_times_out'].extend(iter_times)      # bench.py's CPU arm
    return Z, C, alpha, new_cost, n_iter, float(np.mean(iter_times)), cost_deltas


def iterate_kernel_aa(K, Z, C, alpha, delta=0, update_weights=True,
                      update_dictionary=True, update_scale_factors=True,
                      tolerance=1e-6, max_iterations=1000, **kwargs):
    """archetypal_analysis.py:399-531."""
    def left(c):
        ck = c.dot(K)
        return ck, ck.dot(c.T)

    trace_K = K.trace()
    return _aa_loop(
        left, lambda z: K.dot(z), trace_K, Z, C, alpha, delta,
        lambda c, a, kz, ztz, kw: update_kernel_aa_dictionary(K, c, a, trace_K, kz, ztz, **kw),
        Z.shape[0], update_weights, update_dictionary, update_scale_factors,
        tolerance, max_iterations, kwargs)


def iterate_aa(X, Z, C, alpha, delta=0, update_weights=True,
               update_dictionary=True, update_scale_factors=True,
               tolerance=1e-6, max_iterations=1000, trace_XXt=None, **kwargs):
    """archetypal_analysis.py:534-670.

    ``trace_XXt`` may be supplied to skip the T x T product at :552 (it is only
    used for its trace, which equals ||X||_F^2).
    """
    def left(c):
        cx = c.dot(X)
        return cx.dot(X.T), cx.dot(cx.T)

    if trace_XXt is None:
        trace_XXt = np.trace(X.dot(X.T))
    return _aa_loop(
        left, lambda z: X.dot(X.T.dot(z)), trace_XXt, Z, C, alpha, delta,
        lambda c, a, xxtz, ztz, kw: update_aa_dictionary(X, c, a, trace_XXt, xxtz, ztz, **kw),
        Z.shape[0], update_weights, update_dictionary, update_scale_factors,
        tolerance, max_iterations, kwargs)


def init_kernel_aa(K, k, init, rng, **kwargs):
    """archetypal_analysis.py:51-164: dictionary first, then weights."""
    T = K.shape[0]
    if init is None:
        init = 'furthest_sum'
    if init == 'random':
        C = right_stochastic_matrix((k, T), rng)
    elif init == 'furthest_sum':
        start = kwargs.get('start_index', None)
        if start is None:
            start = rng.randint(T)
        excl = kwargs.get('exclude', None)
        if excl is None:
            excl = np.array([], dtype='i8')
        picks = furthest_sum(dissimilarity_from_kernel(K), k, start, excl,
                             kwargs.get('n_extra_steps', 10))
        C = np.zeros((k, T), dtype=K.dtype)
        for i in range(k):
            C[i, picks[i]] = 1
    else:
        raise ValueError('Invalid init parameter')
    Z = right_stochastic_matrix((T, k), rng)
    return C, Z


def aa_transform(archetypes, data, Z0, max_iterations, **w_kw):
    """ArchetypalAnalysis.transform, archetypal_analysis.py:1151-1199."""
    kw = {k: v for k, v in w_kw.items() if k != 'max_iterations'}
    Z = weights_update(archetypes.dot(archetypes.T), archetypes.dot(data.T), Z0,
                       True, max_iterations=max_iterations, **kw)
    cost = 0.5 * np.linalg.norm(data - Z.dot(archetypes)) ** 2 / data.shape[0]
    return Z, cost


# ---------------------------------------------------------------------------
# GPNH convex coding (gpnh_convex_coding.py)
# ---------------------------------------------------------------------------

def gpnh_regularization(W):
    """gpnh_convex_coding.py:179-196."""
    d, k = W.shape
    if k == 1:
        return 0.0
    acc = 0.0
    for i in range(k):
        for j in range(i + 1, k):
            acc += np.linalg.norm(W[:, i] - W[:, j]) ** 2
    return 2.0 / (k * d * (k - 1.0)) * acc


def gpnh_cost(X, Z, W, lambda_W=0):
    """gpnh_convex_coding.py:199-210."""
    cost = 0.5 * np.linalg.norm(X - Z.dot(W.T)) ** 2 / X.shape[0]
    if lambda_W != 0:
        cost += lambda_W * gpnh_regularization(W)
    return cost


def gpnh_GW(d, k):
    """gpnh_convex_coding.py:296-300."""
    if k > 1:
        return 4.0 / (d * k * (k - 1)) * (k * np.eye(k) - 1)
    return np.zeros((k, k))


def update_gpnh_dictionary(X, Z, ZtZ, GW, lambda_W=0):
    """gpnh_convex_coding.py:213-226 (min-norm least squares, rcond=None)."""
    n = X.shape[0]
    return np.linalg.lstsq(ZtZ / n + lambda_W * GW, Z.T.dot(X) / n, rcond=None)[0].T


def update_gpnh_weights(X, Z, W, **kw):
    """gpnh_convex_coding.py:254-279."""
    return weights_update(W.T.dot(W), X.dot(W), Z, False, **kw)


def iterate_gpnh(X, Z, W, lambda_W=0, update_weights=True, update_dictionary=True,
                 tolerance=1e-6, max_iterations=1000, trace_XtX=None, **kwargs):
    """gpnh_convex_coding.py:282-402.

    ``trace_XtX`` may be supplied (= ||X||_F^2) to skip the d x d product at
    :302, which the reference itself cannot form at d = 44 000.
    """
    d = X.shape[1]
    T, k = Z.shape
    WtXt = W.T.dot(X.T)
    ZtZ = Z.T.dot(Z)
    WtW = W.T.dot(W)
    GW = gpnh_GW(d, k)
    if trace_XtX is None:
        trace_XtX = X.T.dot(X).trace()
    penalty = lambda_W * gpnh_regularization(W) if lambda_W != 0 else 0

    def cost_now():
        return 0.5 * (trace_XtX - 2 * WtXt.dot(Z).trace() +
                      ZtZ.dot(WtW).trace()) / T + penalty

    new_cost = cost_now()
    require = kwargs.get('require_monotonic_cost_decrease', True)
    stop = make_stopping_rule(kwargs.get('stopping_criterion', 'abs_delta_f'))
    w_kw = kwargs.get('weights_solver_kwargs', {})
    if kwargs.get('dictionary_solver_kwargs', {}):
        raise TypeError('_update_gpnh_dictionary() got an unexpected keyword argument')
    iter_times, cost_deltas = [], []
    n_iter = -1
    for n_iter in range(max_iterations):
        t0 = time.perf_counter()
        old_cost = new_cost
        if update_dictionary:
            W = update_gpnh_dictionary(X, Z, ZtZ, GW, lambda_W=lambda_W)
            WtXt = W.T.dot(X.T)
            WtW = W.T.dot(W)
            penalty = lambda_W * gpnh_regularization(W) if lambda_W != 0 else 0
            new_cost = cost_now()
            if require and cost_increased(old_cost, new_cost, tolerance):
                raise RuntimeError('factorization cost increased after dictionary update')
        if update_weights:
            Z = update_gpnh_weights(X, Z, W, **w_kw)
            ZtZ = Z.T.dot(Z)
            new_cost = cost_now()
            if require and cost_increased(old_cost, new_cost, tolerance):
                raise RuntimeError('factorization cost increased after weights update')
        iter_times.append(time.perf_counter() - t0)
        cost_deltas.append(new_cost - old_cost)
        if stop(old_cost, new_cost, tolerance):
            break
    if kwargs.get('iter_times_out') is not None:
        kwargs['iter_times_out'].extend(iter_times)      # bench.py's CPU arm
    return Z, W, new_cost, n_iter, float(np.mean(iter_times)), cost_deltas


def init_gpnh(X, k, init, rng, **kwargs):
    """gpnh_convex_coding.py:41-143: dictionary first, then weights."""
    T, d = X.shape
    if init is None:
        init = 'random'
    if init == 'random':
        W = np.sqrt(np.abs(X).mean() / k) * rng.randn(d, k)
    elif init == 'furthest_sum':
        K = X.dot(X.T)
        start = kwargs.get('start_index', None)
        if start is None:
            start = rng.randint(T)
        excl = kwargs.get('exclude', None)
        if excl is None:
            excl = np.array([], dtype='i8')
        picks = furthest_sum(dissimilarity_from_kernel(K), k, start, excl,
                             kwargs.get('n_extra_steps', 10))
        W = np.zeros((d, k), dtype=K.dtype)
        for i in range(k):
            W[:, i] = X[picks[i]]
    else:
        raise ValueError('Invalid init parameter')
    Z = right_stochastic_matrix((T, k), rng)
    return W, Z


# ---------------------------------------------------------------------------
# k-means Lloyd (scikit-learn 1.9.0, third party; see module docstring)
# ---------------------------------------------------------------------------

def kmeans_lloyd(X, init_centres, tol=1e-4, max_iter=300):
    """Restatement of ``KMeans(init=array, n_init=1, algorithm='lloyd').fit``.

    sklearn/cluster/_kmeans.py (1.9.0): data are mean-centred (:1486-1493),
    ``tol`` is scaled by the mean per-feature variance (:285-294), the E-step
    takes the first minimum of ||c||^2 - 2 x.c (_k_means_lloyd.pyx:193-213),
    empty clusters are relocated to the points farthest from their centres
    (_k_means_common.pyx:167-212), iteration stops on identical labels or
    squared centre shift <= tol, followed by one more E-step when not strictly
    converged (:703-757).  Returns labels (int32), centres, inertia, n_iter.
    """
    X = np.asarray(X, dtype=np.float64)
    mean = X.mean(axis=0)
    Xc = X - mean
    centres = np.asarray(init_centres, dtype=np.float64) - mean
    k = centres.shape[0]
    tol_abs = np.mean(np.var(Xc, axis=0)) * tol
    x_sq = (Xc * Xc).sum(axis=1)

    def e_step(c):
        scores = (c * c).sum(axis=1)[np.newaxis, :] - 2.0 * Xc.dot(c.T)
        return np.argmin(scores, axis=1).astype(np.int32)

    def inertia_of(labels, c):
        diff = Xc - c[labels]
        return float((diff * diff).sum())

    labels = np.full(X.shape[0], -1, dtype=np.int32)
    labels_old = labels.copy()
    strict = False
    n_iter = 0
    for it in range(max_iter):
        n_iter = it + 1
        labels = e_step(centres)
        counts = np.bincount(labels, minlength=k).astype(np.float64)
        sums = np.zeros_like(centres)
        np.add.at(sums, labels, Xc)
        empty = np.where(counts == 0)[0]
        if empty.size:
            diff = Xc - centres[labels]
            dist = (diff * diff).sum(axis=1)
            far = np.argpartition(dist, -empty.size)[:-empty.size - 1:-1]
            if np.max(dist) != 0:
                for idx, cid in enumerate(empty):
                    far_idx = far[idx]
                    old = labels[far_idx]
                    sums[old] -= Xc[far_idx]
                    sums[cid] = Xc[far_idx]
                    counts[cid] = 1
                    counts[old] -= 1
        new_centres = sums.copy()
        for j in range(k):                       # _average_centers: scale by 1/w
            if counts[j] > 0:
                new_centres[j] *= 1.0 / counts[j]
        shift = ((new_centres - centres) ** 2).sum()
        centres = new_centres
        if np.array_equal(labels, labels_old):
            strict = True
            break
        if shift <= tol_abs:
            break
        labels_old = labels.copy()
    if not strict:
        labels = e_step(centres)
    return labels, centres + mean, inertia_of(labels, centres), n_iter
