"""Spectral projected gradient solvers (reference ``spg.py``).

``spg`` keeps the reference's callback interface: the control flow (non-monotone
Armijo search, Barzilai-Borwein step, stopping tests) is host logic identical to
``spg.py:46-283`` and works on whatever array type the callbacks return.  The
library itself does not use it on the hot path: the dictionary update runs the
same iteration fused on the device (``cdr_aa_spg_*``), and the per-sample QPs run
in ``cdr_quad_simplex_spg_batched``.  ``spg`` is still used for the k-vector
scale-factor problem (``delta != 0``).
"""

import time
import warnings

import numpy as np

from . import _backend as be


def spg_line_search_step_length(current_step_length, delta, f_old, f_new,
                                sigma_one=0.1, sigma_two=0.9):
    """Return next step length for line search (spg.py:19-33)."""
    step_length_tmp = (-0.5 * current_step_length ** 2 * delta /
                       (f_new - f_old - current_step_length * delta))
    if sigma_one <= step_length_tmp <= sigma_two * current_step_length:
        return step_length_tmp
    return 0.5 * current_step_length


def spg_line_search_cauchy_step_size(beta, sksk, alpha_min=1e-3, alpha_max=1e3):
    """Return next value of Cauchy step size parameter (spg.py:36-43)."""
    if beta <= 0:
        return alpha_max
    return min(alpha_max, max(alpha_min, sksk / beta))


def spg(f, df, x0, project=None, gamma=1e-4, memory=1,
        sigma_one=0.1, sigma_two=0.9, lambda_min=1e-10,
        alpha0=None, alpha_min=1e-5, alpha_max=1e3,
        epsilon_one=1e-10, epsilon_two=1e-6,
        use_infinity_norm=True, verbose=0,
        max_iterations=10000, max_feval=1000000):
    """Perform gradient descent steps with non-monotone line-search.

    Same parameters, defaults, warnings and return value
    ``(sol, fmin, n_iter, n_feval)`` as the reference (spg.py:46-283).
    """
    is_multivariate = not np.isscalar(x0)
    x = x0.copy() if is_multivariate else x0
    if project is not None:
        x = project(x)

    alpha = alpha0
    f_mem = np.zeros(memory)
    f_old = f(x)
    n_feval = 1

    if verbose:
        print('{:<12s} | {:<12s} | {:<13s} | {:<13s} | {:<12s}'.format(
            'n_iter', 'n_feval', 'f', 'conv_crit', 'time'))
        print('-' * 79)
        print('{:12d} | {:12d} | {: 12.6e} | {: 12.6e} | {: 12.6e}'.format(
            0, n_feval, f_old, -1, 0))

    has_converged = False
    n_iter = -1
    for n_iter in range(max_iterations):
        start_time = time.perf_counter()
        x_old = x.copy() if is_multivariate else x
        gk = df(x)

        if alpha is None:
            if project is None:
                alpha = 1.0 / np.max(np.abs(gk))
            else:
                alpha_inv = np.max(np.abs(project(x - gk) - x))
                alpha = 1.0 / alpha_inv if abs(alpha_inv) > 1e-12 else 1.0

        dk = -alpha * gk
        if project is not None:
            dk = project(x + dk)
            dk -= x

        f_mem = np.roll(f_mem, 1)
        f_mem[0] = f_old
        f_max = None
        for previous_value in f_mem:
            if f_max is None or previous_value >= f_max:
                f_max = previous_value

        delta = np.sum(dk * gk)
        lam = 1
        x = x_old + dk
        f_new = f(x)
        n_feval += 1

        while f_new > f_max + gamma * lam * delta:
            lam = spg_line_search_step_length(
                lam, delta, f_old, f_new, sigma_one=sigma_one, sigma_two=sigma_two)
            x = x_old + lam * dk
            f_new = f(x)
            n_feval += 1
            if abs(lam) < lambda_min:
                warnings.warn('step size below tolerance in SPG line search', UserWarning)
                break

        yk = gk.copy() if is_multivariate else gk
        gk = df(x)
        yk = gk - yk

        sksk = lam ** 2 * np.sum(dk * dk)
        betak = lam * np.sum(dk * yk)
        alpha = spg_line_search_cauchy_step_size(
            betak, sksk, alpha_min=alpha_min, alpha_max=alpha_max)

        f_old = f(x)
        n_feval += 1

        res = -gk if project is None else project(x - gk) - x
        res_norm = np.sum(res ** 2) ** 0.5
        end_time = time.perf_counter()

        if verbose:
            print('{:12d} | {:12d} | {: 12.6e} | {: 12.6e} | {: 12.6e}'.format(
                n_iter + 1, n_feval, f_old, res_norm, end_time - start_time))

        has_converged = res_norm < epsilon_two
        if use_infinity_norm:
            has_converged = has_converged or np.max(np.abs(res)) < epsilon_one

        if has_converged:
            if verbose:
                print('-' * 79)
                print('*** Converged at iteration {:d} ***'.format(n_iter + 1))
            break

        if n_feval > max_feval:
            warnings.warn('maximum number of function evaluations exceeded in SPG',
                          UserWarning)
            break

    if n_iter == max_iterations - 1 and not has_converged:
        warnings.warn('maximum number of iterations exceeded in SPG', UserWarning)

    return x, f_old, n_iter, n_feval


def quad_simplex_spg(A, b, x0, gamma=1e-4, memory=1,
                     sigma_one=0.1, sigma_two=0.9, lambda_min=1e-10,
                     alpha0=-1.0, alpha_min=1e-5, alpha_max=1e3,
                     epsilon_one=1e-10, epsilon_two=1e-6,
                     max_iterations=1000, max_feval=2000):
    """Solve one quadratic program constrained to the standard simplex.

    Minimises ``x.T A x / 2 + b.x`` (spg.py:286-398) with the batched CUDA
    solver on a batch of one.
    """
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    k = b.shape[0]
    params = be.make_spg_params(dict(
        gamma=gamma, memory=memory, sigma_one=sigma_one, sigma_two=sigma_two,
        lambda_min=lambda_min, alpha0=alpha0, alpha_min=alpha_min, alpha_max=alpha_max,
        epsilon_one=epsilon_one, epsilon_two=epsilon_two, max_iterations=max_iterations,
        max_feval=max_feval))
    dA = be.to_device(A)
    dB = be.to_device(-b[np.newaxis, :])       # the kernel's linear term is -B[t]
    dZ = be.to_device(np.asarray(x0, dtype=np.float64)[np.newaxis, :])
    be.quad_simplex_spg_batched(dA, None, dB, k, 1, dZ, 1, k, params)
    return dZ.cpu().numpy()[0]
