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


class _SpgTrace:
    """Iteration log in the reference's verbose format (spg.py:159-164, 256-259)."""

    HEADER = '{:<12s} | {:<12s} | {:<13s} | {:<13s} | {:<12s}'
    ROW = '{:12d} | {:12d} | {: 12.6e} | {: 12.6e} | {: 12.6e}'

    def __init__(self, enabled):
        self.enabled = bool(enabled)

    def start(self, n_feval, value):
        if self.enabled:
            print(self.HEADER.format('n_iter', 'n_feval', 'f', 'conv_crit', 'time'))
            print('-' * 79)
            print(self.ROW.format(0, n_feval, value, -1, 0))

    def row(self, *fields):
        if self.enabled:
            print(self.ROW.format(*fields))

    def converged(self, iteration):
        if self.enabled:
            print('-' * 79)
            print('*** Converged at iteration {:d} ***'.format(iteration))


def _first_step_length(x, grad, project):
    """Inverse infinity norm of the first (projected) increment (spg.py:178-189)."""
    if project is None:
        return 1.0 / np.max(np.abs(grad))
    reach = np.max(np.abs(project(x - grad) - x))
    return 1.0 / reach if abs(reach) > 1e-12 else 1.0


def _reference_value(history):
    """Largest stored objective value, scanning with ``>=`` like spg.py:199-203."""
    best = None
    for value in history:
        if best is None or value >= best:
            best = value
    return best


def spg(f, df, x0, project=None, gamma=1e-4, memory=1,
        sigma_one=0.1, sigma_two=0.9, lambda_min=1e-10,
        alpha0=None, alpha_min=1e-5, alpha_max=1e3,
        epsilon_one=1e-10, epsilon_two=1e-6,
        use_infinity_norm=True, verbose=0,
        max_iterations=10000, max_feval=1000000):
    """Perform gradient descent steps with non-monotone line-search.

    Same parameters, defaults, warnings and return value ``(sol, fmin, n_iter, n_feval)``
    as the reference's ``spg`` (spg.py:46-283): ``f`` / ``df`` are callables returning the
    objective and its gradient, ``project`` an optional projection onto the feasible set.
    Works on scalars and on arrays of any type the callables accept.
    """
    copy = (lambda v: v.copy()) if not np.isscalar(x0) else (lambda v: v)
    point = copy(x0)
    if project is not None:
        point = project(point)

    step = alpha0                      # None -> initialised from the first gradient
    history = np.zeros(memory)         # zeros, not NaN (spg.py:153)
    value = f(point)
    evaluations = 1
    log = _SpgTrace(verbose)
    log.start(evaluations, value)

    done = False
    iteration = -1
    for iteration in range(max_iterations):
        tic = time.perf_counter()
        origin = copy(point)
        grad = df(point)
        if step is None:
            step = _first_step_length(point, grad, project)

        direction = -step * grad
        if project is not None:
            direction = project(point + direction)
            direction -= point

        history = np.roll(history, 1)
        history[0] = value
        ceiling = _reference_value(history)

        slope = np.sum(direction * grad)
        lam = 1
        point = origin + direction
        trial = f(point)
        evaluations += 1
        while trial > ceiling + gamma * lam * slope:
            lam = spg_line_search_step_length(lam, slope, value, trial,
                                              sigma_one=sigma_one, sigma_two=sigma_two)
            point = origin + lam * direction
            trial = f(point)
            evaluations += 1
            if abs(lam) < lambda_min:
                warnings.warn('step size below tolerance in SPG line search', UserWarning)
                break

        previous_grad = copy(grad)
        grad = df(point)
        change = grad - previous_grad
        step = spg_line_search_cauchy_step_size(
            lam * np.sum(direction * change), lam ** 2 * np.sum(direction * direction),
            alpha_min=alpha_min, alpha_max=alpha_max)

        value = f(point)
        evaluations += 1

        residual = -grad if project is None else project(point - grad) - point
        residual_norm = np.sum(residual ** 2) ** 0.5
        log.row(iteration + 1, evaluations, value, residual_norm, time.perf_counter() - tic)

        done = residual_norm < epsilon_two
        if use_infinity_norm:
            done = done or np.max(np.abs(residual)) < epsilon_one
        if done:
            log.converged(iteration + 1)
            break
        if evaluations > max_feval:
            warnings.warn('maximum number of function evaluations exceeded in SPG', UserWarning)
            break

    if iteration == max_iterations - 1 and not done:
        warnings.warn('maximum number of iterations exceeded in SPG', UserWarning)

    return point, value, iteration, evaluations


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
