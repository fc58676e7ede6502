"""GPU parity tests of the individual kernels behind the C ABI (``-m gpu``).

Each test drives ``libcdr_b200.so`` through the ctypes binding on seeded inputs
and compares with the CPU oracle (``oracle/``, pinned against the reference by
``tests/test_oracle_golden.py``), with NumPy for plain contractions, or with the
golden vectors generated from the reference itself.

Tolerances: index / label outputs are bit-exact; fp64 contractions agree to
1e-12 relative (summation order differs); iterative solvers agree to 1e-9 when
the iteration count is pinned and to the solver's own stopping tolerance
(epsilon_two = 1e-6 on the projected-gradient norm) when run to convergence.
"""

import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
if not torch.cuda.is_available():          # pragma: no cover
    pytest.skip('needs a CUDA device', allow_module_level=True)

from oracle import convex_oracle as orc                       # noqa: E402
from convex_dim_red import _backend as be                       # noqa: E402
from convex_dim_red import simplex_projection as sp             # noqa: E402
from convex_dim_red.furthest_sum import furthest_sum            # noqa: E402
from convex_dim_red.spg import quad_simplex_spg                 # noqa: E402


def close(a, b, rtol=1e-12, atol=1e-13):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def test_device_is_sm100():
    assert be.library().cdr_device_check() == 0


# ---------------------------------------------------------------- simplex
@pytest.mark.parametrize('n', [1, 2, 3, 5, 8, 17, 64, 317])
def test_simplex_vector_golden(golden, n):
    x = golden['simplex/vec%d/x' % n]
    close(sp.simplex_project_vector(x), golden['simplex/vec%d/out' % n], rtol=1e-13, atol=1e-15)


def test_simplex_known_answers(golden):
    # reference tests/test_simplex_projection.py:13-57, 166-176
    close(sp.simplex_project_vector(np.array([0.8, 0.8])), [0.5, 0.5], atol=1e-15)
    close(sp.simplex_project_vector(np.array([0.5, -0.5])), [1.0, 0.0], atol=1e-15)
    close(sp.simplex_project_vector(np.array([3.0])), [1.0], atol=1e-15)
    A = np.array([[0.5, 0.5], [0.5, 1.0], [0.0, -0.5]])
    close(sp.simplex_project_rows(A), [[0.5, 0.5], [0.25, 0.75], [0.75, 0.25]], atol=1e-15)
    for name in ('feasible', 'ties'):
        close(sp.simplex_project_vector(golden['simplex/%s/x' % name]),
              golden['simplex/%s/out' % name], rtol=1e-13, atol=1e-15)


def test_simplex_rows_cols_golden(golden):
    A = golden['simplex/rows/A']
    close(sp.simplex_project_rows(A), golden['simplex/rows/out'], rtol=1e-13, atol=1e-15)
    close(sp.simplex_project_columns(A), golden['simplex/cols/out'], rtol=1e-13, atol=1e-15)
    A = golden['simplex/rows_wide/A']
    close(sp.simplex_project_rows(A), golden['simplex/rows_wide/out'], rtol=1e-12, atol=1e-15)
    out = np.empty_like(A)
    assert sp.simplex_project_rows(A, out=out) is out
    close(out, golden['simplex/rows_wide/out'], rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize('shape', [(8, 1620), (3, 18000), (2, 30000), (341, 317), (5, 1)])
def test_simplex_rows_vs_oracle(shape):
    rs = np.random.RandomState(shape[1])
    A = rs.standard_normal(shape) * (0.02 if shape[1] > 1000 else 1.0) + 1.0 / shape[1]
    out = sp.simplex_project_rows(A)
    close(out, orc.simplex_project_rows(A), rtol=1e-11, atol=1e-15)
    assert np.all(out >= 0)
    close(out.sum(axis=1), 1.0, atol=1e-12)        # reference test: row sums within 1e-14/1e-15


def test_simplex_empty_and_idempotent():
    assert sp.simplex_project_rows(np.zeros((0, 4))).shape == (0, 4)
    rs = np.random.RandomState(0)
    Z = orc.right_stochastic_matrix((50, 9), rs)
    close(sp.simplex_project_rows(Z), Z, atol=1e-15)
    close(sp.simplex_project_columns(Z.T.copy()), Z.T, atol=1e-15)


# ---------------------------------------------------------------- QP
@pytest.mark.parametrize('k', [2, 3, 8, 16])
def test_qp_single_golden(golden, k):
    A, b, x0 = (golden['qp/k%d/%s' % (k, n)] for n in ('A', 'b', 'x0'))
    close(quad_simplex_spg(A, b, x0, max_iterations=3), golden['qp/k%d/x_it3' % k], rtol=1e-9)
    close(quad_simplex_spg(A, b, x0), golden['qp/k%d/x' % k], rtol=0, atol=5e-6)


def _qp_case(T, k, seed):
    rs = np.random.RandomState(seed)
    M = rs.standard_normal((k + 4, k))
    A = M.T.dot(M)
    B = rs.standard_normal((k, T)) * 2.0
    Z0 = orc.right_stochastic_matrix((T, k), rs)
    alpha = rs.uniform(0.8, 1.2, size=k)
    return A, B, Z0, alpha


def _run_qp(A, alpha, B, Z0, k_by_T, **kw):
    T, k = Z0.shape
    dA, dB, dZ = be.to_device(A), be.to_device(B), be.to_device(Z0)
    dal = None if alpha is None else be.to_device(alpha)
    n_iter = torch.zeros(T, dtype=torch.int32, device='cuda')
    n_feval = torch.zeros(T, dtype=torch.int32, device='cuda')
    sb = (1, B.shape[1]) if k_by_T else (B.shape[1], 1)
    be.quad_simplex_spg_batched(dA, dal, dB, sb[0], sb[1], dZ, T, k, be.make_spg_params(kw),
                                n_iter, n_feval)
    return dZ.cpu().numpy(), n_iter.cpu().numpy(), n_feval.cpu().numpy()


@pytest.mark.parametrize('k', [1, 3, 8, 9, 16, 20, 33, 64])
def test_qp_batched_pinned_iterations(k):
    """With the iteration count pinned the trajectories agree to rounding."""
    T = 203
    A, B, Z0, alpha = _qp_case(T, k, 100 + k)
    da = np.diag(alpha)
    for its in (1, 4):
        Z, n_it, n_fe = _run_qp(A, alpha, B, Z0, True, max_iterations=its)
        ref, r_it, r_fe = orc.weights_update(da.dot(A).dot(da), da.dot(B), Z0, True,
                                             return_counts=True, max_iterations=its)
        close(Z, ref, rtol=1e-9, atol=1e-12)
        assert np.array_equal(n_it, r_it)
        assert np.mean(n_fe == r_fe) > 0.98       # a rounding-level Armijo tie may flip a trial
        close(Z.sum(axis=1), 1.0, atol=1e-12)
        assert np.all(Z >= 0)


@pytest.mark.parametrize('k', [1, 2, 5, 8, 13, 64])
def test_qp_projection_of_the_start_on_hard_inputs(k):
    """With max_iterations = 0 the batched solver returns P(z0) (spg.py:300), i.e. the
    per-sample threshold search on its own (k <= 8: rank / ballot form, k > 8: Michelot):
    ties, vertices, all-equal, already feasible, feasible up to rounding, huge and tiny
    magnitudes, against simplex_projection.py:13-27 as restated by the oracle."""
    rs = np.random.RandomState(40 + k)
    rows = [np.full(k, 1.0 / k), np.zeros(k), np.full(k, 7.5), -np.arange(k, dtype=float)]
    e = np.zeros(k)
    e[k // 2] = 1.0
    rows.append(e.copy())
    e[k // 2] = 100.0
    rows.append(e.copy())
    v = rs.rand(k)
    v[: max(1, k // 2)] = v[0]
    rows.append(v.copy())                                  # ties among the largest
    v = rs.rand(k)
    v /= v.sum()
    rows += [v.copy(), v + 1e-17, v * (1 + 1e-15), v - 1e-16]
    rows += [rs.standard_normal(k) * 1e6, rs.standard_normal(k) * 1e-9, 1.0 + rs.standard_normal(k) * 1e-13]
    for scale in (0.1, 1.0, 10.0):
        rows += list(rs.standard_normal((100, k)) * scale)
    Z0 = np.array(rows)
    T = Z0.shape[0]
    Z, n_it, _ = _run_qp(np.eye(k), None, np.zeros((k, T)), Z0, True, max_iterations=0)
    ref = orc.simplex_project_rows(Z0)
    scale = np.maximum(1.0, np.abs(Z0).max(axis=1))[:, None]
    err = np.abs(Z - ref).max(axis=1)
    tol = 4e-16 * k * scale[:, 0]
    assert np.all(err <= tol), (np.argmax(err / tol), Z0[np.argmax(err / tol)], err.max())
    assert np.all(Z >= 0)
    # the outputs sum to one as accurately as the oracle's do
    sum_err, ref_sum_err = np.abs(Z.sum(axis=1) - 1.0), np.abs(ref.sum(axis=1) - 1.0)
    worst = np.argmax(sum_err - ref_sum_err - tol)
    assert np.all(sum_err <= ref_sum_err + tol), (worst, Z0[worst], sum_err[worst], ref_sum_err[worst])


@pytest.mark.parametrize('k', [3, 8, 20, 64])
def test_qp_batched_converged(k):
    T = 150
    A, B, Z0, _ = _qp_case(T, k, 7 + k)
    Z, n_it, _ = _run_qp(A, None, B.T.copy(), Z0, False)
    ref, r_it, _ = orc.weights_update(A, B.T.copy(), Z0, False, return_counts=True)
    close(Z, ref, rtol=0, atol=1e-5)
    # objective values agree far more tightly than the minimisers
    def obj(Zm):
        return 0.5 * np.einsum('ti,ij,tj->t', Zm, A, Zm) - np.einsum('ti,it->t', Zm, B)
    close(obj(Z), obj(ref), rtol=1e-9, atol=1e-10)
    assert abs(np.median(n_it) - np.median(r_it)) <= 2


def test_qp_batched_golden(golden):
    A, CK, Z0 = golden['qp/batch/A'], golden['qp/batch/CK'], golden['qp/batch/Z0']
    Z, _, _ = _run_qp(A, None, CK, Z0, True, max_iterations=1)
    close(Z, golden['qp/batch/Z_aa_it1'], rtol=1e-9)
    Z, _, _ = _run_qp(A, None, CK, Z0, True)
    close(Z, golden['qp/batch/Z_aa'], rtol=0, atol=5e-6)
    Z, _, _ = _run_qp(A, golden['qp/batch/alpha2'], CK, Z0, True)
    close(Z, golden['qp/batch/Z_aa_alpha2'], rtol=0, atol=5e-6)


def test_qp_ragged_sizes_and_done_flag():
    for T in (1, 2, 5, 127, 129):
        A, B, Z0, _ = _qp_case(T, 8, T)
        Z, _, _ = _run_qp(A, None, B, Z0, True, max_iterations=2)
        close(Z, orc.weights_update(A, B, Z0, True, max_iterations=2), rtol=1e-9, atol=1e-12)
    # a set `done` flag turns the launch into a no-op
    A, B, Z0, _ = _qp_case(16, 8, 3)
    state = be.DeviceState(1e-6, 10, 'abs_delta_f', True)
    state.write_field('done', 1)
    dZ = be.to_device(Z0)
    be.quad_simplex_spg_batched(be.to_device(A), None, be.to_device(B), 1, 16, dZ, 16, 8,
                                be.make_spg_params({}), flags=state.buf)
    assert np.array_equal(dZ.cpu().numpy(), Z0)


# ---------------------------------------------------------------- streaming contractions
@pytest.mark.parametrize('T,d,k', [(64, 96, 8), (203, 777, 5), (50, 33, 1), (130, 1000, 16),
                                   (97, 515, 20), (77, 260, 33), (70, 300, 64), (1620, 2048, 8)])
def test_reduce_samples_and_features(T, d, k):
    rs = np.random.RandomState(T + d + k)
    X = rs.standard_normal((T, d))
    L = rs.standard_normal((k, T))
    Z = rs.uniform(size=(T, k))
    E = rs.standard_normal((k, k))
    Xd = be.to_device_padded(X)
    ws = be.Workspace(T, d, k)
    out = be.zeros(k, Xd.stride(0))
    Ld = be.to_device_padded(L)
    be.reduce_samples(Ld, Ld.stride(0), 1, Xd, T, d, k, out, ws)
    close(be.to_host(out, k, d), L.dot(X), rtol=1e-12, atol=1e-11)
    assert np.all(be.to_host(out)[:, d:] == 0)                 # padding stays zero
    # strided left operand (weights.T) with the k x k epilogue
    Zd = be.to_device(Z)
    be.reduce_samples(Zd, 1, k, Xd, T, d, k, out, ws, E=be.to_device(E))
    close(be.to_host(out, k, d), E.dot(Z.T.dot(X)), rtol=1e-11, atol=1e-10)
    # reduce over features
    M = rs.standard_normal((k, d))
    Md = be.to_device_padded(M)
    outT = be.zeros(k, be.round_up(T))
    be.reduce_features(Md, Xd, T, d, k, outT, ws)
    close(be.to_host(outT, k, T), M.dot(X.T), rtol=1e-12, atol=1e-11)


def test_reduce_samples_split_path_and_determinism():
    # few features, many samples: the sample axis is split and reduced in fixed order
    rs = np.random.RandomState(5)
    T, d, k = 2000, 100, 8
    X = rs.standard_normal((T, d))
    L = rs.standard_normal((k, T))
    Xd, Ld = be.to_device_padded(X), be.to_device_padded(L)
    ws = be.Workspace(T, d, k)
    out1, out2 = be.zeros(k, Xd.stride(0)), be.zeros(k, Xd.stride(0))
    be.reduce_samples(Ld, Ld.stride(0), 1, Xd, T, d, k, out1, ws)
    be.reduce_samples(Ld, Ld.stride(0), 1, Xd, T, d, k, out2, ws)
    close(be.to_host(out1, k, d), L.dot(X), rtol=1e-12, atol=1e-11)
    assert torch.equal(out1, out2)


def test_gram_frobenius_residual():
    rs = np.random.RandomState(9)
    T, d, k = 150, 400, 6
    X = rs.standard_normal((T, d))
    Xd = be.to_device_padded(X)
    K = be.gram(Xd, T, d)
    close(be.to_host(K, T, T), X.dot(X.T), rtol=1e-12, atol=1e-10)
    close(float(be.frobenius_sq(Xd, T, d).item()), np.sum(X * X), rtol=1e-13)
    Z = orc.right_stochastic_matrix((T, k), rs)
    A = rs.standard_normal((k, d))
    out, part = be.zeros(1), be.zeros(T)
    Ad, Zd = be.to_device_padded(A), be.to_device(Z)
    be.check(be.library().cdr_residual_sq(Xd.data_ptr(), Xd.stride(0), T, d, Zd.data_ptr(), k,
                                          Ad.data_ptr(), Ad.stride(0), out.data_ptr(),
                                          part.data_ptr(), be.stream_ptr()), 'residual')
    close(float(out.item()), np.linalg.norm(X - Z.dot(A)) ** 2, rtol=1e-13)


@pytest.mark.parametrize('T,d', [(1, 5), (37, 70), (128, 64), (129, 333), (257, 1000), (700, 4180)])
def test_syrk_gram_matches_numpy_and_the_slab_path(T, d):
    """K = X X' (archetypal_analysis.py:1032) by the SYRK kernel: symmetric to the bit, equal
    to NumPy and to the round-1 slab path to rounding, deterministic; tile edges (T not a
    multiple of 128, d not a multiple of 32) included."""
    rs = np.random.RandomState(T + d)
    X = rs.standard_normal((T, d))
    Xd = be.to_device_padded(X)
    K = be.gram(Xd, T, d)
    Kh = be.to_host(K, T, T)
    assert np.array_equal(Kh, Kh.T)
    scale = np.abs(X.dot(X.T)).max()
    close(Kh, X.dot(X.T), rtol=0, atol=1e-13 * scale * np.sqrt(d))
    close(Kh, be.to_host(be.gram_slabs(Xd, T, d), T, T), rtol=0, atol=1e-13 * scale * np.sqrt(d))
    assert torch.equal(K, be.gram(Xd, T, d))
    assert float(K[:, T:].abs().sum()) == 0.0


@pytest.mark.parametrize('k,n', [(1, 10), (5, 33), (8, 1620), (20, 700), (64, 5000)])
def test_small_gram(k, n):
    rs = np.random.RandomState(k * n)
    A = rs.standard_normal((k, n))
    Zm = rs.standard_normal((n, k))
    Ad, Zd = be.to_device(A), be.to_device(Zm)
    o1, o2, o3 = be.zeros(k, k), be.zeros(k, k), be.zeros(k, k)
    ws = be.Workspace(8, 8, k)
    be.small_gram([(Ad, n, 1, k, Ad, n, 1, k, n, o1, 1.0, 0),
                   (Ad, n, 1, k, Zd, 1, k, k, n, o2, 0.5, 0),
                   (Ad, n, 1, k, Ad, n, 1, k, n, o3, 1.0, 1)], ws)
    close(o1.cpu().numpy(), A.dot(A.T), rtol=1e-12, atol=1e-11)
    close(o2.cpu().numpy(), 0.5 * A.dot(Zm), rtol=1e-12, atol=1e-11)
    pair = ((A[:, None, :] - A[None, :, :]) ** 2).sum(-1)
    close(o3.cpu().numpy(), pair, rtol=1e-12, atol=1e-11)


@pytest.mark.parametrize('singular', [False, True])
@pytest.mark.parametrize('k', [1, 2, 3, 8, 20, 64])
def test_gpnh_solve_matrix_matches_lstsq(k, singular):
    """Cholesky fast path (regular Z'Z) and Jacobi pseudo-inverse (singular Z'Z: the
    minimum-norm branch of numpy.linalg.lstsq) against lstsq itself."""
    rs = np.random.RandomState(k)
    T, d = 300, 57
    Z = orc.right_stochastic_matrix((T, k), rs)
    if singular and k > 3:
        Z[:, 2] = 0.0
        Z /= Z.sum(axis=1)[:, None]
    ZtZ = Z.T.dot(Z)
    rhs = rs.standard_normal((k, 11))
    for lam in (0.0, 3.2):
        P = be.zeros(k, k)
        be.check(be.library().cdr_gpnh_solve_matrix(
            be.to_device(ZtZ).data_ptr(), k, T, d, lam, P.data_ptr(), None, 0, None,
            be.stream_ptr()), 'solve')
        lhs = ZtZ / T + lam * orc.gpnh_GW(d, k)
        ref = np.linalg.lstsq(lhs, rhs / T, rcond=None)[0]
        got = P.cpu().numpy().dot(rhs)
        close(got, ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())
    S = rs.standard_normal((k + 2, k))
    S = S.T.dot(S)
    P = be.zeros(k, k)
    be.check(be.library().cdr_sym_pinv(be.to_device(S).data_ptr(), k, P.data_ptr(), None,
                                       be.stream_ptr()), 'pinv')
    close(P.cpu().numpy().dot(S), np.eye(k), rtol=0, atol=1e-8)


# ---------------------------------------------------------------- furthest sum
def test_furthest_sum_golden(golden):
    D = golden['fs/D']
    for i, (k, start, extra) in enumerate(golden['fs/cases']):
        sel = furthest_sum(D, int(k), int(start), list(golden['fs/case%d/exclude' % i]), int(extra))
        assert sel.dtype == np.int64
        assert np.array_equal(sel, golden['fs/case%d/selected' % i])
    K3 = golden['fs/k3/D']
    for start in range(3):
        for extra in range(1, 11):
            assert set(furthest_sum(K3, 2, start, [], extra)) == {0, 2}
    assert np.array_equal(furthest_sum(K3, 2, 1, [], 3), golden['fs/k3/sel'])


def test_furthest_sum_ties_match_stable_sort_order():
    """Integer-valued dissimilarities produce exact ties; the picks must follow the
    reference's stable-sort / pop(-1) order bit for bit."""
    rs = np.random.RandomState(123)
    for trial in range(6):
        n = 30 + trial
        D = rs.randint(0, 4, size=(n, n)).astype(np.float64)
        D = D + D.T
        np.fill_diagonal(D, 0.0)
        for k, start, extra, excl in ((4, 0, 10, []), (7, n - 1, 13, [1, 2]), (1, 3, 4, []),
                                      (n - 3, 2, 5, [0])):
            ref = orc.furthest_sum(D, k, start, excl, extra)
            got = furthest_sum(D, k, start, excl, extra)
            assert np.array_equal(got, np.asarray(ref, dtype=np.int64)), (trial, k, start)


def test_furthest_sum_validation():
    D = np.zeros((4, 4))
    assert list(furthest_sum(D, 0, 0)) == []
    with pytest.raises(ValueError):
        furthest_sum(np.zeros((3, 4)), 2, 0)
    with pytest.raises(ValueError):
        furthest_sum(D, 2, 7)
    with pytest.raises(ValueError):
        furthest_sum(D, 2, 1, [1])
    with pytest.raises(ValueError):
        furthest_sum(D, 4, 0, [1])
    pts = np.array([[0, 0], [.5, .5], [.2, .7], [.4, .1], [1, 0], [.6, .6], [0, 1],
                    [.3, .3], [.7, .2], [1, 1]], dtype=float)
    Dm = np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1))
    assert set(furthest_sum(Dm, 4, 1, [], 10)) == {0, 4, 6, 9}


def test_furthest_sum_large_random_vs_oracle():
    rs = np.random.RandomState(77)
    P = rs.standard_normal((700, 9))
    D = np.sqrt(np.maximum(((P[:, None, :] - P[None, :, :]) ** 2).sum(-1), 0))
    ref = orc.furthest_sum(D, 8, 11, [], 10)
    assert np.array_equal(furthest_sum(D, 8, 11, [], 10), np.asarray(ref, dtype=np.int64))


# ---------------------------------------------------------------- TMA-pipelined passes
@pytest.mark.parametrize('T,d,k', [(300, 9000, 8), (1000, 5000, 16), (130, 20000, 5),
                                   (257, 4800, 20), (70, 6016, 32), (1620, 44000, 8),
                                   (64, 9472, 1), (203, 12345, 3)])
def test_tma_streaming_passes(T, d, k, monkeypatch):
    """Shapes that take the cp.async.bulk + mbarrier pipeline: against NumPy and against
    the direct-load kernels (CDR_DISABLE_TMA=1)."""
    rs = np.random.RandomState(T + d + k)
    X = rs.standard_normal((T, d))
    Z = rs.uniform(size=(T, k))
    L = rs.standard_normal((k, T))
    E = rs.standard_normal((k, k))
    M = rs.standard_normal((k, d))
    Xd, Zd, Ld, Ed, Md = (be.to_device_padded(X), be.to_device(Z), be.to_device_padded(L),
                          be.to_device(E), be.to_device_padded(M))
    ws = be.Workspace(T, d, k)
    res = {}
    for mode in ('tma', 'direct'):
        monkeypatch.setenv('CDR_DISABLE_TMA', '1' if mode == 'direct' else '0')
        o1, o2 = be.zeros(k, Xd.stride(0)), be.zeros(k, Xd.stride(0))
        o3 = be.zeros(k, be.round_up(T))
        be.reduce_samples(Ld, Ld.stride(0), 1, Xd, T, d, k, o1, ws)
        be.reduce_samples(Zd, 1, k, Xd, T, d, k, o2, ws, E=Ed)
        be.reduce_features(Md, Xd, T, d, k, o3, ws)
        torch.cuda.synchronize()
        res[mode] = (be.to_host(o1), be.to_host(o2), be.to_host(o3, k, T))
        close(res[mode][0][:, :d], L.dot(X), rtol=1e-12, atol=1e-10)
        assert np.all(res[mode][0][:, d:] == 0)
        close(res[mode][1][:, :d], E.dot(Z.T.dot(X)), rtol=1e-11, atol=1e-9)
        close(res[mode][2], M.dot(X.T), rtol=1e-12, atol=1e-10)
    for a, b in zip(res['tma'], res['direct']):
        close(a, b, rtol=1e-12, atol=1e-10)
    # deterministic: a second run is bit-identical
    monkeypatch.setenv('CDR_DISABLE_TMA', '0')
    o3b = be.zeros(k, be.round_up(T))
    be.reduce_features(Md, Xd, T, d, k, o3b, ws)
    assert np.array_equal(be.to_host(o3b, k, T), res['tma'][2])
