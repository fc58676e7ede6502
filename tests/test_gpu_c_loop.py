"""The whole-iteration C entry points on their own: a fit driven through ctypes only --
cdr_gpnh_prepare_enqueue / cdr_gpnh_iterate_enqueue and cdr_aa_prepare_enqueue /
cdr_aa_iterate_enqueue on caller-allocated device buffers, no engine class in between --
against the CPU oracle.  This is the sequence INTEGRATION.md section 3 shows in C; it replaces
the loops `_iterate_gpnh_convex_coding` (gpnh_convex_coding.py:282-402) and `_iterate_aa`
(archetypal_analysis.py:534-670).  Shapes: one the three- / eight-kernel paths cover (wide
enough for the strip kernels) and one they do not (general kernel sequence)."""

import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
if not torch.cuda.is_available():          # pragma: no cover
    pytest.skip('needs a CUDA device', allow_module_level=True)

from oracle import convex_oracle as orc                                     # noqa: E402
from convex_dim_red import _backend as be                                   # noqa: E402
from convex_dim_red.datasets import synthetic_field                         # noqa: E402


def _state(tolerance, max_iterations, trace):
    st = be.LoopState()
    st.tolerance, st.max_iterations, st.stopping_rule, st.require_monotone = tolerance, max_iterations, 0, 1
    st.trace_data = trace
    return torch.from_numpy(np.frombuffer(bytes(st), dtype=np.uint8).copy()).cuda()


def _read(buf):
    return be.LoopState.from_buffer_copy(buf.cpu().numpy().tobytes())


def _run(lib, iterate, problem, state_buf, max_iterations):
    stream = be.stream_ptr()
    for _ in range(max_iterations + 2):          # calls past `done` are empty launches
        be.check(iterate(ctypes.byref(problem), stream), 'iterate')
    return _read(state_buf)


@pytest.mark.parametrize('T,d,k,fused', [(300, 9600, 5, 1), (60, 200, 4, 0), (200, 25600, 12, 1)])
def test_gpnh_fit_through_the_c_entry_points(T, d, k, fused):
    lib = be.library()
    assert lib.cdr_gpnh_fused_applicable(T, d, k) == fused
    X = synthetic_field(T, d, seed=21)
    rs = np.random.RandomState(3)
    W0, Z0 = orc.init_gpnh(X, k, 'random', rs)
    n_it, lam = 6, 0.3
    trace = float(np.sum(X * X))
    ref = orc.iterate_gpnh(X, Z0.copy(), W0.copy(), lambda_W=lam, tolerance=1e-12,
                           max_iterations=n_it, trace_XtX=trace)

    Xd = be.to_device_padded(X)
    ldx, ldt = Xd.stride(0), be.round_up(T)
    Z = be.to_device(Z0)
    WT = be.to_device_padded(np.ascontiguousarray(W0.T))
    XWt = be.zeros(k, ldt)
    kk = [be.zeros(k, k) for _ in range(5)]
    state = _state(1e-12, n_it, trace)
    deltas = be.zeros(n_it)
    ws = torch.empty(lib.cdr_gpnh_workspace_bytes(T, d, k) // 8 + 1, dtype=torch.float64, device='cuda')
    p = be.GpnhProblem(Xd.data_ptr(), ldx, T, d, k, T, lam, Z.data_ptr(), WT.data_ptr(),
                       XWt.data_ptr(), ldt, *[m.data_ptr() for m in kk], state.data_ptr(),
                       deltas.data_ptr(), be.make_spg_params({}), ws.data_ptr(), ws.numel() * 8,
                       None, T)
    be.check(lib.cdr_gpnh_prepare_enqueue(ctypes.byref(p), be.stream_ptr()), 'prepare')
    st = _run(lib, lib.cdr_gpnh_iterate_enqueue, p, state, n_it)
    assert st.done and st.error_stage == 0 and st.n_iter == n_it == ref[3] + 1
    np.testing.assert_allclose(st.cost, ref[2], rtol=1e-8)
    np.testing.assert_allclose(deltas.cpu().numpy(), ref[5], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(Z.cpu().numpy(), ref[0], rtol=0, atol=2e-5)
    np.testing.assert_allclose(be.to_host(WT, k, d).T, ref[1], rtol=0, atol=2e-5)


@pytest.mark.parametrize('T,d,k,inner,fused', [(300, 9600, 5, 1, 1), (300, 9600, 5, 2, 0),
                                                (60, 200, 4, 1, 0), (200, 25600, 12, 1, 1)])
def test_aa_fit_through_the_c_entry_points(T, d, k, inner, fused):
    lib = be.library()
    assert lib.cdr_aa_fused_applicable(T, d, k, inner) == fused
    X = synthetic_field(T, d, seed=22)
    rs = np.random.RandomState(4)
    C0 = orc.right_stochastic_matrix((k, T), rs)
    Z0 = orc.right_stochastic_matrix((T, k), rs)
    n_it = 5
    trace = float(np.sum(X * X))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref = orc.iterate_aa(X, Z0.copy(), C0.copy(), np.ones(k), tolerance=1e-12,
                             max_iterations=n_it, trace_XXt=trace,
                             dictionary_solver_kwargs=dict(max_iterations=inner))

    Xd = be.to_device_padded(X)
    ldx, ldt = Xd.stride(0), be.round_up(T)
    Z = be.to_device(Z0)
    C = be.to_device_padded(C0)
    kt = [be.zeros(k, ldt) for _ in range(5)]              # G, D, CK, DK, KZt
    alpha = be.to_device(np.ones(k))
    kk = [be.zeros(k, k) for _ in range(5)]                # ZtZ, CKCt, CKZ, G01, G11
    scratch = be.zeros(8 * k)
    state = _state(1e-12, n_it, trace)
    deltas = be.zeros(n_it)
    tmp_kd = be.zeros(k, ldx)
    buf = be.AaBuffers(C.data_ptr(), *[m.data_ptr() for m in kt], alpha.data_ptr(),
                       *[m.data_ptr() for m in kk], scratch.data_ptr(), state.data_ptr(),
                       deltas.data_ptr(), k, T, ldt, 1.0 / T, 1.0 / k)
    ws = torch.empty(lib.cdr_aa_workspace_bytes(T, d, k) // 8 + 1, dtype=torch.float64, device='cuda')
    d_params = be.make_spg_params(dict(max_iterations=inner), max_feval=1000000)
    p = be.AaProblem(Xd.data_ptr(), ldx, T, d, buf, Z.data_ptr(), tmp_kd.data_ptr(), d_params,
                     be.make_spg_params({}), ws.data_ptr(), ws.numel() * 8, 0, T, None)
    be.check(lib.cdr_aa_prepare_enqueue(ctypes.byref(p), be.stream_ptr()), 'prepare')
    st = _run(lib, lib.cdr_aa_iterate_enqueue, p, state, n_it)
    assert st.done and st.error_stage == 0 and st.n_iter == n_it == ref[4] + 1
    np.testing.assert_allclose(st.cost, ref[3], rtol=1e-8)
    np.testing.assert_allclose(deltas.cpu().numpy(), ref[6], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(Z.cpu().numpy(), ref[0], rtol=0, atol=2e-5)
    np.testing.assert_allclose(be.to_host(C, k, T), ref[1], rtol=0, atol=2e-5)


def test_kmeans_device_loop_relocates_an_empty_cluster():
    """cdr_kmeans_iterate_enqueue at a shape its device loop covers, with an initial centre so
    far away that its cluster is empty after the first E step: the device stops before the
    centre update, the host relocates (_k_means_common.pyx:167-212) and the loop resumes.
    Labels bit-exact against scikit-learn."""
    from sklearn.cluster import KMeans as SkKMeans
    from convex_dim_red.kmeans import kmeans_lloyd
    T, d, k = 300, 9600, 4
    assert be.library().cdr_kmeans_fused_applicable(T, d, k) == 1
    rs = np.random.RandomState(8)
    centres = rs.standard_normal((3, d)) * 2.0
    X = centres[rs.randint(3, size=T)] + rs.standard_normal((T, d))
    init = np.vstack([X[[0, 1, 2]], 50.0 + np.zeros((1, d))])
    stats = {}
    labels, cent, inertia, n_iter = kmeans_lloyd(X, init, tol=1e-4, max_iter=300, stats=stats)
    sk = SkKMeans(n_clusters=k, init=init, n_init=1, algorithm='lloyd', tol=1e-4, max_iter=300).fit(X)
    assert stats['device_loop']
    assert np.array_equal(labels, sk.labels_) and n_iter == sk.n_iter_
    np.testing.assert_allclose(inertia, sk.inertia_, rtol=1e-10)
    np.testing.assert_allclose(cent, sk.cluster_centers_, rtol=0, atol=1e-9)


def test_kmeans_device_loop_twelve_clusters():
    """The device Lloyd loop for 8 < k <= 16 (16 lanes per sample in the assignment kernel):
    labels, iteration count and centres against scikit-learn from the same initial centres."""
    from sklearn.cluster import KMeans as SkKMeans
    from convex_dim_red.kmeans import kmeans_lloyd
    T, d, k = 600, 25600, 12
    assert be.library().cdr_kmeans_fused_applicable(T, d, k) == 1
    rs = np.random.RandomState(9)
    centres = rs.standard_normal((k, d)) * 1.5
    X = centres[rs.randint(k, size=T)] + rs.standard_normal((T, d))
    init = X[rs.choice(T, size=k, replace=False)].copy()
    stats = {}
    labels, cent, inertia, n_iter = kmeans_lloyd(X, init, tol=1e-4, max_iter=300, stats=stats)
    sk = SkKMeans(n_clusters=k, init=init, n_init=1, algorithm='lloyd', tol=1e-4, max_iter=300).fit(X)
    assert stats['device_loop']
    assert np.array_equal(labels, sk.labels_) and n_iter == sk.n_iter_
    np.testing.assert_allclose(inertia, sk.inertia_, rtol=1e-10)
    np.testing.assert_allclose(cent, sk.cluster_centers_, rtol=0, atol=1e-9)
