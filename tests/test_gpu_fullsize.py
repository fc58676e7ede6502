"""Full-size (BASELINE.json shape) checks.

Two kinds: (i) the CUDA path against the CPU oracle on the same seeded inputs at the
BASELINE.json shapes themselves -- 1620 x 44 000 k = 8 (configs[0] AA, configs[1] GPNH),
700 x 41 800 k = 8 (configs[2]: FurthestSum picks and k-means labels, bit-exact) and the
production width d = 44 000 with k = 64 (configs[4] components, T = 2000 rows) -- the oracle
needs a few seconds for a handful of outer iterations at these sizes; (ii) size-independent
properties -- monotone cost, feasibility of the factors, agreement of the trace-form cost
with the direct residual, linearity of the streaming passes and a checksum of checksums.

Tolerances: iteration counts exact, cost rtol 1e-8, factors atol 2e-5
(the per-sample SPG stops on a 1e-6 projected-gradient norm, spg.py:392), picks / labels
bit-exact."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
if not torch.cuda.is_available():          # pragma: no cover
    pytest.skip('needs a CUDA device', allow_module_level=True)

from convex_dim_red import _backend as be                                   # noqa: E402
from convex_dim_red import archetypal_analysis as aa                        # noqa: E402
from convex_dim_red import gpnh_convex_coding as gp                         # noqa: E402
from convex_dim_red.datasets import synthetic_field                         # noqa: E402
from convex_dim_red.stochastic_matrices import right_stochastic_matrix      # noqa: E402

from oracle import convex_oracle as orc                                     # noqa: E402

T, D, K = 1620, 44000, 8


@pytest.fixture(scope='module')
def hadisst():
    return synthetic_field(T, D, seed=0)


def test_gpnh_full_size_invariants(hadisst):
    X = hadisst
    rs = np.random.RandomState(0)
    W0 = np.sqrt(np.abs(X).mean() / K) * rs.randn(D, K)
    Z0 = right_stochastic_matrix((T, K), random_state=rs)
    Z, W, cost, n_iter, _, deltas = gp._iterate_gpnh_convex_coding(
        X, Z0, W0, lambda_W=0.05, tolerance=0.0, max_iterations=6)
    assert n_iter == 5 and len(deltas) == 6
    assert all(d <= 1e-9 for d in deltas)                       # monotone decrease
    assert np.all(Z >= 0) and np.allclose(Z.sum(axis=1), 1, 1e-12)
    assert W.shape == (D, K)
    # the loop's trace-form cost equals the direct residual cost (gpnh_convex_coding.py:199-210)
    direct = gp._gpnh_cost(X, Z, W, lambda_W=0.05)
    np.testing.assert_allclose(cost, direct, rtol=1e-9)
    # the dictionary is the least-squares solution for the final... previous weights: one more
    # dictionary step cannot increase the cost
    W2 = gp._update_gpnh_dictionary(X, Z, Z.T.dot(Z), 4.0 / (D * K * (K - 1)) * (K * np.eye(K) - 1),
                                    lambda_W=0.05)
    assert gp._gpnh_cost(X, Z, W2, lambda_W=0.05) <= direct * (1 + 1e-12)


def test_aa_full_size_invariants(hadisst):
    X = hadisst
    rs = np.random.RandomState(1)
    C0 = right_stochastic_matrix((K, T), random_state=rs)
    Z0 = right_stochastic_matrix((T, K), random_state=rs)
    Z, C, alpha, cost, n_iter, _, deltas = aa._iterate_aa(
        X, Z0, C0, np.ones(K), tolerance=0.0, max_iterations=5,
        dictionary_solver_kwargs=dict(max_iterations=1))
    assert n_iter == 4
    assert all(d <= 1e-9 for d in deltas)
    assert np.all(Z >= 0) and np.allclose(Z.sum(axis=1), 1, 1e-12)
    assert np.all(C >= 0) and np.allclose(C.sum(axis=1), 1, 1e-12)
    # cost = 1/2 ||X - Z C X||^2 / T, evaluated directly with the residual kernel
    archetypes = C.dot(X)
    direct = gp._gpnh_cost(X, Z, archetypes.T, lambda_W=0)
    np.testing.assert_allclose(cost, direct, rtol=1e-8)


def test_streaming_passes_linearity_and_checksum(hadisst):
    X = hadisst
    rs = np.random.RandomState(2)
    Xd = be.to_device_padded(X)
    ws = be.Workspace(T, D, K)
    L1, L2 = rs.standard_normal((K, T)), rs.standard_normal((K, T))
    outs = []
    for L in (L1, L2, L1 + 2.0 * L2):
        Ld = be.to_device_padded(L)
        o = be.zeros(K, Xd.stride(0))
        be.reduce_samples(Ld, Ld.stride(0), 1, Xd, T, D, K, o, ws)
        outs.append(be.to_host(o, K, D))
    scale = np.abs(outs[2]).max()
    np.testing.assert_allclose(outs[2], outs[0] + 2.0 * outs[1], rtol=0, atol=1e-12 * scale)
    # checksum of checksums with random weights u (the columns of X have zero mean, so plain
    # sums would only compare rounding noise): sum_f u_f (L X)[j, f] = L[j, :] . (X u)
    u = rs.standard_normal(D)
    np.testing.assert_allclose(outs[0].dot(u), L1.dot(X.dot(u)), rtol=1e-9)

    M1, M2 = rs.standard_normal((K, D)), rs.standard_normal((K, D))
    fouts = []
    for M in (M1, M2, M1 - 0.5 * M2):
        Md = be.to_device_padded(M)
        o = be.zeros(K, be.round_up(T))
        be.reduce_features(Md, Xd, T, D, K, o, ws)
        fouts.append(be.to_host(o, K, T))
    scale = np.abs(fouts[2]).max()
    np.testing.assert_allclose(fouts[2], fouts[0] - 0.5 * fouts[1], rtol=0, atol=1e-12 * scale)
    v = rs.standard_normal(T)
    np.testing.assert_allclose(fouts[0].dot(v), M1.dot(v.dot(X)), rtol=1e-9)
    # spot rows against NumPy
    idx = [0, 1, 811, T - 1]
    np.testing.assert_allclose(fouts[0][:, idx], M1.dot(X[idx].T), rtol=1e-11, atol=1e-9)
    cols = [0, 7, 21999, D - 1]
    np.testing.assert_allclose(outs[0][:, cols], L1.dot(X[:, cols]), rtol=1e-11, atol=1e-9)


def test_weights_update_idempotent_at_full_batch(hadisst):
    """Re-solving the per-sample QPs from their own solution leaves it unchanged to solver
    accuracy (idempotence), T = 18 000 samples, k = 8."""
    rs = np.random.RandomState(3)
    n, k = 18000, 8
    Mx = rs.standard_normal((k + 3, k))
    A = Mx.T.dot(Mx)
    B = rs.standard_normal((k, n)) * 3.0
    Z0 = right_stochastic_matrix((n, k), random_state=rs)
    Z1 = aa._update_kernel_aa_weights(Z0, np.ones(k), B, A)
    Z2 = aa._update_kernel_aa_weights(Z1, np.ones(k), B, A)
    assert np.all(Z1 >= 0) and np.allclose(Z1.sum(axis=1), 1, 1e-12)
    np.testing.assert_allclose(Z2, Z1, rtol=0, atol=5e-6)
    obj = lambda Zm: 0.5 * np.einsum('ti,ij,tj->t', Zm, A, Zm) - np.einsum('ti,it->t', Zm, B)
    assert np.all(obj(Z1) <= obj(Z0) + 1e-12)


# ---------------------------------------------------------------------------
# CUDA path vs the CPU oracle at the BASELINE.json shapes
# ---------------------------------------------------------------------------

def _close(a, b, rtol, atol=0.0):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


@pytest.mark.parametrize('lam', [0.0, 0.05])
def test_gpnh_full_size_vs_oracle(hadisst, lam):
    """_iterate_gpnh_convex_coding (gpnh_convex_coding.py:282-402), BASELINE configs[1]."""
    X = hadisst
    rs = np.random.RandomState(10)
    W0, Z0 = orc.init_gpnh(X, K, 'random', rs)
    trace = float(np.sum(X * X))
    ref = orc.iterate_gpnh(X, Z0.copy(), W0.copy(), lambda_W=lam, tolerance=1e-12,
                           max_iterations=5, trace_XtX=trace)
    got = gp._iterate_gpnh_convex_coding(X, Z0.copy(), W0.copy(), lambda_W=lam,
                                         tolerance=1e-12, max_iterations=5)
    assert got[3] == ref[3] == 4
    _close(got[2], ref[2], rtol=1e-8)
    _close(got[5], ref[5], rtol=1e-5, atol=1e-9)
    _close(got[0], ref[0], rtol=0, atol=2e-5)
    _close(got[1], ref[1], rtol=0, atol=2e-5)


def test_aa_full_size_vs_oracle(hadisst):
    """_iterate_aa (archetypal_analysis.py:534-670), BASELINE configs[0], the drivers'
    dictionary_solver_kwargs=dict(max_iterations=1) (bin/run_hadisst_aa.py:160-166)."""
    X = hadisst
    rs = np.random.RandomState(11)
    C0 = right_stochastic_matrix((K, T), random_state=rs)
    Z0 = right_stochastic_matrix((T, K), random_state=rs)
    trace = float(np.sum(X * X))
    kw = dict(tolerance=1e-12, max_iterations=5, dictionary_solver_kwargs=dict(max_iterations=1))
    ref = orc.iterate_aa(X, Z0.copy(), C0.copy(), np.ones(K), trace_XXt=trace, **kw)
    got = aa._iterate_aa(X, Z0.copy(), C0.copy(), np.ones(K), **kw)
    assert got[4] == ref[4] == 4
    _close(got[3], ref[3], rtol=1e-8)
    _close(got[6], ref[6], rtol=1e-5, atol=1e-9)
    _close(got[0], ref[0], rtol=0, atol=2e-5)
    _close(got[1], ref[1], rtol=0, atol=2e-5)


@pytest.mark.parametrize('n_rows', [7000, 14000])
def test_aa_long_dictionary_rows_vs_oracle(n_rows):
    """The eight-kernel iteration with dictionary rows too long for the head kernel's
    all-in-shared-memory layout: 6000 < T <= 13500 keeps only the projected row resident,
    longer rows are re-read from global memory (csrc/iterate_aa.cu).  On one GPU these layouts
    are reached only with many samples; sample-sharded fits reach them through T_total."""
    d, k = 9600, 5
    assert be.library().cdr_aa_fused_applicable(n_rows, d, k, 1) == 1
    X = synthetic_field(n_rows, d, seed=31)
    rs = np.random.RandomState(32)
    C0 = right_stochastic_matrix((k, n_rows), random_state=rs)
    Z0 = right_stochastic_matrix((n_rows, k), random_state=rs)
    trace = float(np.sum(X * X))
    kw = dict(tolerance=1e-12, max_iterations=3, dictionary_solver_kwargs=dict(max_iterations=1))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref = orc.iterate_aa(X, Z0.copy(), C0.copy(), np.ones(k), trace_XXt=trace, **kw)
        got = aa._iterate_aa(X, Z0.copy(), C0.copy(), np.ones(k), **kw)
    assert got[4] == ref[4] == 2
    _close(got[3], ref[3], rtol=1e-8)
    _close(got[6], ref[6], rtol=1e-5, atol=1e-9)
    _close(got[0], ref[0], rtol=0, atol=2e-5)
    _close(got[1], ref[1], rtol=0, atol=2e-5)


def test_aa_full_size_inner_spg_iterations_vs_oracle(hadisst):
    """Two inner SPG iterations per dictionary update exercise the Barzilai-Borwein step and
    the residual test of spg.py:231-281 at full size."""
    X = hadisst
    rs = np.random.RandomState(12)
    C0 = right_stochastic_matrix((K, T), random_state=rs)
    Z0 = right_stochastic_matrix((T, K), random_state=rs)
    trace = float(np.sum(X * X))
    kw = dict(tolerance=1e-12, max_iterations=2, dictionary_solver_kwargs=dict(max_iterations=2))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref = orc.iterate_aa(X, Z0.copy(), C0.copy(), np.ones(K), trace_XXt=trace, **kw)
        got = aa._iterate_aa(X, Z0.copy(), C0.copy(), np.ones(K), **kw)
    assert got[4] == ref[4]
    _close(got[3], ref[3], rtol=1e-8)
    _close(got[1], ref[1], rtol=0, atol=2e-5)
    _close(got[0], ref[0], rtol=0, atol=2e-5)


def test_config3_furthest_sum_and_kmeans_bit_exact():
    """BASELINE configs[2]: k-means k = 8 with FurthestSum initialisation on a 700 x 41 800
    field.  Picks vs the oracle's restatement of furthest_sum.py:23-127, labels vs
    scikit-learn's KMeans itself (the third-party code the drivers call,
    bin/run_hadisst_kmeans.py:128-131) and vs the oracle: all bit-exact."""
    from sklearn.cluster import KMeans as SkKMeans
    from convex_dim_red.furthest_sum import furthest_sum
    from convex_dim_red.kmeans import kmeans_lloyd
    Tk, Dk = 700, 41800
    X = synthetic_field(Tk, Dk, seed=5)
    start = int(np.random.RandomState(0).randint(Tk))
    Kmat = X.dot(X.T)
    Dmat = orc.dissimilarity_from_kernel(Kmat)
    ref_picks = np.asarray(orc.furthest_sum(Dmat, K, start, [], 10))
    # device Gram -> dissimilarity -> selection, through the estimator's initialiser
    C = aa._initialize_kernel_aa_dictionary_furthest_sum(
        aa._LazyKernel(X), K, start_index=start, n_extra_steps=10)
    picks = np.argmax(C, axis=1)
    assert np.array_equal(picks, ref_picks)
    assert np.array_equal(np.asarray(furthest_sum(Dmat, K, start, [], 10)), ref_picks)
    labels, centres, inertia, n_iter = kmeans_lloyd(X, X[picks], tol=1e-4, max_iter=300)
    sk = SkKMeans(n_clusters=K, init=X[ref_picks], n_init=1, algorithm='lloyd', tol=1e-4,
                  max_iter=300).fit(X)
    assert labels.dtype == np.int32
    assert np.array_equal(labels, sk.labels_)
    assert n_iter == sk.n_iter_
    _close(inertia, sk.inertia_, rtol=1e-10)
    _close(centres, sk.cluster_centers_, rtol=0, atol=1e-9)
    olabels, ocentres, oinertia, on_iter = orc.kmeans_lloyd(X, X[ref_picks], tol=1e-4,
                                                            max_iter=300)
    assert np.array_equal(labels, olabels) and n_iter == on_iter


@pytest.fixture(scope='module')
def wide_k64():
    return synthetic_field(2000, D, seed=6)


def test_gpnh_k64_production_width_vs_oracle(wide_k64):
    """k = 64 at d = 44 000 (BASELINE configs[4] components and width, T = 2000 rows): the
    kernels chosen for k > 16 at production width."""
    X = wide_k64
    k = 64
    rs = np.random.RandomState(13)
    W0, Z0 = orc.init_gpnh(X, k, 'random', rs)
    trace = float(np.sum(X * X))
    ref = orc.iterate_gpnh(X, Z0.copy(), W0.copy(), lambda_W=0.0, tolerance=1e-12,
                           max_iterations=3, trace_XtX=trace)
    got = gp._iterate_gpnh_convex_coding(X, Z0.copy(), W0.copy(), lambda_W=0.0,
                                         tolerance=1e-12, max_iterations=3)
    assert got[3] == ref[3]
    _close(got[2], ref[2], rtol=1e-8)
    _close(got[0], ref[0], rtol=0, atol=2e-5)
    _close(got[1], ref[1], rtol=0, atol=2e-5)


def test_aa_k64_production_width_vs_oracle(wide_k64):
    X = wide_k64
    k, Tn = 64, wide_k64.shape[0]
    rs = np.random.RandomState(14)
    C0 = right_stochastic_matrix((k, Tn), random_state=rs)
    Z0 = right_stochastic_matrix((Tn, k), random_state=rs)
    trace = float(np.sum(X * X))
    kw = dict(tolerance=1e-12, max_iterations=3, dictionary_solver_kwargs=dict(max_iterations=1))
    ref = orc.iterate_aa(X, Z0.copy(), C0.copy(), np.ones(k), trace_XXt=trace, **kw)
    got = aa._iterate_aa(X, Z0.copy(), C0.copy(), np.ones(k), **kw)
    assert got[4] == ref[4]
    _close(got[3], ref[3], rtol=1e-8)
    _close(got[0], ref[0], rtol=0, atol=2e-5)
    _close(got[1], ref[1], rtol=0, atol=2e-5)
