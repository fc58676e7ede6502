"""Generate golden input/output vectors by running the REAL reference.

Run in the build container only (the reference is mounted at /root/reference,
which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports ``convex_dim_red`` from ``/root/reference/src`` unmodified (after the
NumPy-2 shim ``np.NaN = np.nan`` that ``spg.py:310`` needs) and scikit-learn
1.9.0 for the k-means cases, runs small seeded problems through the reference
functions on the hot path and stores inputs and outputs in
``tests/golden/golden_v1.npz``.  The fixtures pin ``oracle/`` (CPU tests) and
the CUDA path (GPU tests).
"""

import os
import sys
import warnings

import numpy as np

np.NaN = np.nan          # noqa: NumPy-2 shim for reference spg.py:310
np.product = np.prod

REF_SRC = os.environ.get('CDR_REFERENCE_SRC', '/root/reference/src')
sys.path.insert(0, REF_SRC)

warnings.filterwarnings('ignore')

import convex_dim_red as ref                                     # noqa: E402
from convex_dim_red import archetypal_analysis as raa            # noqa: E402
from convex_dim_red import gpnh_convex_coding as rgp             # noqa: E402
from convex_dim_red import simplex_projection as rsp             # noqa: E402
rspg = sys.modules['convex_dim_red.spg']    # the package re-exports the function `spg`
from sklearn.cluster import KMeans                               # noqa: E402
import sklearn                                                   # noqa: E402

OUT = {}


def put(name, value):
    OUT[name] = np.asarray(value)


def synth(T, d, k0, seed, sigma=0.5):
    """Structured synthetic anomalies: X = Z0 A0 + sigma E, column mean removed
    (SURVEY.md section 8d)."""
    rs = np.random.RandomState(seed)
    Z0 = ref.right_stochastic_matrix((T, k0), random_state=rs)
    A0 = rs.standard_normal((k0, d))
    E = rs.standard_normal((T, d))
    X = Z0.dot(A0) + sigma * E
    return X - X.mean(axis=0)


# ---------------------------------------------------------------- simplex
def gen_simplex():
    rs = np.random.RandomState(11)
    for n in (1, 2, 3, 5, 8, 17, 64, 317):
        x = rs.uniform(-2, 2, size=n)
        put('simplex/vec%d/x' % n, x)
        put('simplex/vec%d/out' % n, rsp.simplex_project_vector(x))
    # a vector already on the simplex and one with ties / zeros
    x = np.array([0.25, 0.25, 0.0, 0.5, 0.0])
    put('simplex/feasible/x', x)
    put('simplex/feasible/out', rsp.simplex_project_vector(x))
    x = np.array([1.0, 1.0, 1.0, -3.0, 1.0, 0.2])
    put('simplex/ties/x', x)
    put('simplex/ties/out', rsp.simplex_project_vector(x))
    A = rs.uniform(-1, 1, size=(7, 13))
    put('simplex/rows/A', A)
    put('simplex/rows/out', rsp.simplex_project_rows(A))
    put('simplex/cols/out', rsp.simplex_project_columns(A))
    A = rs.standard_normal((9, 700)) * 0.01 + 1.0 / 700
    put('simplex/rows_wide/A', A)
    put('simplex/rows_wide/out', rsp.simplex_project_rows(A))


# ---------------------------------------------------------------- QP
def gen_qp():
    rs = np.random.RandomState(5)
    for k in (2, 3, 8, 16):
        M = rs.standard_normal((k + 3, k))
        A = M.T.dot(M)
        b = rs.standard_normal(k)
        x0 = rs.uniform(size=k)
        x0 /= x0.sum()
        put('qp/k%d/A' % k, A)
        put('qp/k%d/b' % k, b)
        put('qp/k%d/x0' % k, x0)
        put('qp/k%d/x' % k, rspg.quad_simplex_spg(A, b, x0))
        put('qp/k%d/x_it3' % k, rspg.quad_simplex_spg(A, b, x0, max_iterations=3))
    # batched: AA-style (k x T linear term) and GPNH-style (T x k)
    T, k = 60, 5
    M = rs.standard_normal((k + 2, k))
    A = M.T.dot(M)
    CK = rs.standard_normal((k, T))
    Z0 = ref.right_stochastic_matrix((T, k), random_state=rs)
    alpha = np.ones(k)
    put('qp/batch/A', A)
    put('qp/batch/CK', CK)
    put('qp/batch/Z0', Z0)
    put('qp/batch/Z_aa', raa._update_kernel_aa_weights(Z0, alpha, CK, A))
    put('qp/batch/Z_aa_it1',
        raa._update_kernel_aa_weights(Z0, alpha, CK, A, max_iterations=1))
    alpha2 = rs.uniform(0.8, 1.2, size=k)
    put('qp/batch/alpha2', alpha2)
    put('qp/batch/Z_aa_alpha2', raa._update_kernel_aa_weights(Z0, alpha2, CK, A))


# ---------------------------------------------------------------- spg (generic)
def gen_spg():
    # the two known-answer problems of the reference's tests/test_spg.py
    def f(x):
        return x ** 4 + 2 * x ** 2 + 1

    def df(x):
        return 4 * x ** 3 + 4 * x

    def project(x):
        return min(max(x, -1.0), 0.5)

    sol = rspg.spg(f, df, 0.4, project=project)
    put('spg/quartic/out', np.array(sol, dtype=np.float64))
    # vector problem: least squares on a box
    rs = np.random.RandomState(3)
    M = rs.standard_normal((12, 6))
    y = rs.standard_normal(12)
    x0 = rs.uniform(size=6)

    def f2(x):
        r = M.dot(x) - y
        return 0.5 * r.dot(r)

    def df2(x):
        return M.T.dot(M.dot(x) - y)

    def p2(x):
        return np.fmin(np.fmax(x, 0.0), 0.3)

    x, fx, n_it, n_fe = rspg.spg(f2, df2, x0, project=p2)
    put('spg/box/M', M)
    put('spg/box/y', y)
    put('spg/box/x0', x0)
    put('spg/box/x', x)
    put('spg/box/stats', np.array([fx, n_it, n_fe], dtype=np.float64))


# ---------------------------------------------------------------- furthest sum
def gen_furthest_sum():
    rs = np.random.RandomState(7)
    P = rs.standard_normal((40, 6))
    K = P.dot(P.T)
    kd = np.diag(K)
    D = np.sqrt(np.tile(kd, (40, 1)) - 2 * K + np.tile(kd[:, None], (1, 40)))
    D = np.nan_to_num(D)
    put('fs/D', D)
    put('fs/P', P)
    cases = [(5, 3, [], 10), (5, 3, [], 0), (8, 0, [4, 9], 10),
             (1, 17, [], 3), (6, 39, [0, 1, 2], 1), (40, 2, [], 2)]
    put('fs/cases', np.array([[c[0], c[1], c[3]] for c in cases]))
    for i, (k, start, excl, extra) in enumerate(cases):
        put('fs/case%d/exclude' % i, np.array(excl, dtype=np.int64))
        put('fs/case%d/selected' % i,
            np.asarray(ref.furthest_sum(D, k, start, excl, extra), dtype=np.int64))
    # known-answer tests of tests/test_furthest_sum.py:142-194
    K3 = np.array([[0, 1, 2.0], [1, 0, 0.5], [2, 0.5, 0]])
    put('fs/k3/D', K3)
    put('fs/k3/sel', np.asarray(ref.furthest_sum(K3, 2, 1, [], 3), dtype=np.int64))


# ---------------------------------------------------------------- kernel AA / AA
def gen_aa():
    T, d, k = 48, 30, 4
    X = synth(T, d, 5, seed=21)
    K = X.dot(X.T)
    rs = np.random.RandomState(2)
    C0 = ref.right_stochastic_matrix((k, T), random_state=rs)
    Z0 = ref.right_stochastic_matrix((T, k), random_state=rs)
    alpha = np.ones(k)
    put('aa/X', X)
    put('aa/C0', C0)
    put('aa/Z0', Z0)
    put('aa/kernel_cost', raa._kernel_aa_cost(K, Z0, C0, alpha))

    ZtZ = Z0.T.dot(Z0)
    KZ = K.dot(Z0)
    trace_K = K.trace()
    # single-function pins
    XXtZD = KZ.copy()
    put('aa/dict_cost', raa._aa_dictionary_cost(X, C0, trace_K, XXtZD, ZtZ))
    put('aa/dict_grad', raa._aa_dictionary_gradient(X, C0, XXtZD, ZtZ))
    put('aa/kdict_cost', raa._kernel_aa_dictionary_cost(K, C0, trace_K, KZ, ZtZ))
    put('aa/kdict_grad', raa._kernel_aa_dictionary_gradient(K, C0, KZ, ZtZ))
    for it in (1, 2, 5):
        put('aa/kdict_update_it%d' % it, raa._update_kernel_aa_dictionary(
            K, C0, alpha, trace_K, KZ, ZtZ, max_iterations=it))
        put('aa/dict_update_it%d' % it, raa._update_aa_dictionary(
            X, C0, alpha, trace_K, KZ, ZtZ, max_iterations=it))
    CK = C0.dot(K)
    CKCt = CK.dot(C0.T)
    put('aa/weights_update', raa._update_kernel_aa_weights(Z0, alpha, CK, CKCt))

    for name, fn, data in (('kaa', raa._iterate_kernel_aa, K),
                           ('faa', raa._iterate_aa, X)):
        for tag, kw in (('d1', dict(dictionary_solver_kwargs=dict(max_iterations=1))),
                        ('d3w5', dict(dictionary_solver_kwargs=dict(max_iterations=3),
                                      weights_solver_kwargs=dict(max_iterations=5))),
                        ('rel', dict(dictionary_solver_kwargs=dict(max_iterations=1),
                                     weights_solver_kwargs=dict(max_iterations=1),
                                     stopping_criterion='rel_delta_f'))):
            Z, C, a, cost, n_iter, _, deltas = fn(
                data, Z0.copy(), C0.copy(), alpha.copy(), tolerance=1e-9,
                max_iterations=12, **kw)
            put('aa/%s_%s/Z' % (name, tag), Z)
            put('aa/%s_%s/C' % (name, tag), C)
            put('aa/%s_%s/stats' % (name, tag), np.array([cost, n_iter]))
            put('aa/%s_%s/deltas' % (name, tag), np.array(deltas))
    # frozen-factor runs
    Z, C, a, cost, n_iter, _, deltas = raa._iterate_kernel_aa(
        K, Z0.copy(), C0.copy(), alpha.copy(), update_dictionary=False,
        tolerance=1e-9, max_iterations=5)
    put('aa/kaa_frozenC/Z', Z)
    put('aa/kaa_frozenC/stats', np.array([cost, n_iter]))
    # delta != 0 (scale factors), few iterations
    Z, C, a, cost, n_iter, _, deltas = raa._iterate_kernel_aa(
        K, Z0.copy(), C0.copy(), alpha.copy(), delta=0.2, tolerance=1e-9,
        max_iterations=6, dictionary_solver_kwargs=dict(max_iterations=1))
    put('aa/kaa_delta/Z', Z)
    put('aa/kaa_delta/C', C)
    put('aa/kaa_delta/alpha', a)
    put('aa/kaa_delta/stats', np.array([cost, n_iter]))

    # estimator classes (random and furthest-sum init), to convergence-ish
    for init in ('random', 'furthest_sum'):
        m = ref.ArchetypalAnalysis(
            n_components=k, init=init, tolerance=1e-6, max_iterations=40,
            random_state=0, dictionary_solver_kwargs=dict(max_iterations=1))
        Z = m.fit_transform(X)
        put('aa/est_%s/Z' % init, Z)
        put('aa/est_%s/C' % init, m.dictionary)
        put('aa/est_%s/archetypes' % init, m.archetypes)
        put('aa/est_%s/stats' % init, np.array([m.cost, m.n_iter]))
        put('aa/est_%s/deltas' % init, np.array(m.cost_deltas))
        Xv = synth(12, d, 5, seed=22)
        Zv, cv = m.transform(Xv)
        put('aa/est_%s/Xv' % init, Xv)
        put('aa/est_%s/Zv' % init, Zv)
        put('aa/est_%s/cost_v' % init, cv)
        mk = ref.KernelAA(
            n_components=k, init=init, tolerance=1e-6, max_iterations=40,
            random_state=0, dictionary_solver_kwargs=dict(max_iterations=2))
        Zk = mk.fit_transform(K)
        put('aa/kest_%s/Z' % init, Zk)
        put('aa/kest_%s/C' % init, mk.dictionary)
        put('aa/kest_%s/stats' % init, np.array([mk.cost, mk.n_iter]))

    # triangle vertex recovery (reference tests/test_archetypal_analysis.py:496-606
    # style): index-exact pin
    rs = np.random.RandomState(0)
    basis = np.array([[0.0, 0.0], [1.0, 0.0], [0.5, 1.0]])
    n = 40
    W = ref.right_stochastic_matrix((n, 3), random_state=rs)
    vertex_idx = np.array([5, 27, 32])
    for i, v in enumerate(vertex_idx):
        W[v] = 0
        W[v, i] = 1
    Xtri = W.dot(basis)
    Ktri = Xtri.dot(Xtri.T)
    mk = ref.KernelAA(n_components=3, init='furthest_sum', tolerance=1e-8,
                      max_iterations=200, random_state=0,
                      dictionary_solver_kwargs=dict(max_iterations=20))
    mk.fit_transform(Ktri)
    put('aa/tri/X', Xtri)
    put('aa/tri/C', mk.dictionary)
    put('aa/tri/picks', np.sort(np.argmax(mk.dictionary, axis=1)))
    put('aa/tri/stats', np.array([mk.cost, mk.n_iter]))


# ---------------------------------------------------------------- GPNH
def gen_gpnh():
    T, d, k = 50, 24, 4
    X = synth(T, d, 5, seed=31)
    rs = np.random.RandomState(4)
    W0 = np.sqrt(np.abs(X).mean() / k) * rs.randn(d, k)
    Z0 = ref.right_stochastic_matrix((T, k), random_state=rs)
    put('gpnh/X', X)
    put('gpnh/W0', W0)
    put('gpnh/Z0', Z0)
    put('gpnh/reg', rgp._gpnh_regularization(W0))
    put('gpnh/cost0', rgp._gpnh_cost(X, Z0, W0, 0.0))
    put('gpnh/cost_l', rgp._gpnh_cost(X, Z0, W0, 3.2))
    ZtZ = Z0.T.dot(Z0)
    GW = 4.0 / (d * k * (k - 1)) * (k * np.eye(k) - 1)
    for lam in (0.0, 3.2):
        tag = 'l%g' % lam
        put('gpnh/dict_update_%s' % tag,
            np.ascontiguousarray(rgp._update_gpnh_dictionary(X, Z0, ZtZ, GW, lambda_W=lam)))
        put('gpnh/weights_update', rgp._update_gpnh_weights(X, Z0, W0))
        put('gpnh/weights_update_it2',
            rgp._update_gpnh_weights(X, Z0, W0, max_iterations=2))
        for wtag, kw in (('full', {}), ('w1', dict(weights_solver_kwargs=dict(max_iterations=1)))):
            Z, W, cost, n_iter, _, deltas = rgp._iterate_gpnh_convex_coding(
                X, Z0.copy(), W0.copy(), lambda_W=lam, tolerance=1e-9,
                max_iterations=10, **kw)
            put('gpnh/it_%s_%s/Z' % (tag, wtag), Z)
            put('gpnh/it_%s_%s/W' % (tag, wtag), np.ascontiguousarray(W))
            put('gpnh/it_%s_%s/stats' % (tag, wtag), np.array([cost, n_iter]))
            put('gpnh/it_%s_%s/deltas' % (tag, wtag), np.array(deltas))
    # frozen dictionary
    Z, W, cost, n_iter, _, deltas = rgp._iterate_gpnh_convex_coding(
        X, Z0.copy(), W0.copy(), update_dictionary=False, tolerance=1e-9,
        max_iterations=4)
    put('gpnh/frozenW/Z', Z)
    put('gpnh/frozenW/stats', np.array([cost, n_iter]))

    for init in ('random', 'furthest_sum'):
        m = ref.GPNHConvexCoding(n_components=k, lambda_W=0.5, init=init,
                                 tolerance=1e-6, max_iterations=30, random_state=0)
        Z = m.fit_transform(X)
        put('gpnh/est_%s/Z' % init, Z)
        put('gpnh/est_%s/W' % init, np.ascontiguousarray(m.dictionary))
        put('gpnh/est_%s/stats' % init, np.array([m.cost, m.n_iter]))
        put('gpnh/est_%s/deltas' % init, np.array(m.cost_deltas))
        Xv = synth(10, d, 5, seed=32)
        Zv, cv = m.transform(Xv)
        put('gpnh/est_%s/Xv' % init, Xv)
        put('gpnh/est_%s/Zv' % init, Zv)
        put('gpnh/est_%s/cost_v' % init, cv)


# ---------------------------------------------------------------- k-means
def gen_kmeans():
    T, d, k = 120, 20, 5
    X = synth(T, d, 6, seed=41, sigma=0.3)
    K = X.dot(X.T)
    kd = np.diag(K)
    D = np.sqrt(np.maximum(
        np.tile(kd, (T, 1)) - 2 * K + np.tile(kd[:, None], (1, T)), 0))
    picks = np.asarray(ref.furthest_sum(D, k, 7, [], 10), dtype=np.int64)
    put('km/X', X)
    put('km/picks', picks)
    for tag, tol, max_iter in (('conv', 1e-4, 300), ('it2', 1e-4, 2), ('tol0', 0.0, 300)):
        km = KMeans(n_clusters=k, init=X[picks].copy(), n_init=1, algorithm='lloyd',
                    tol=tol, max_iter=max_iter).fit(X.copy())
        put('km/%s/labels' % tag, km.labels_.astype(np.int32))
        put('km/%s/centres' % tag, km.cluster_centers_)
        put('km/%s/stats' % tag, np.array([km.inertia_, km.n_iter_]))
    # an init that produces an empty cluster (relocation path)
    init = X[picks].copy()
    init[2] = X.max(axis=0) * 50.0
    km = KMeans(n_clusters=k, init=init, n_init=1, algorithm='lloyd',
                tol=1e-4, max_iter=300).fit(X.copy())
    put('km/empty/init', init)
    put('km/empty/labels', km.labels_.astype(np.int32))
    put('km/empty/centres', km.cluster_centers_)
    put('km/empty/stats', np.array([km.inertia_, km.n_iter_]))


def main():
    gen_simplex()
    gen_qp()
    gen_spg()
    gen_furthest_sum()
    gen_aa()
    gen_gpnh()
    gen_kmeans()
    put('meta/versions', np.array([np.__version__, sklearn.__version__]))
    here = os.path.dirname(os.path.abspath(__file__))
    path = os.path.join(here, 'golden_v1.npz')
    np.savez_compressed(path, **OUT)
    print('wrote %s: %d arrays, %.1f KiB' % (path, len(OUT), os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
