"""Worker for the multi-process tests (launched through torch.distributed.run).

--mode gloo : CPU tensors; exercises convex_dim_red._dist (shard bounds, all-reduce,
              ragged column all-gather) and the sample-sharded algebra of one GPNH
              dictionary step against the single-process oracle.
--mode nccl : one GPU per rank; runs the sample-sharded GPNH and AA engines and saves
              rank 0's gathered result for comparison with a single-GPU run.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'matrix-factorization-case-studies_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch                      # noqa: E402
import torch.distributed as dist  # noqa: E402


def problem(T=203, d=300, k=5, seed=11, path=None):
    """Seeded test problem.  ``path``: load it from / save it to an .npz, so that the worker
    processes and the test process use bit-identical inputs (the BLAS product inside the
    generator rounds differently under a different thread count)."""
    if path is not None and os.path.exists(path):
        f = np.load(path)
        return f['X'], f['Z0'], f['W0'], f['C0']
    made = _generate_problem(T, d, k, seed)
    if path is not None:
        np.savez(path, X=made[0], Z0=made[1], W0=made[2], C0=made[3])
    return made


def _generate_problem(T, d, k, seed):
    from oracle import convex_oracle as orc
    from convex_dim_red.datasets import synthetic_field
    X = synthetic_field(T, d, seed=seed)
    rs = np.random.RandomState(seed)
    W0 = 0.3 * rs.randn(d, k)
    Z0 = orc.right_stochastic_matrix((T, k), rs)
    C0 = orc.right_stochastic_matrix((k, T), rs)
    return X, Z0, W0, C0


DATA = [None]      # path of the shared problem file (--data), if any


class _FakeModel:
    """Stands in for an estimator: its cost is a function of the draws its initialisation
    takes from the shared RNG, so a wrong draw order changes the winner."""

    def __init__(self, rng, shape):
        C = rng.uniform(size=shape)
        Z = rng.uniform(size=shape[::-1])
        self.cost = float(np.round(np.abs(C.sum() - Z.sum()), 1))    # rounded: forces ties
        self.draws = (C, Z)


def replicated_restarts_check(comm):
    """best_of_restarts over the group returns, on every rank, exactly the model the serial
    loop keeps (first minimum, shared RNG sequence), for several restart counts."""
    from convex_dim_red.model_selection import best_of_restarts, _skip_aa_initialisation
    shape = (3, 17)
    for n_init in (1, 2, 5, 8):
        rng = np.random.RandomState(5)
        serial = best_of_restarts(n_init, lambda: _FakeModel(rng, shape))
        serial_state = rng.get_state()[1].copy()
        rng = np.random.RandomState(5)
        got = best_of_restarts(
            n_init, lambda: _FakeModel(rng, shape),
            lambda: _skip_aa_initialisation(rng, 'random', shape[1], shape[0], 0), comm)
        assert got.cost == serial.cost
        assert np.array_equal(got.draws[0], serial.draws[0])
        assert np.array_equal(got.draws[1], serial.draws[1])
        assert np.array_equal(rng.get_state()[1], serial_state)


def row_helpers_check(comm, X):
    """Helpers of the row-sharded k-means / FurthestSum / estimator paths on CPU tensors."""
    from convex_dim_red._dist import shard_bounds, shard_sizes
    T, d = X.shape
    rank, world = comm.rank, comm.world
    lo, hi = shard_bounds(T, world, rank)
    sizes = shard_sizes(T, world)
    assert comm.device() == 'cpu'
    assert comm.local_rows(hi - lo) == (lo, T)
    try:
        comm.local_rows(hi - lo + (1 if rank == 0 else -1))     # total unchanged, split wrong
    except ValueError:
        pass
    else:
        raise AssertionError('unbalanced row blocks must be rejected')
    assert comm.sum_scalars([1.0, float(rank)]) == [float(world), 1.0]

    # device-side all-gather of row blocks (ragged: 203 rows over 2 ranks)
    Xl = torch.from_numpy(X[lo:hi].copy())
    full = comm.allgather_row_blocks(Xl, sizes)
    assert full.shape == (T, d) and np.array_equal(full.numpy(), X)
    even = comm.allgather_row_blocks(Xl[:100].contiguous(), [100, 100])
    assert np.array_equal(even.numpy(), np.concatenate([X[:100], X[102:202]]))

    # column moments merged from per-rank moments (k-means centring and tolerance)
    Y = X + 50.0                                            # large mean: no cancellation allowed
    Yl = Y[lo:hi]
    mean, var = comm.merge_column_moments(torch.from_numpy(Yl.mean(axis=0)),
                                          torch.from_numpy(Yl.var(axis=0)), hi - lo, T)
    np.testing.assert_allclose(mean.numpy(), Y.mean(axis=0), rtol=1e-14)
    np.testing.assert_allclose(var.numpy(), Y.var(axis=0), rtol=1e-11)

    # globally farthest samples (empty-cluster relocation): values with ties across ranks
    vals = np.round(np.abs(X[:, 0]) * 3.0)
    got = comm.global_top(vals[lo:hi], lo, 6)
    order = sorted(range(T), key=lambda i: (-vals[i], i))[:6]
    assert [g for _, g in got] == order and [v for v, _ in got] == [vals[i] for i in order]
    assert comm.global_top(vals[lo:hi], lo, 0) == []
    assert len(comm.global_top(vals[lo:lo + 1], lo, 5)) == world      # fewer rows than asked

    t = torch.full((3,), float(rank))
    comm.broadcast(t, 1)
    assert t.tolist() == [1.0, 1.0, 1.0]
    assert comm.broadcast_object({'rank': rank}, 1) == {'rank': 1}
    assert comm.allgather_objects(rank) == list(range(world))


def run_gloo(out):
    from oracle import convex_oracle as orc
    from convex_dim_red._dist import Comm, shard_bounds, shard_sizes
    dist.init_process_group('gloo')
    comm = Comm()
    rank, world = comm.rank, comm.world
    assert world == 2 and comm.enabled
    X, Z0, W0, C0 = problem()
    T, d = X.shape
    k = Z0.shape[1]
    lo, hi = shard_bounds(T, world, rank)
    sizes = shard_sizes(T, world)
    assert sum(sizes) == T and sizes[rank] == hi - lo and max(sizes) - min(sizes) <= 1
    Xl, Zl = X[lo:hi], Z0[lo:hi]

    # statistics that reduce over samples: local partial + sum all-reduce
    stats = torch.from_numpy(np.stack([Zl.T.dot(Zl), W0.T.dot(Xl.T).dot(Zl)]))
    comm.allreduce_sum(stats)
    ZtZ = stats[0].numpy()
    np.testing.assert_allclose(ZtZ, Z0.T.dot(Z0), rtol=1e-12)
    np.testing.assert_allclose(stats[1].numpy(), W0.T.dot(X.T).dot(Z0), rtol=1e-11, atol=1e-9)

    # dictionary step: W' = P sum_g Z_g' X_g with P = pinv(Z'Z / T + lambda G) / T
    lam = 0.7
    P = np.linalg.pinv(ZtZ / T + lam * orc.gpnh_GW(d, k)) / T
    WT = torch.from_numpy(P.dot(Zl.T.dot(Xl)))
    comm.allreduce_sum(WT)
    ref = orc.update_gpnh_dictionary(X, Z0, Z0.T.dot(Z0), orc.gpnh_GW(d, k), lambda_W=lam)
    np.testing.assert_allclose(WT.numpy().T, ref, rtol=1e-9, atol=1e-11)

    # products that reduce over features are local; their columns are all-gathered
    local = torch.from_numpy(np.ascontiguousarray(C0.dot(X).dot(Xl.T)))       # k x T_local
    padded = torch.zeros((k, local.shape[1] + 5), dtype=torch.float64)
    padded[:, :local.shape[1]] = local
    full = torch.zeros((k, T + 3), dtype=torch.float64)
    comm.allgather_columns(padded, full, sizes)
    np.testing.assert_allclose(full[:, :T].numpy(), C0.dot(X).dot(X.T), rtol=1e-12, atol=1e-9)
    assert float(full[:, T:].abs().sum()) == 0.0

    # equal shards take the single strided-copy path
    eq = torch.arange(k * 7, dtype=torch.float64).reshape(k, 7) + 100.0 * rank
    full_eq = torch.zeros((k, 20), dtype=torch.float64)
    comm.allgather_columns(eq, full_eq, [7, 7])
    want = torch.cat([torch.arange(k * 7, dtype=torch.float64).reshape(k, 7) + 100.0 * r
                      for r in range(world)], dim=1)
    assert torch.equal(full_eq[:, :14], want) and float(full_eq[:, 14:].abs().sum()) == 0.0

    rows = comm.allgather_rows(Zl)
    assert np.array_equal(rows, Z0)
    # ranks seeded differently (random_state=None in the estimators): after the sync every
    # rank draws what rank 0 would have drawn
    rng = np.random.RandomState(100 + rank)
    comm.sync_random_state(rng)
    draws = comm.allgather_objects(rng.uniform(size=4).tolist())
    assert draws[0] == draws[1] == np.random.RandomState(100).uniform(size=4).tolist()
    replicated_restarts_check(comm)
    row_helpers_check(comm, X)
    mx = torch.tensor([float(rank)], dtype=torch.float64)
    comm.allreduce_max(mx)
    assert mx.item() == world - 1
    if rank == 0:
        np.savez(out, ok=np.array([1]))
    dist.barrier()
    dist.destroy_process_group()


def wider_cases(X, lo, hi, comm=None):
    """Estimator-level fits, replicated restarts and k-means; run row-sharded over the group
    by the worker and on one GPU by the test (comm=None), which compares the two."""
    import warnings
    from convex_dim_red import ArchetypalAnalysis, GPNHConvexCoding
    from convex_dim_red.kmeans import KMeans
    from convex_dim_red.model_selection import fit_aa_model, fit_gpnh_model
    warnings.simplefilter('ignore', UserWarning)
    kw = dict(comm=comm) if comm is not None else {}
    Xl = np.ascontiguousarray(X[lo:hi])
    res = {}

    m = ArchetypalAnalysis(5, init='furthest_sum', random_state=3, max_iterations=5,
                           tolerance=1e-12, dictionary_solver_kwargs=dict(max_iterations=1))
    m.fit_transform(Xl, **kw)
    res.update(e_aa_Z=m.weights, e_aa_C=m.dictionary, e_aa_A=m.archetypes, e_aa_cost=m.cost)
    m = GPNHConvexCoding(4, lambda_W=0.1, init='random', random_state=4, max_iterations=5,
                         tolerance=1e-12)
    m.fit_transform(Xl, **kw)
    res.update(e_gp_Z=m.weights, e_gp_W=m.dictionary, e_gp_cost=m.cost)
    m = GPNHConvexCoding(4, init='furthest_sum', random_state=5, max_iterations=3,
                         tolerance=1e-12)
    m.fit_transform(Xl, **kw)
    res.update(e_gf_Z=m.weights, e_gf_W=m.dictionary, e_gf_cost=m.cost)

    # replicated restarts work on the whole matrix on every rank
    m = fit_gpnh_model(X, n_components=4, lambda_W=0.05, n_init=3, max_iterations=4,
                       tolerance=1e-12, random_state=7, **kw)
    res.update(r_gp_Z=m.weights, r_gp_W=m.dictionary, r_gp_cost=m.cost)
    m = fit_aa_model(X, n_components=4, init='furthest_sum', n_init=3, max_iterations=4,
                     tolerance=1e-12, random_state=8, **kw)
    res.update(r_aa_Z=m.weights, r_aa_C=m.dictionary, r_aa_cost=m.cost)

    for name, init in (('fs', 'furthest_sum'), ('rnd', 'random')):
        km = KMeans(n_clusters=5, init=init, random_state=1, max_iter=50).fit(Xl, **kw)
        res.update({'k_%s_labels' % name: km.labels_, 'k_%s_centres' % name: km.cluster_centers_,
                    'k_%s_inertia' % name: km.inertia_, 'k_%s_n' % name: km.n_iter_})
    # a centre far from every sample starts empty and is moved onto the farthest sample
    init = np.concatenate([X[[0, 57, 111, 180]], np.full((1, X.shape[1]), 1e3)])
    km = KMeans(n_clusters=5, init=init, max_iter=50).fit(Xl, **kw)
    res.update(k_emp_labels=km.labels_, k_emp_centres=km.cluster_centers_,
               k_emp_inertia=km.inertia_, k_emp_n=km.n_iter_)
    return res


def wider_nccl_cases(comm, X, lo, hi):
    return wider_cases(X, lo, hi, comm)


def run_nccl(out):
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from convex_dim_red._dist import Comm, shard_bounds
    from convex_dim_red import archetypal_analysis as aa
    from convex_dim_red import gpnh_convex_coding as gp
    comm = Comm()
    X, Z0, W0, C0 = problem(T=403, d=2600, k=8, path=DATA[0])
    T = X.shape[0]
    lo, hi = shard_bounds(T, comm.world, comm.rank)
    g = gp._iterate_gpnh_convex_coding(X[lo:hi].copy(), Z0[lo:hi].copy(), W0.copy(), lambda_W=0.2,
                                       tolerance=1e-12, max_iterations=6, comm=comm)
    a = aa._iterate_aa(X[lo:hi].copy(), Z0[lo:hi].copy(), C0.copy(), np.ones(C0.shape[0]),
                       tolerance=1e-12, max_iterations=6,
                       dictionary_solver_kwargs=dict(max_iterations=2), comm=comm)
    Zg = comm.allgather_rows(g[0])
    Za = comm.allgather_rows(a[0])
    extra = wider_nccl_cases(comm, X, lo, hi)
    if comm.rank == 0:
        np.savez(out, gZ=Zg, gW=g[1], gcost=g[2], gn=g[3], aZ=Za, aC=a[1], acost=a[3], an=a[4],
                 **extra)
    dist.barrier()
    dist.destroy_process_group()


def run_peer(out):
    """Peer-memory collectives (CDR_PEER_COLLECTIVES=1) against NCCL / the unfused kernels."""
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    os.environ['CDR_PEER_COLLECTIVES'] = '1'
    from convex_dim_red import _backend as be
    from convex_dim_red._dist import Comm, shard_bounds, shard_sizes
    from convex_dim_red import archetypal_analysis as aa
    from convex_dim_red import gpnh_convex_coding as gp
    comm = Comm()
    rank, world = comm.rank, comm.world
    report = {}

    # --- stand-alone collectives on tensors of several sizes (odd, tiny, multi-chunk, 2.8 MB)
    T, d, k = 1620 // world, 44000, 8
    ldx = be.round_up(d)
    shapes = [(3, 5, 5), (2,), (k, k), (4100,), (k, ldx), (k, be.round_up(203))]
    peer = comm.setup_peer(shapes + [(k, ldx)], (k, ldx))
    assert peer is not None
    gen = torch.Generator(device='cuda').manual_seed(100 + rank)
    tensors = [comm.zeros(*shape) for shape in shapes[:5]]
    for rep in range(3):                                   # several epochs of every flag
        for i, t in enumerate(tensors):
            t.copy_(torch.randn(t.shape, generator=gen, device='cuda', dtype=torch.float64))
            want = t.clone()
            dist.all_reduce(want)
            assert peer.can_allreduce(t)
            comm.allreduce_sum(t)
            torch.cuda.synchronize()
            exact = bool(torch.equal(t, want))
            close = bool(torch.allclose(t, want, rtol=1e-13, atol=1e-13))
            report['allreduce_%d_%d' % (rep, i)] = (exact if world == 2 else close)
    # views: an even-length prefix goes through the peer kernel, an odd-length middle view
    # falls back to NCCL -- both must give the right sum
    stats = tensors[0]
    stats.copy_(torch.randn(stats.shape, generator=gen, device='cuda', dtype=torch.float64))
    want = stats.clone()
    dist.all_reduce(want)
    assert peer.can_allreduce(stats[:2]) and not peer.can_allreduce(stats[1])
    comm.allreduce_sum(stats[:2])
    comm.allreduce_sum(stats[2])
    torch.cuda.synchronize()
    report['allreduce_views'] = bool(torch.allclose(stats, want, rtol=1e-13, atol=1e-13))

    # --- all-gather of ragged column blocks
    sizes = shard_sizes(203, world)
    outk = comm.zeros(k, be.round_up(203))
    local = torch.randn((k, be.round_up(sizes[rank])), generator=gen, device='cuda',
                        dtype=torch.float64)
    comm.allgather_columns(local, outk, sizes)
    torch.cuda.synchronize()
    parts = comm.allgather_objects(local[:, :sizes[rank]].cpu())
    report['allgather'] = bool(torch.equal(outk[:, :203].cpu(), torch.cat(parts, dim=1)))

    # --- fused reduce-over-samples + all-reduce against the unfused pair
    Xl = torch.randn((T, ldx), generator=gen, device='cuda', dtype=torch.float64)
    Xl[:, d:] = 0
    Z = torch.rand((T, k), generator=gen, device='cuda', dtype=torch.float64)
    E = torch.randn((k, k), generator=gen, device='cuda', dtype=torch.float64)
    fused = tensors[4]
    ws = be.Workspace(T, d, k)
    for name, Em in (('plain', None), ('epilogue', E)):
        ref = be.zeros(k, ldx)
        be.reduce_samples(Z, 1, k, Xl, T, d, k, ref, ws, E=Em)
        dist.all_reduce(ref)
        for rep in range(3):
            fused.zero_()
            ok = peer.reduce_samples_allreduce(Z, 1, k, Xl, T, T, d, k, fused, E=Em)
            torch.cuda.synchronize()
            report['fused_%s_%d' % (name, rep)] = bool(ok) and (
                bool(torch.equal(fused, ref)) if world == 2 else
                bool(torch.allclose(fused, ref, rtol=1e-13, atol=1e-12)))
    peer.check()

    # --- the sharded engines with peer collectives against the same engines over NCCL
    X, Z0, W0, C0 = problem(T=403, d=2600, k=8, path=DATA[0])       # direct-load kernels
    lo, hi = shard_bounds(X.shape[0], world, rank)
    big = np.random.RandomState(3).randn(1300, 44000) * 0.1          # strip kernels (fused path)
    blo, bhi = shard_bounds(big.shape[0], world, rank)
    bZ = np.random.RandomState(4).rand(1300, 8)
    bZ /= bZ.sum(axis=1, keepdims=True)
    bW = 0.1 * np.random.RandomState(5).randn(44000, 8)
    bC = np.random.RandomState(6).rand(8, 1300)
    bC /= bC.sum(axis=1, keepdims=True)
    results = {}
    for mode in ('1', '0'):
        os.environ['CDR_PEER_COLLECTIVES'] = mode
        g = gp._iterate_gpnh_convex_coding(X[lo:hi].copy(), Z0[lo:hi].copy(), W0.copy(),
                                           lambda_W=0.2, tolerance=1e-12, max_iterations=6,
                                           comm=comm)
        a = aa._iterate_aa(X[lo:hi].copy(), Z0[lo:hi].copy(), C0.copy(), np.ones(8),
                           tolerance=1e-12, max_iterations=6,
                           dictionary_solver_kwargs=dict(max_iterations=2), comm=comm)
        gb = gp._iterate_gpnh_convex_coding(big[blo:bhi].copy(), bZ[blo:bhi].copy(), bW.copy(),
                                            tolerance=1e-12, max_iterations=12, comm=comm)
        ab = aa._iterate_aa(big[blo:bhi].copy(), bZ[blo:bhi].copy(), bC.copy(), np.ones(8),
                            tolerance=1e-12, max_iterations=8,
                            dictionary_solver_kwargs=dict(max_iterations=1), comm=comm)
        results[mode] = (g, a, gb, ab)
    for idx, name in enumerate(('gpnh_small', 'aa_small', 'gpnh_strip', 'aa_strip')):
        p, n = results['1'][idx], results['0'][idx]
        cost_p, cost_n = (p[2], n[2]) if name.startswith('gpnh') else (p[3], n[3])
        report['engine_%s_cost' % name] = bool(np.isclose(cost_p, cost_n, rtol=1e-10, atol=0))
        report['engine_%s_Z' % name] = bool(np.allclose(p[0], n[0], rtol=0, atol=1e-7))
        report['engine_%s_D' % name] = bool(np.allclose(p[1], n[1], rtol=0, atol=1e-7))
    comm.close_peer()
    reports = comm.allgather_objects(report)
    if rank == 0:
        bad = sorted({key for rep in reports for key, ok in rep.items() if not ok})
        np.savez(out, n_checks=np.array([len(report)]), failed=np.array(bad, dtype=str))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--mode', required=True)
    ap.add_argument('--out', required=True)
    ap.add_argument('--data', default=None)
    args = ap.parse_args()
    DATA[0] = args.data
    {'gloo': run_gloo, 'nccl': run_nccl, 'peer': run_peer}[args.mode](args.out)
