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


def problem(T=203, d=300, k=5, seed=11):
    from oracle import convex_oracle as orc
    from convex_dim_red.datasets import synthetic_field
    X = synthetic_field(T, d, seed=seed)
    rs = np.random.RandomState(seed)
    W0 = 0.3 * rs.randn(d, k)
    Z0 = orc.right_stochastic_matrix((T, k), rs)
    C0 = orc.right_stochastic_matrix((k, T), rs)
    return X, Z0, W0, C0


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
    mx = torch.tensor([float(rank)], dtype=torch.float64)
    comm.allreduce_max(mx)
    assert mx.item() == world - 1
    if rank == 0:
        np.savez(out, ok=np.array([1]))
    dist.barrier()
    dist.destroy_process_group()


def run_nccl(out):
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from convex_dim_red._dist import Comm, shard_bounds
    from convex_dim_red import archetypal_analysis as aa
    from convex_dim_red import gpnh_convex_coding as gp
    comm = Comm()
    X, Z0, W0, C0 = problem(T=403, d=2600, k=8)
    T = X.shape[0]
    lo, hi = shard_bounds(T, comm.world, comm.rank)
    g = gp._iterate_gpnh_convex_coding(X[lo:hi].copy(), Z0[lo:hi].copy(), W0.copy(), lambda_W=0.2,
                                       tolerance=1e-12, max_iterations=6, comm=comm)
    a = aa._iterate_aa(X[lo:hi].copy(), Z0[lo:hi].copy(), C0.copy(), np.ones(C0.shape[0]),
                       tolerance=1e-12, max_iterations=6,
                       dictionary_solver_kwargs=dict(max_iterations=2), comm=comm)
    Zg = comm.allgather_rows(g[0])
    Za = comm.allgather_rows(a[0])
    if comm.rank == 0:
        np.savez(out, gZ=Zg, gW=g[1], gcost=g[2], gn=g[3], aZ=Za, aC=a[1], acost=a[3], an=a[4])
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--mode', required=True)
    ap.add_argument('--out', required=True)
    args = ap.parse_args()
    (run_gloo if args.mode == 'gloo' else run_nccl)(args.out)
