"""Sample-axis sharding helpers (one process per GPU, ``torch.distributed``).

The alternating-update path shards naturally along the sample (time) axis
(SURVEY.md section 8e): rank g owns the rows ``X_g`` and ``Z_g``; every product that
reduces over samples (``Z'X``, ``C X``, ``Z'Z``, traces) is a local partial followed by
a sum all-reduce of a k x d or k x k buffer; products that reduce over features
(``X W``, ``(C X) X'``) are purely local and yield the rank's own columns, which are
all-gathered when a replicated k x T matrix is needed (archetypal analysis).  The
per-sample QPs need no communication.

NCCL is used on GPUs; the same code runs over ``gloo`` on CPU tensors, which is
how the host-side logic is tested without a GPU.
"""

import numpy as np


def shard_bounds(n_samples, world, rank):
    """Contiguous, balanced row range [lo, hi) of ``rank`` (first ranks get the remainder)."""
    base, rem = divmod(int(n_samples), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_sizes(n_samples, world):
    return [shard_bounds(n_samples, world, r)[1] - shard_bounds(n_samples, world, r)[0]
            for r in range(world)]


class Comm:
    """Thin wrapper over a ``torch.distributed`` process group (or no group at all)."""

    def __init__(self, group=None, enabled=None):
        import torch.distributed as dist
        self._dist = dist
        if enabled is None:
            enabled = dist.is_available() and dist.is_initialized()
        self.enabled = bool(enabled) and dist.get_world_size(group) > 1
        self.group = group
        self.world = dist.get_world_size(group) if self.enabled else 1
        self.rank = dist.get_rank(group) if self.enabled else 0

    def allreduce_sum(self, t):
        """In-place sum over ranks (a no-op without a group)."""
        if self.enabled:
            self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
        return t

    def allreduce_max(self, t):
        if self.enabled:
            self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX, group=self.group)
        return t

    def allgather_columns(self, local, out, sizes, scratch=None):
        """Gather the (k, sizes[r]) column blocks of every rank into ``out[:, :sum(sizes)]``.

        ``local`` may be a view with a larger leading dimension.  Blocks travel as
        contiguous (k, max(sizes)) buffers (shards differ by at most one row) and are
        copied into place.
        """
        import torch
        k = local.shape[0]
        if not self.enabled:
            out[:, :sizes[0]].copy_(local[:, :sizes[0]])
            return out
        nmax = max(sizes)
        if scratch is None:
            scratch = torch.zeros((self.world + 1, k, nmax), dtype=local.dtype,
                                  device=local.device)
        mine = scratch[self.world]
        mine[:, :sizes[self.rank]].copy_(local[:, :sizes[self.rank]])
        self._dist.all_gather([scratch[r] for r in range(self.world)], mine, group=self.group)
        if min(sizes) == nmax:
            # equal shards: one strided copy (world, k, n) -> (k, world * n)
            out.as_strided((k, self.world, nmax), (out.stride(0), nmax, 1)).copy_(
                scratch[:self.world].permute(1, 0, 2))
            return out
        lo = 0
        for r, n in enumerate(sizes):
            out[:, lo:lo + n].copy_(scratch[r, :, :n])
            lo += n
        return out

    def allgather_rows(self, local_np):
        """Gather NumPy row blocks of all ranks (used for the final weights)."""
        if not self.enabled:
            return local_np
        parts = [None] * self.world
        self._dist.all_gather_object(parts, local_np, group=self.group)
        return np.concatenate(parts, axis=0)
