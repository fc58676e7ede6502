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
        self.peer = None          # _peer.PeerGroup once setup_peer() has run ...
        self.peer_on = False      # ... and whether the current fit uses it

    # -- peer-memory collectives (default for NCCL groups; CDR_PEER_COLLECTIVES=0 disables) ---
    def setup_peer(self, shapes, inbox_shape):
        """Prepare the symmetric region for the fp64 tensors ``shapes`` (allocated afterwards,
        in this order, with :meth:`zeros`) and a per-rank inbox slot of ``inbox_shape``.
        Collective: every rank calls it with the same arguments.  Returns the PeerGroup, or
        None when peer collectives are off (then :meth:`zeros` gives ordinary tensors and the
        collectives go through ``torch.distributed``)."""
        from . import _peer
        self.peer_on = (self.enabled and self.device() == 'cuda' and
                        _peer.peer_collectives_enabled())
        if not self.peer_on:
            return None
        count = lambda shape: int(np.prod(shape))
        data_bytes = sum(_peer.round_up(count(s) * 8 + 8) for s in shapes)
        # the fused exchange keeps two alternating sets of 16-byte tagged words per double
        slot_bytes = 4 * count(inbox_shape) * 8
        if self.peer is not None and not self.peer.fits(data_bytes, slot_bytes):
            self.peer.close()
            self.peer = None
        if self.peer is None:
            self.peer = _peer.PeerGroup(self, data_bytes, slot_bytes)
        else:
            self.peer.reset()
        return self.peer

    def close_peer(self):
        """Unmap and free the symmetric regions (collective)."""
        if self.peer is not None:
            self.peer.close()
            self.peer = None
        self.peer_on = False

    def zeros(self, *shape):
        """fp64 zeros on the device: inside the symmetric region when peer collectives are
        set up (so the tensor can take part in them), an ordinary tensor otherwise."""
        if self.peer_on:
            return self.peer.zeros(*shape)
        from . import _backend as be
        return be.zeros(*shape)

    def _peer_tensor(self, t):
        import torch
        return (self.peer_on and t.is_cuda and t.dtype == torch.float64 and
                self.peer.contains(t))

    def allreduce_sum(self, t):
        """In-place sum over ranks (a no-op without a group)."""
        if self.enabled:
            if self._peer_tensor(t) and self.peer.can_allreduce(t):
                return self.peer.allreduce(t)
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
        if self._peer_tensor(out):
            return self.peer.allgather_columns(local, out, sizes)
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

    def device(self):
        """Where tensors handed to the collectives must live ('cuda' under NCCL)."""
        if self.enabled and self._dist.get_backend(self.group) == 'nccl':
            return 'cuda'
        return 'cpu'

    def sum_scalars(self, values):
        """Element-wise sum over ranks of a short list of Python floats."""
        import torch
        if not self.enabled:
            return [float(v) for v in values]
        t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=self.device())
        return [float(v) for v in self.allreduce_sum(t).cpu()]

    def local_rows(self, n_local):
        """(first global row, total rows) of this rank's block of ``n_local`` rows; the
        engines use balanced contiguous blocks (``shard_bounds``), which is checked here."""
        total = int(round(self.sum_scalars([n_local])[0]))
        lo, hi = shard_bounds(total, self.world, self.rank)
        if hi - lo != int(n_local):
            raise ValueError('rank %d holds %d rows; a balanced split of %d rows over %d ranks '
                             'gives it %d (see convex_dim_red._dist.shard_bounds)'
                             % (self.rank, n_local, total, self.world, hi - lo))
        return lo, total

    def broadcast(self, t, src):
        """In-place broadcast of a tensor from rank ``src``."""
        if self.enabled:
            self._dist.broadcast(t, src=src, group=self.group)
        return t

    def allgather_row_blocks(self, local, sizes):
        """Stack the (sizes[r], ld) row blocks of every rank into one (sum(sizes), ld) tensor
        (a device-to-device all-gather of the local rows of X).  Blocks travel as equal
        (max(sizes), ld) buffers; with equal shards the receive buffer is the result."""
        import torch
        if not self.enabled:
            return local[:sizes[0]]
        ld = local.shape[1]
        nmax = max(sizes)
        scratch = torch.empty((self.world, nmax, ld), dtype=local.dtype, device=local.device)
        if local.shape[0] == nmax and local.is_contiguous():
            mine = local
        else:
            mine = torch.zeros((nmax, ld), dtype=local.dtype, device=local.device)
            mine[:sizes[self.rank]].copy_(local[:sizes[self.rank]])
        self._dist.all_gather([scratch[r] for r in range(self.world)], mine, group=self.group)
        if min(sizes) == nmax:
            return scratch.view(self.world * nmax, ld)
        return torch.cat([scratch[r, :n] for r, n in enumerate(sizes)], dim=0)

    def merge_column_moments(self, mean, var, n_local, n_total):
        """Column mean and (biased) variance of the stacked rows of all ranks from each
        rank's own ``mean`` / ``var`` over its ``n_local`` rows:
        mean = sum_r n_r mean_r / n,  var = sum_r n_r (var_r + (mean_r - mean)^2) / n
        (two vector all-reduces; no E[x^2] - mean^2 cancellation)."""
        if not self.enabled:
            return mean, var
        total_mean = self.allreduce_sum(mean * float(n_local)) / float(n_total)
        spread = (var + (mean - total_mean) ** 2) * float(n_local)
        return total_mean, self.allreduce_sum(spread) / float(n_total)

    def global_top(self, values, offset, n):
        """The ``n`` largest entries over the concatenated ``values`` (NumPy, one block per
        rank starting at global index ``offset``): list of ``(value, global_index)``, largest
        first, lower index first among equals."""
        values = np.asarray(values)
        m = min(int(n), values.size)
        if m > 0:
            local = np.argsort(-values, kind='stable')[:m]
            cand = [(float(values[i]), int(offset + i)) for i in local]
        else:
            cand = []
        merged = [c for part in self.allgather_objects(cand) for c in part]
        merged.sort(key=lambda c: (-c[0], c[1]))
        return merged[:int(n)]

    def allgather_objects(self, obj):
        """One picklable object per rank, in rank order."""
        if not self.enabled:
            return [obj]
        parts = [None] * self.world
        self._dist.all_gather_object(parts, obj, group=self.group)
        return parts

    def broadcast_object(self, obj, src):
        """``obj`` of rank ``src`` on every rank."""
        if not self.enabled:
            return obj
        box = [obj if self.rank == src else None]
        self._dist.broadcast_object_list(box, src=src, group=self.group)
        return box[0]

    def sync_random_state(self, rng):
        """Make rank 0's generator state the state of ``rng`` on every rank (collective).

        Sample-sharded fits draw the full-size initial factors on every rank and rely on the
        replicas being bit-identical; with ``random_state=None`` each rank would otherwise
        seed from OS entropy and the replicas -- and with them the device-side `done` flags
        and the collective sequences -- would silently diverge.  A no-op when every rank was
        given the same seed."""
        if self.enabled:
            rng.set_state(self.broadcast_object(rng.get_state(), src=0))
        return rng

    def allgather_rows(self, local_np):
        """Gather NumPy row blocks of all ranks (used for the final weights)."""
        if not self.enabled:
            return local_np
        parts = [None] * self.world
        self._dist.all_gather_object(parts, local_np, group=self.group)
        return np.concatenate(parts, axis=0)
