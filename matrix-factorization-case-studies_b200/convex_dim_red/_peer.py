"""Peer-memory collectives of the sample-sharded fit (``cdr_peer_*`` in include/cdr_b200.h).

One process per GPU.  Every rank allocates one *symmetric region* with the library, exports
it through CUDA IPC and maps the regions of the other ranks; the 64-byte handles travel
through the ``torch.distributed`` group (the only thing the group is used for here).  Tensors
that take part in a collective are carved out of the region by :meth:`PeerGroup.zeros`, so
the same tensor sits at the same offset on every rank and the library's kernels can push /
pull it over NVLink directly: the sum over ranks of a k x d partial, fused into the epilogue
of the reduce-over-samples kernel where the strip kernel applies, the sum of the small k x k
statistics, and the all-gather of k x T_local column blocks.

On by default for NCCL process groups (validated on 2 and 8 B200s, profiles/r02);
``CDR_PEER_COLLECTIVES=0`` falls back to NCCL through ``torch.distributed``.  PyTorch still
owns streams; the region itself is ``cudaMalloc`` memory owned by the library, wrapped
zero-copy as tensors.
"""

import ctypes
import os

from . import _backend as be

ALIGN = 512


def peer_collectives_enabled():
    return os.environ.get('CDR_PEER_COLLECTIVES', '1') != '0'


def round_up(n, m=ALIGN):
    return (int(n) + m - 1) // m * m


def region_layout(world, data_bytes, inbox_slot_bytes):
    """(inbox offset, data offset, total bytes) of a region:
    header | world inbox slots + one result slot | data."""
    slot = round_up(inbox_slot_bytes)
    inbox_offset = be.PEER_HEADER_BYTES
    data_offset = inbox_offset + (world + 1) * slot
    return inbox_offset, data_offset, data_offset + round_up(data_bytes)


class _DeviceSpan:
    """Exposes raw device memory through ``__cuda_array_interface__`` (zero-copy tensor)."""

    def __init__(self, address, n_doubles):
        self.__cuda_array_interface__ = {'shape': (int(n_doubles),), 'typestr': '<f8',
                                         'data': (int(address), False), 'version': 3}


class PeerGroup:
    """The symmetric regions of all ranks of ``comm`` plus a bump allocator over the local one."""

    def __init__(self, comm, data_bytes, inbox_slot_bytes):
        torch = be.require_cuda()
        if comm.world > be.MAX_PEERS:
            raise be.BackendError('peer collectives support at most %d ranks' % be.MAX_PEERS)
        lib = be.library()
        self.comm = comm
        self.slot_bytes = round_up(inbox_slot_bytes)
        self.inbox_offset, self.data_offset, self.nbytes = region_layout(
            comm.world, data_bytes, inbox_slot_bytes)
        base = ctypes.c_void_p()
        be.check(lib.cdr_peer_region_alloc(self.nbytes, ctypes.byref(base)),
                 'cdr_peer_region_alloc')
        self.base = base.value
        handle = ctypes.create_string_buffer(64)
        be.check(lib.cdr_peer_export(self.base, handle), 'cdr_peer_export')
        handles = comm.allgather_objects(handle.raw)
        self.struct = be.PeerGroupStruct()
        self.struct.world, self.struct.rank = comm.world, comm.rank
        self.struct.region_bytes = self.nbytes
        self.struct.inbox_offset = self.inbox_offset
        self.struct.inbox_slot_bytes = self.slot_bytes
        self.imported = []
        for r, raw in enumerate(handles):
            if r == comm.rank:
                self.struct.region[r] = self.base
                continue
            mapped = ctypes.c_void_p()
            be.check(lib.cdr_peer_import(raw, ctypes.byref(mapped)), 'cdr_peer_import')
            self.struct.region[r] = mapped.value
            self.imported.append(mapped.value)
        self.cursor = self.data_offset
        self._keep = []
        self._data_end = set()           # addresses at which an allocation's data ends
        torch.cuda.synchronize()
        comm.allgather_objects(0)        # every region is cleared and mapped before first use

    # -- allocation -----------------------------------------------------------
    def reset(self):
        """Start a new arena (the tensors of the previous fit are dead)."""
        self.cursor = self.data_offset
        self._keep = []
        self._data_end = set()

    def fits(self, data_bytes, inbox_slot_bytes):
        return (round_up(inbox_slot_bytes) <= self.slot_bytes and
                self.data_offset + round_up(data_bytes) <= self.nbytes)

    def zeros(self, *shape):
        """A zero-filled fp64 tensor inside the region (same offset on every rank, provided
        all ranks allocate the same shapes in the same order)."""
        torch = be.torch_mod()
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = round_up(n * 8 + 8)                 # at least one spare double after the data
        if self.cursor + nbytes > self.nbytes:
            raise be.BackendError('symmetric region exhausted')
        span = _DeviceSpan(self.base + self.cursor, nbytes // 8)
        flat = torch.as_tensor(span, device='cuda')
        self._keep.append(span)
        self._data_end.add(self.base + self.cursor + n * 8)
        self.cursor += nbytes
        flat.zero_()
        return flat[:n].view(*shape)

    def contains(self, t):
        return (t.is_cuda and self.base + self.data_offset <= t.data_ptr() <
                self.base + self.nbytes)

    def offset(self, t):
        return t.data_ptr() - self.base

    # -- collectives ------------------------------------------------------------
    def can_allreduce(self, t):
        """Whether :meth:`allreduce` handles this tensor: contiguous, inside the region, 16-byte
        aligned, and of even length -- or ending where an allocation's data ends (the spare
        double that follows is zero on every rank and simply rides along; a view that ends
        inside an allocation would drag a live neighbour into the sum)."""
        if not (self.contains(t) and t.is_contiguous()) or t.data_ptr() % 16:
            return False
        return t.numel() % 2 == 0 or (t.data_ptr() + t.numel() * 8) in self._data_end

    def allreduce(self, t, flags=None):
        """In-place sum over ranks of a fp64 tensor (view) living in the region."""
        if not self.can_allreduce(t):
            raise be.BackendError('tensor is not eligible for the peer all-reduce')
        n = t.numel() + t.numel() % 2
        be.check(be.library().cdr_peer_allreduce(
            ctypes.byref(self.struct), self.offset(t), n, be.ptr(flags), be.stream_ptr()),
            'cdr_peer_allreduce')
        return t

    def allgather_columns(self, local, out, sizes, flags=None):
        """Push this rank's (k, sizes[rank]) block into columns of ``out`` on every rank."""
        rank = self.comm.rank
        k = local.shape[0]
        be.check(be.library().cdr_peer_allgather_columns(
            ctypes.byref(self.struct), local.data_ptr(), local.stride(0), self.offset(out),
            out.stride(0), k, sum(sizes[:rank]), sizes[rank], max(sizes), be.ptr(flags),
            be.stream_ptr()), 'cdr_peer_allgather_columns')
        return out

    def reduce_samples_allreduce(self, L, sLi, sLt, X, T, T_min, d, k, out, E=None, flags=None):
        """``out = sum_r E (L_r X_r)`` with the exchange fused into the kernel's epilogue.
        Returns False when the fused kernel does not cover the shape (nothing was launched)."""
        rc = be.library().cdr_reduce_samples_allreduce(
            ctypes.byref(self.struct), be.ptr(L), sLi, sLt, be.ptr(X), X.stride(0), T, T_min, d,
            k, be.ptr(E), self.offset(out), out.stride(0), be.ptr(flags), be.stream_ptr())
        if rc == be.ERR_NOT_APPLICABLE:
            return False
        be.check(rc, 'cdr_reduce_samples_allreduce')
        return True

    # -- housekeeping -----------------------------------------------------------
    def check(self):
        """Raise if a wait inside a collective kernel timed out since the last check."""
        err = ctypes.c_int(0)
        be.check(be.library().cdr_peer_error(ctypes.byref(self.struct), ctypes.byref(err),
                                             be.stream_ptr()), 'cdr_peer_error')
        if err.value:
            what = {1: 'start barrier', 2: 'finish barrier', 3: 'strip tiles', 4: 'strip sums',
                    5: 'statistics of a peer'}
            raise be.BackendError('peer collective timed out waiting for %s (rank %d)'
                                  % (what.get(err.value, err.value), self.comm.rank))

    def close(self):
        torch = be.torch_mod()
        torch.cuda.synchronize()
        self.comm.allgather_objects(0)    # nobody is still inside a collective
        lib = be.library()
        for mapped in self.imported:
            lib.cdr_peer_release(mapped)
        self.imported = []
        self._keep = []
        if self.base:
            lib.cdr_peer_region_free(self.base)
            self.base = None
