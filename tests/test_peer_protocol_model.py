"""Model check of the synchronisation protocols of the peer-memory collectives (csrc/peer.cuh,
csrc/peer.cu, the SamplesExchange variant in csrc/stream_tma.cu) -- no GPU involved.

The CUDA kernels synchronise GPUs through each other's memory: the stand-alone collectives
with epoch flags around the data (st.release.sys / ld.acquire.sys), the fused
reduce-over-samples + all-reduce kernel and the one-shot exchanges in the kernel tails with
"tagged words" -- every 8-byte word carries 32 data bits and the 32-bit epoch of its exchange,
so data and flag are one store, the reader polls the words themselves, and two buffer sets
alternate by epoch parity.  Whether such a
protocol can deadlock, read a tile of the wrong launch, or overwrite data a slower rank
still needs does not depend on CUDA: it is a property of the order of flag and data
accesses.  This file restates that order as Python generators (one per CTA, one `yield`
per memory access), runs all CTAs of all ranks under many random interleavings -- ranks
drift apart by whole launches where the protocol allows it -- and asserts the invariants
on tagged data.  Memory is sequentially consistent here; the fences of the real kernels
(st.release.sys / ld.acquire.sys around the data accesses) are what reduces the hardware to
this model.  The kernels must keep the access order written here.
"""

import random

import pytest


class Region:
    """One rank's symmetric region (flags, counters, data tagged with the launch they belong to)."""

    def __init__(self, world, max_strips, max_ctas):
        # fused reduce-over-samples + all-reduce
        self.fused_epoch = 0
        self.fused_tickets = 0
        self.ready = [[0] * world for _ in range(max_strips)]
        self.done = [0] * max_strips
        self.inbox = [[None] * max_strips for _ in range(world)]   # [from rank][strip] -> epoch
        self.consumed = [[True] * max_strips for _ in range(world)]
        self.out = [0] * max_strips                                 # strip -> epoch of the sum
        self.out_readers = 0                                        # later kernels reading `out`
        # stand-alone all-reduce
        self.coll_seq = [0] * max_ctas
        self.coll_start = [[0] * max_ctas for _ in range(world)]
        self.coll_finish = [[0] * max_ctas for _ in range(world)]
        self.buf = []                                               # chunk -> ('partial'|'sum', n)
        self.launch = 0                                             # launch this rank's stream is in
        # all-gather of column blocks
        self.gathered = [0] * world                                 # block of rank r -> launch
        self.gather_readers = 0
        # tagged-word buffers (two sets by epoch parity): tile inbox [set][from][strip], result
        # buffer [set][strip], small-exchange slots [set][from]; value = epoch tag, and whether
        # the current content has been read by its consumer
        self.ll_inbox = [[[0] * max_strips for _ in range(world)] for _ in range(2)]
        self.ll_inbox_read = [[[True] * max_strips for _ in range(world)] for _ in range(2)]
        self.ll_res = [[0] * max_strips for _ in range(2)]
        self.ll_res_read = [[True] * max_strips for _ in range(2)]
        self.small_epoch = 0
        self.small_ll = [[0] * world for _ in range(2)]
        self.small_read = [[True] * world for _ in range(2)]


def fused_cta(mem, world, rank, cta, grid, nstrips):
    """One CTA of reduce_samples_tma_kernel<..., SamplesExchange> (consumer warps): tagged
    words, no flags, no fences (stream_tma.cu)."""
    mine = mem[rank]
    epoch = mine.fused_epoch + 1
    cur = epoch & 1
    yield
    strips = range(cta, nstrips, grid)
    for strip in strips:
        yield                                                   # accumulate the strip
        owner = strip % world
        # the epilogue pushes the tagged tile into this rank's slot of the owner's inbox
        assert mem[owner].ll_inbox_read[cur][rank][strip], 'tile overwritten before the owner read it'
        mem[owner].ll_inbox[cur][rank][strip] = epoch
        mem[owner].ll_inbox_read[cur][rank][strip] = False
        yield
        if owner == rank:
            for r in range(world):
                while mine.ll_inbox[cur][r][strip] != epoch:    # poll the words of rank r's tile
                    yield
                mine.ll_inbox_read[cur][r][strip] = True
                yield
            assert mine.out_readers == 0 or mine.out[strip] == epoch, \
                'sum stored while a later kernel of the previous launch still reads out'
            mine.out[strip] = epoch
            yield
            for r in range(world):
                if r != rank:
                    assert mem[r].ll_res_read[cur][strip], 'result overwritten before it was read'
                    mem[r].ll_res[cur][strip] = epoch
                    mem[r].ll_res_read[cur][strip] = False
                    yield
    for strip in strips:
        if strip % world == rank:
            continue
        while mine.ll_res[cur][strip] != epoch:                 # poll the result words
            yield
        mine.ll_res_read[cur][strip] = True
        assert mine.out_readers == 0 or mine.out[strip] == epoch
        mine.out[strip] = epoch
        yield
    # the last CTA to leave publishes the epoch
    mine.fused_tickets += 1
    if mine.fused_tickets == grid:
        mine.fused_tickets = 0
        yield
        mine.fused_epoch = epoch
    yield


def small_cta(mem, world, rank):
    """peer::cta_allreduce_small: the last CTA of a fused kernel exchanges its rank's sums."""
    mine = mem[rank]
    epoch = mine.small_epoch + 1
    cur = epoch & 1
    yield
    for r in range(world):
        assert mem[r].small_read[cur][rank], 'statistics overwritten before the peer read them'
        mem[r].small_ll[cur][rank] = epoch
        mem[r].small_read[cur][rank] = False
        yield
    for r in range(world):
        while mine.small_ll[cur][r] != epoch:
            yield
        mine.small_read[cur][r] = True
        yield
    mine.small_epoch = epoch
    yield


def reader_kernel(mem, rank, nstrips, launch):
    """The kernels that follow the fused one in the stream: they read the whole `out`."""
    mine = mem[rank]
    mine.out_readers += 1
    for strip in range(nstrips):
        assert mine.out[strip] == launch, 'out changed under a reader (or is stale)'
        yield
    mine.out_readers -= 1
    yield


def allreduce_cta(mem, world, rank, cta, grid, nchunks):
    """One CTA of peer_allreduce_kernel (two-shot, pull + push)."""
    mine = mem[rank]
    epoch = mine.coll_seq[cta] + 1
    yield
    for p in range(world):                                      # start barrier
        mem[p].coll_start[rank][cta] = epoch
        yield
    for p in range(world):
        while mine.coll_start[p][cta] < epoch:
            yield
    for c in range(rank + world * cta, nchunks, world * grid):   # chunks this CTA owns
        for r in range(world):
            kind, n = mem[r].buf[c]
            assert (kind, n) == ('partial', mine.launch), 'pulled a chunk of another launch'
            yield
        for r in range(world):
            mem[r].buf[c] = ('sum', mem[rank].launch)
            yield
    for p in range(world):                                      # finish barrier
        mem[p].coll_finish[rank][cta] = epoch
        yield
    for p in range(world):
        while mine.coll_finish[p][cta] < epoch:
            yield
    mine.coll_seq[cta] = epoch
    yield


def allgather_cta(mem, world, rank, cta, grid, launch):
    """One CTA of peer_allgather_columns_kernel (push between a start and a finish barrier)."""
    mine = mem[rank]
    epoch = mine.coll_seq[cta] + 1
    yield
    for p in range(world):
        mem[p].coll_start[rank][cta] = epoch
        yield
    for p in range(world):
        while mine.coll_start[p][cta] < epoch:
            yield
    if cta == 0:                                                # this CTA's share of the block
        for p in range(world):
            assert mem[p].gather_readers == 0, 'block pushed under a reader of the old contents'
            mem[p].gathered[rank] = launch
            yield
    for p in range(world):
        mem[p].coll_finish[rank][cta] = epoch
        yield
    for p in range(world):
        while mine.coll_finish[p][cta] < epoch:
            yield
    mine.coll_seq[cta] = epoch
    yield


def gather_reader_kernel(mem, world, rank, launch):
    mine = mem[rank]
    mine.gather_readers += 1
    for r in range(world):
        assert mine.gathered[r] == launch, 'gathered matrix incomplete or from another launch'
        yield
    mine.gather_readers -= 1
    yield


def producer_kernel(mem, rank, nchunks, launch):
    """The kernel that writes this rank's partial before the all-reduce of `launch`."""
    mine = mem[rank]
    mine.launch = launch
    mine.buf = mine.buf[:nchunks] + [None] * max(0, nchunks - len(mine.buf))
    for c in range(nchunks):
        mine.buf[c] = ('partial', launch)
        yield


def checker_kernel(mem, rank, nchunks, launch):
    for c in range(nchunks):
        assert mem[rank].buf[c] == ('sum', launch), 'buffer not fully reduced after the kernel'
        yield


def rank_stream(mem, world, rank, plan, rng):
    """The stream of one rank: kernels run one after another; the CTAs of a kernel are all
    resident and interleave arbitrarily."""
    for launch, (kind, grid, size) in enumerate(plan, start=1):
        if kind == 'fused':
            kernels = [[fused_cta(mem, world, rank, b, grid, size) for b in range(grid)],
                       [reader_kernel(mem, rank, size, launch_of(plan, launch, 'fused'))]]
        elif kind == 'small':
            kernels = [[small_cta(mem, world, rank)]]
        elif kind == 'allgather':
            kernels = [[allgather_cta(mem, world, rank, b, grid, launch) for b in range(grid)],
                       [gather_reader_kernel(mem, world, rank, launch)]]
        else:
            kernels = [[producer_kernel(mem, rank, size, launch)],
                       [allreduce_cta(mem, world, rank, b, grid, size) for b in range(grid)],
                       [checker_kernel(mem, rank, size, launch)]]
        for ctas in kernels:
            live = list(ctas)
            while live:
                g = rng.choice(live)
                try:
                    next(g)
                except StopIteration:
                    live.remove(g)
                yield


def launch_of(plan, upto, kind):
    """How many launches of `kind` the first `upto` entries of the plan contain."""
    return sum(1 for k, _, _ in plan[:upto] if k == kind)


def run(world, plan, seed, bias=None, max_steps=2_000_000):
    rng = random.Random(seed)
    mem = [Region(world, 64, 16) for _ in range(world)]
    streams = [rank_stream(mem, world, r, plan, random.Random(seed * 131 + r)) for r in range(world)]
    live = list(range(world))
    weights = bias or [1] * world
    steps = 0
    while live:
        r = rng.choices(live, weights=[weights[i] for i in live])[0]
        try:
            next(streams[r])
        except StopIteration:
            live.remove(r)
        steps += 1
        assert steps < max_steps, 'no progress: the protocol deadlocked (or livelocked)'
    return mem


@pytest.mark.parametrize('world', [2, 3, 4])
def test_fused_exchange_protocol(world):
    plan = [('fused', 3, 7)] * 5
    for seed in range(40):
        mem = run(world, plan, seed)
        assert all(m.fused_epoch == 5 for m in mem)
    # one rank far slower than the others (they may run ahead only as far as the flags allow)
    for seed in range(20):
        run(world, plan, seed, bias=[1] + [25] * (world - 1))
        run(world, plan, seed, bias=[25] * (world - 1) + [1])


@pytest.mark.parametrize('world', [2, 3, 8])
def test_tagged_word_exchanges_as_in_a_sharded_iteration(world):
    """The sharded GPNH iteration (fused k x d exchange, then the statistics exchange in the
    weights kernel's tail) and the AA one (two fused exchanges and three small ones), with
    ranks of very different speed and with fewer strips than ranks."""
    gpnh = [('fused', 4, 9), ('small', 1, 0)] * 5
    aa = [('fused', 3, 9), ('small', 1, 0), ('small', 1, 0), ('fused', 3, 9), ('small', 1, 0)] * 3
    few = [('fused', 2, 2), ('small', 1, 0)] * 4            # some ranks own no strip
    for plan in (gpnh, aa, few):
        for seed in range(12):
            mem = run(world, plan, seed)
            assert all(m.small_epoch == sum(1 for k, _, _ in plan if k == 'small') for m in mem)
            run(world, plan, seed, bias=[1] + [30] * (world - 1))
            run(world, plan, seed, bias=[30] * (world - 1) + [1])


def test_fused_exchange_survives_a_change_of_grid():
    """A later fit with another strip count / grid reuses the flags of the region: the
    launch-level epoch keeps every strip's flag values monotonic (per-CTA counters would
    not: a CTA index that only exists in the larger grid would start again from 1)."""
    plan = [('fused', 2, 9)] * 4 + [('fused', 5, 11)] * 3 + [('fused', 3, 3)] * 2
    for seed in range(30):
        mem = run(3, plan, seed)
        assert all(m.fused_epoch == 9 for m in mem)


@pytest.mark.parametrize('world', [2, 3, 5])
def test_allreduce_protocol(world):
    plan = [('allreduce', 2, 9), ('allreduce', 2, 9), ('allreduce', 1, 1), ('allreduce', 3, 14),
            ('allreduce', 2, 9)]
    for seed in range(40):
        run(world, plan, seed)
    for seed in range(15):
        run(world, plan, seed, bias=[1] + [30] * (world - 1))


def test_mixed_sequence_as_in_an_iteration():
    """A GPNH-like iteration: fused k x d sum, then the small statistics all-reduce, repeated;
    and an AA-like one with the all-gather of the column blocks in between."""
    plan = [('fused', 4, 6), ('allreduce', 1, 1)] * 6
    for seed in range(30):
        run(3, plan, seed)
    plan = [('fused', 4, 6), ('allgather', 2, 0), ('allreduce', 1, 1), ('allgather', 1, 0)] * 4
    for seed in range(30):
        run(3, plan, seed)
        run(4, plan, seed, bias=[40, 1, 40, 40])


def test_the_model_catches_a_broken_protocol():
    """Sanity of the checker itself: without the final wait on `done` a fast rank overwrites
    an inbox slot the owner has not consumed (or leaves with an incomplete result)."""
    def no_done_wait(mem, world, rank, cta, grid, nstrips):
        mine = mem[rank]
        epoch = mine.fused_epoch + 1
        yield
        for strip in range(cta, nstrips, grid):
            owner = strip % world
            assert mem[owner].consumed[rank][strip], 'tile overwritten before the owner read it'
            mem[owner].inbox[rank][strip] = epoch
            mem[owner].consumed[rank][strip] = False
            yield
            mem[owner].ready[strip][rank] = epoch
            yield
            if owner == rank:
                for r in range(world):
                    while mine.ready[strip][r] < epoch:
                        yield
                for r in range(world):
                    assert mine.inbox[r][strip] == epoch, 'tile of another launch in the sum'
                    mine.consumed[r][strip] = True
                    yield
        mine.fused_tickets += 1
        if mine.fused_tickets == grid:
            mine.fused_tickets = 0
            mine.fused_epoch = epoch
        yield

    import sys
    module = sys.modules[__name__]
    saved = module.fused_cta
    module.fused_cta = no_done_wait
    try:
        caught = 0
        for seed in range(30):
            try:
                run(3, [('fused', 2, 5)] * 4, seed, bias=[1, 30, 30])
            except AssertionError:
                caught += 1
        assert caught > 0
    finally:
        module.fused_cta = saved
