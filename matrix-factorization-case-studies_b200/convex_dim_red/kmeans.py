"""k-means (Lloyd) with FurthestSum initialisation.

The reference has no k-means of its own: its drivers call
``sklearn.cluster.KMeans`` (``bin/run_hadisst_kmeans.py:128-131``,
``bin/run_jra55_kmeans.py:115-132``).  This module restates the Lloyd path of
scikit-learn 1.9.0 (``sklearn/cluster/_kmeans.py``, ``_k_means_lloyd.pyx``,
``_k_means_common.pyx``) on the GPU: data centring, tolerance scaling,
first-minimum assignment, centre update, empty-cluster relocation, the
label-equality / centre-shift stopping rule and the final E-step.  The two
passes over the data per iteration are the streaming contractions
``cdr_reduce_features`` (x.c for all samples and centres) and
``cdr_reduce_samples`` (per-cluster sums through a one-hot matrix).

Sample-sharded fits (``comm=Comm()``, one process per GPU, each holding a contiguous block
of rows): assignment is local; the per-cluster sums (k x d), the counts
and the changed-label count are summed over ranks; the column mean / variance are merged
from per-rank moments; empty clusters are moved onto the globally farthest samples; the
centre update is replicated.  FurthestSum seeding all-gathers the rows once and deals the
slabs of the Gram matrix to the ranks.
"""

import ctypes

import numpy as np
from sklearn.utils import check_random_state

from . import _backend as be
from ._dist import Comm
from .furthest_sum import dissimilarity_from_gram_device, furthest_sum_device


def _group(comm):
    return comm if comm is not None else Comm(enabled=False)


def _row_layout(n_local, comm):
    """(sizes of all ranks' row blocks, first global row of this rank, total rows)."""
    sizes = [int(n) for n in comm.allgather_objects(int(n_local))]
    return sizes, sum(sizes[:comm.rank]), sum(sizes)


def furthest_sum_centres(X, n_clusters, start_index, extra_steps=10, exclude=None, comm=None):
    """Indices of the FurthestSum picks on the rows of X (distances from the Gram matrix,
    as the estimators build them: archetypal_analysis.py:96-100).  With a process group X
    is this rank's row block and the indices are global; every rank gets the same picks."""
    comm = _group(comm)
    X = np.ascontiguousarray(X, dtype=np.float64)
    T, d = X.shape
    Xd = be.to_device_padded(X)
    if comm.enabled:
        sizes, _, T = _row_layout(T, comm)
        Xd = comm.allgather_row_blocks(Xd, sizes)
    K = be.gram(Xd, T, d, comm)
    D = dissimilarity_from_gram_device(K, T)
    return furthest_sum_device(D, T, n_clusters, start_index, exclude, extra_steps)


def gather_rows(X, picks, comm=None):
    """Rows ``picks`` (global indices) of the row-sharded matrix X on every rank: each rank
    fills in the rows it owns and a sum all-reduce completes the rest."""
    comm = _group(comm)
    X = np.asarray(X)
    if not comm.enabled:
        return np.array(X[np.asarray(picks)], dtype=np.float64)
    torch = be.require_cuda()
    _, lo, _ = _row_layout(X.shape[0], comm)
    rows = np.zeros((len(picks), X.shape[1]))
    for i, g in enumerate(picks):
        if lo <= g < lo + X.shape[0]:
            rows[i] = X[int(g) - lo]
    out = torch.from_numpy(rows).cuda()
    comm.allreduce_sum(out)
    return out.cpu().numpy()


def kmeans_lloyd(X, init_centres, tol=1e-4, max_iter=300, verbose=False, comm=None, stats=None,
                 _time_iterations=0):
    """``KMeans(init=init_centres, n_init=1, algorithm='lloyd').fit(X)``.

    ``stats`` (optional dict) receives ``loop_ms`` (the Lloyd iterations alone, CUDA events),
    ``n_iter`` and ``device_loop`` (whether the graph-replayed device loop ran).
    ``_time_iterations=N`` (bench.py) instead times N steady-state Lloyd iterations of the
    device loop from the initial centres -- the state block is re-armed before each, so none
    is skipped after convergence -- and returns the milliseconds per iteration.

    Returns ``(labels int32[T], centres k x d, inertia, n_iter)``.  With a process group X is
    this rank's row block and the labels are those of its rows; centres, inertia and n_iter
    are global and identical on all ranks.
    """
    comm = _group(comm)
    torch = be.require_cuda()
    lib = be.library()
    X = np.ascontiguousarray(X, dtype=np.float64)
    init_centres = np.ascontiguousarray(init_centres, dtype=np.float64)
    T, d = X.shape
    k = init_centres.shape[0]
    if k > be.MAX_COMPONENTS:
        raise ValueError('n_clusters > %d is not supported by the B200 build' % be.MAX_COMPONENTS)
    s = be.stream_ptr
    be.trace('kmeans: enter')
    Xd = be._upload_padded(X)          # private copy: it is centred in place below
    be.trace('kmeans: upload')
    ldx = Xd.stride(0)
    ldt = be.round_up(T)

    # _kmeans.py:285-294 (tolerance) and :1486-1493 (centring); NumPy's row-sequential order
    mean = be.zeros(ldx)
    var = be.zeros(ldx)
    be.check(lib.cdr_column_moments(Xd.data_ptr(), ldx, T, d, mean.data_ptr(), var.data_ptr(), s()),
             'cdr_column_moments')
    if comm.enabled:
        layout, row0, n_total = _row_layout(T, comm)
        mean, var = comm.merge_column_moments(mean, var, T, n_total)
    tol_abs = float(np.mean(var[:d].cpu().numpy()) * tol)
    be.check(lib.cdr_center_columns(Xd.data_ptr(), ldx, T, d, mean.data_ptr(), -1.0, s()),
             'cdr_center_columns')
    centres = be.to_device_padded(init_centres)
    be.check(lib.cdr_center_columns(centres.data_ptr(), ldx, k, d, mean.data_ptr(), -1.0, s()),
             'cdr_center_columns')

    labels = torch.full((T,), -1, dtype=torch.int32, device='cuda')
    onehot = be.zeros(k, ldt)
    xct = be.zeros(k, ldt)
    sums = be.zeros(k, ldx)
    cnorm = be.zeros(k)
    shift = be.zeros(k)
    dist = be.zeros(T)
    tally = torch.zeros(k + 1, dtype=torch.int32, device='cuda')   # counts | changed labels
    counts, changed = tally[:k], tally[k:]
    ws = be.Workspace(T, d, k)

    def e_step():
        be.check(lib.cdr_row_sqnorms(centres.data_ptr(), ldx, k, d, cnorm.data_ptr(), s()),
                 'cdr_row_sqnorms')
        be.reduce_features(centres, Xd, T, d, k, xct, ws)
        tally.zero_()
        be.check(lib.cdr_kmeans_labels(xct.data_ptr(), ldt, cnorm.data_ptr(), T, k,
                                       labels.data_ptr(), onehot.data_ptr(), ldt,
                                       counts.data_ptr(), changed.data_ptr(), s()),
                 'cdr_kmeans_labels')
        comm.allreduce_sum(tally)

    def relocate_sharded(empty, weights):
        """Empty clusters -> the globally farthest samples (the owner of each sample
        broadcasts its row and current label)."""
        sq_distances()
        far = comm.global_top(dist.cpu().numpy(), row0, empty.size)
        if not far or far[0][0] == 0:
            return
        row = be.zeros(ldx)
        for (_, g), cid in zip(far, empty):
            owner = next(r for r in range(comm.world) if g < sum(layout[:r + 1]))
            old = None
            if comm.rank == owner:
                row.copy_(Xd[g - row0, :])
                old = int(labels[g - row0].item())
            comm.broadcast(row, owner)
            old = comm.broadcast_object(old, owner)
            sums[old, :] -= row
            sums[int(cid), :] = row
            weights[int(cid)] = 1.0
            weights[old] -= 1.0

    def sq_distances():
        be.check(lib.cdr_kmeans_sqdist(Xd.data_ptr(), ldx, T, d, centres.data_ptr(), ldx,
                                       labels.data_ptr(), dist.data_ptr(), s()),
                 'cdr_kmeans_sqdist')

    def m_step(host_counts):
        """Centre update from `sums` / `counts` with the empty-cluster relocation of
        _k_means_common.pyx:167-212 (rare, host logic); returns the total squared shift."""
        weights = counts.to(torch.float64)
        if comm.enabled:
            if (host_counts == 0).any():
                relocate_sharded(np.where(host_counts == 0)[0], weights)
        elif (host_counts == 0).any():
            # move empty clusters onto the samples farthest from their current centres
            empty = np.where(host_counts == 0)[0]
            sq_distances()
            dh = dist.cpu().numpy()
            if np.max(dh) != 0:
                far = np.argpartition(dh, -empty.size)[:-empty.size - 1:-1]
                lab = labels.cpu().numpy()
                for idx, cid in enumerate(empty):
                    far_idx = int(far[idx])
                    old = int(lab[far_idx])
                    sums[old, :] -= Xd[far_idx, :]
                    sums[int(cid), :] = Xd[far_idx, :]
                    weights[int(cid)] = 1.0
                    weights[old] -= 1.0
        be.check(lib.cdr_kmeans_update(sums.data_ptr(), ldx, weights.data_ptr(),
                                       centres.data_ptr(), ldx, k, d, shift.data_ptr(), s()),
                 'cdr_kmeans_update')
        sh = shift.cpu().numpy()
        return float((np.sqrt(sh) ** 2).sum())           # _kmeans.py:733: (center_shift**2).sum()

    strict = False
    n_iter = 0
    be.trace('kmeans: moments, centring, buffers')
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    fused = (not comm.enabled) and bool(lib.cdr_kmeans_fused_applicable(T, d, k))
    if _time_iterations and not fused:
        raise be.BackendError('the device Lloyd loop does not cover this shape')
    if fused:
        # device-resident loop (csrc/kmeans_iter.cu): five kernels per Lloyd iteration, the
        # stopping rule on the device, replayed from a CUDA graph; the host reads the state
        # block every few iterations and only steps in for an empty cluster
        init = be.KmeansState()
        init.max_iter = int(max_iter)
        init.tol_abs = tol_abs
        st_buf = torch.from_numpy(np.frombuffer(bytes(init), dtype=np.uint8).copy()).cuda()
        wsb = torch.empty(lib.cdr_kmeans_workspace_bytes(T, d, k) // 8 + 1, dtype=torch.float64,
                          device='cuda')
        prob = be.KmeansProblem(
            Xd.data_ptr(), ldx, T, d, k, centres.data_ptr(), labels.data_ptr(),
            onehot.data_ptr(), ldt, sums.data_ptr(), cnorm.data_ptr(), shift.data_ptr(),
            counts.data_ptr(), st_buf.data_ptr(), wsb.data_ptr(), wsb.numel() * 8)

        def read_state():
            return be.KmeansState.from_buffer_copy(st_buf.cpu().numpy().tobytes())

        def write_state(st):
            st_buf.copy_(torch.from_numpy(np.frombuffer(bytes(st), dtype=np.uint8).copy()))

        def iterate():
            be.check(lib.cdr_kmeans_iterate_enqueue(ctypes.byref(prob), s()),
                     'cdr_kmeans_iterate_enqueue')

        def prepare():
            be.check(lib.cdr_kmeans_prepare_enqueue(ctypes.byref(prob), s()),
                     'cdr_kmeans_prepare_enqueue')

        prepare()

        be.trace('kmeans: device loop set-up')
        iterate()                                        # eager: warms every kernel up
        st = read_state()
        be.trace('kmeans: first iteration')
        if _time_iterations:
            armed = st_buf.clone()
            armed.copy_(torch.from_numpy(np.frombuffer(bytes(init), dtype=np.uint8).copy()))

            def one():
                st_buf.copy_(armed)
                iterate()
            prepare()
            graph = be.capture_graph(one)
            for _ in range(3):
                graph.replay()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(int(_time_iterations)):
                graph.replay()
            t1.record()
            torch.cuda.synchronize()
            return t0.elapsed_time(t1) / int(_time_iterations)
        graph, chunk, launched = None, 1, 1
        graph_after = be.graph_after(True)               # a Lloyd iteration is 4 long kernels
        while True:
            if st.done and st.needs_relocation:
                # the device stopped before the centre update: finish this iteration here
                shift_tot = m_step(counts.cpu().numpy())
                st.n_iter += 1
                st.needs_relocation = 0
                st.done = 0
                if not st.changed:
                    st.strict, st.done = 1, 1
                elif shift_tot <= tol_abs or st.n_iter >= max_iter:
                    st.done = 1
                st.changed = 0
                write_state(st)
                prepare()                                # the centres changed on this side
            if st.done:
                break
            if verbose:
                print('Iteration %d, center shift %.6e' % (st.n_iter - 1, st.shift_total))
            elif graph is None and launched >= graph_after and not be.graphs_disabled():
                graph = be.capture_graph(iterate)
                be.trace('kmeans: graph capture')
            for _ in range(1 if verbose else chunk):
                if graph is not None:
                    graph.replay()
                else:
                    iterate()
            launched += 1 if verbose else chunk
            st = read_state()
            chunk = min(2 * chunk, 16)
        n_iter, strict = int(st.n_iter), bool(st.strict)
    for it in (range(max_iter) if not fused else ()):
        n_iter = it + 1
        e_step()
        be.reduce_samples(onehot, ldt, 1, Xd, T, d, k, sums, ws)
        comm.allreduce_sum(sums)
        shift_tot = m_step(counts.cpu().numpy())
        moved = int(changed.item())
        if not moved:
            strict = True
            break
        if verbose:
            print('Iteration %d, center shift %.6e' % (it, shift_tot))
        if shift_tot <= tol_abs:
            break
    ev1.record()
    be.trace('kmeans: remaining iterations')
    if stats is not None:
        torch.cuda.synchronize()
        stats.update(loop_ms=ev0.elapsed_time(ev1), n_iter=n_iter, device_loop=bool(fused))
    if not strict:
        e_step()                                         # _kmeans.py:745-757
    sq_distances()
    total = be.zeros(1)
    be.check(lib.cdr_sum_vector(dist.data_ptr(), T, total.data_ptr(), s()), 'cdr_sum_vector')
    comm.allreduce_sum(total)
    inertia = float(total.item())
    be.check(lib.cdr_center_columns(centres.data_ptr(), ldx, k, d, mean.data_ptr(), 1.0, s()),
             'cdr_center_columns')
    out = (labels.cpu().numpy(), be.to_host(centres, k, d), inertia, n_iter)
    be.trace('kmeans: final E step, inertia, results to host')
    return out


def _row_sqnorms_and_device(X):
    """Device copy of X and its squared row norms (sklearn row_norms(X, squared=True))."""
    lib = be.library()
    T, d = X.shape
    Xd = be.to_device_padded(X)
    norms = be.zeros(T)
    be.check(lib.cdr_row_sqnorms(Xd.data_ptr(), Xd.stride(0), T, d, norms.data_ptr(),
                                 be.stream_ptr()), 'cdr_row_sqnorms')
    return Xd, norms.cpu().numpy()


def kmeans_plusplus(X, n_clusters, random_state=None, n_local_trials=None):
    """k-means++ seeding, restating ``sklearn.cluster.kmeans_plusplus`` /
    ``_kmeans_plusplus`` (scikit-learn 1.9.0 ``_kmeans.py:180-282``): same RNG calls in the
    same order, same candidate selection.  The inner products candidates . samples are the
    reduce-over-features pass; the O(n_samples) bookkeeping (potentials, cumulative sums,
    ``searchsorted``) is host logic on vectors.  Returns ``(centers, indices)``.
    """
    rng = check_random_state(random_state)
    X = np.ascontiguousarray(X, dtype=np.float64)
    T, d = X.shape
    if n_local_trials is None:
        n_local_trials = 2 + int(np.log(n_clusters))
    Xd, x_sq = _row_sqnorms_and_device(X)
    ws = be.Workspace(T, d, max(1, min(n_local_trials, be.MAX_COMPONENTS)))
    ldt = be.round_up(T)
    sample_weight = np.ones(T)

    def sq_distances(ids):
        """_euclidean_distances(X[ids], X, squared=True): (-2 x.y + |x|^2) + |y|^2, clipped."""
        ids = np.asarray(ids)
        M = be.to_device_padded(X[ids])
        dots = be.zeros(len(ids), ldt)
        be.reduce_features(M, Xd, T, d, len(ids), dots, ws)
        dist = -2.0 * be.to_host(dots, len(ids), T)
        dist += x_sq[ids][:, np.newaxis]
        dist += x_sq[np.newaxis, :]
        np.maximum(dist, 0, out=dist)
        return dist

    centers = np.empty((n_clusters, d))
    indices = np.full(n_clusters, -1, dtype=int)
    center_id = rng.choice(T, p=sample_weight / sample_weight.sum())
    centers[0] = X[center_id]
    indices[0] = center_id
    closest = sq_distances([center_id])
    current_pot = closest @ sample_weight
    for c in range(1, n_clusters):
        rand_vals = rng.uniform(size=n_local_trials) * current_pot
        candidate_ids = np.searchsorted(np.cumsum(sample_weight * closest, dtype=np.float64),
                                        rand_vals)
        np.clip(candidate_ids, None, closest.size - 1, out=candidate_ids)
        dist = sq_distances(candidate_ids)
        np.minimum(closest, dist, out=dist)
        candidates_pot = dist @ sample_weight.reshape(-1, 1)
        best = np.argmin(candidates_pot)
        current_pot = candidates_pot[best]
        closest = dist[best][np.newaxis, :]
        best_id = candidate_ids[best]
        centers[c] = X[best_id]
        indices[c] = best_id
    return centers, indices


def _is_same_clustering(labels1, labels2, n_clusters):
    """Same partition up to a relabelling (sklearn ``_is_same_clustering``)."""
    mapping = np.full(n_clusters, -1, dtype=np.int64)
    for a, b in zip(labels1, labels2):
        if mapping[a] == -1:
            mapping[a] = b
        elif mapping[a] != b:
            return False
    return True


class KMeans():
    """Minimal estimator around :func:`kmeans_lloyd` with the scikit-learn attribute
    names the reference's drivers read (``cluster_centers_``, ``labels_``, ``inertia_``,
    ``n_iter_``; bin/run_hadisst_kmeans.py:128-137).

    ``init`` is an explicit (n_clusters, n_features) array, ``'k-means++'`` or ``'random'``
    (both with scikit-learn's RNG call sequence, so a seeded fit picks the same seeds), or
    ``'furthest_sum'`` (start index drawn from ``random_state``, 10 replacement passes).
    With ``n_init > 1`` the run with the lowest inertia is kept (sklearn's selection rule).

    ``fit(X_local, comm=Comm())`` is the sample-sharded fit: every rank passes its contiguous
    block of rows (rank order = row order) and the same ``random_state``; ``labels_`` then
    holds the labels of all rows on every rank.  ``init='k-means++'`` is single-GPU only.
    """

    def __init__(self, n_clusters=8, init='furthest_sum', n_init=1, max_iter=300, tol=1e-4,
                 verbose=0, random_state=None, extra_steps=10):
        self.n_clusters = n_clusters
        self.init = init
        self.n_init = n_init
        self.max_iter = max_iter
        self.tol = tol
        self.verbose = verbose
        self.random_state = random_state
        self.extra_steps = extra_steps

    def _initial_centres(self, X, rng, Xc=None, comm=None):
        if comm is not None and comm.enabled and isinstance(self.init, str):
            n_total = _row_layout(X.shape[0], comm)[2]
            if self.init == 'furthest_sum':
                start = rng.randint(n_total)
                picks = furthest_sum_centres(X, self.n_clusters, start, self.extra_steps,
                                             comm=comm)
            elif self.init == 'random':
                picks = rng.choice(n_total, size=self.n_clusters, replace=False,
                                   p=np.ones(n_total) / n_total)
            else:
                raise NotImplementedError("sample-sharded k-means supports init = array, "
                                          "'furthest_sum' or 'random'; got %r" % self.init)
            return gather_rows(X, picks, comm)
        if isinstance(self.init, str):
            if self.init == 'furthest_sum':
                start = rng.randint(X.shape[0])
                picks = furthest_sum_centres(X, self.n_clusters, start, self.extra_steps)
                return X[picks].copy()
            if self.init == 'k-means++':
                # sklearn seeds on the mean-centred data (_kmeans.py:1486-1512)
                _, picks = kmeans_plusplus(Xc, self.n_clusters, random_state=rng)
                return X[picks].copy()
            if self.init == 'random':
                n = X.shape[0]
                picks = rng.choice(n, size=self.n_clusters, replace=False, p=np.ones(n) / n)
                return X[picks].copy()
            raise ValueError("init must be an array, 'k-means++', 'furthest_sum' or 'random'; "
                             "got %r" % self.init)
        init = np.asarray(self.init, dtype=np.float64)
        if init.shape != (self.n_clusters, X.shape[1]):
            raise ValueError('The shape of the initial centers %s does not match the number '
                             'of clusters %d and features %d' %
                             (init.shape, self.n_clusters, X.shape[1]))
        return init

    def fit(self, X, y=None, comm=None):
        X = np.ascontiguousarray(X, dtype=np.float64)
        rng = check_random_state(self.random_state)
        n_init = 1 if not isinstance(self.init, str) else max(1, int(self.n_init))
        best = None
        sharded = comm is not None and comm.enabled
        use_pp = isinstance(self.init, str) and self.init == 'k-means++' and not sharded
        Xc = X - X.mean(axis=0) if use_pp else None
        with be.DeviceCache([Xc] if Xc is not None else []):
            for _ in range(n_init):
                centres0 = self._initial_centres(X, rng, Xc, comm)
                result = kmeans_lloyd(X, centres0, tol=self.tol, max_iter=self.max_iter,
                                      verbose=bool(self.verbose), comm=comm)
                if sharded:
                    result = (comm.allgather_rows(result[0]),) + result[1:]
                # _kmeans.py:1530-1540: lower inertia and a genuinely different clustering
                if best is None or (result[2] < best[2] and
                                    not _is_same_clustering(result[0], best[0], self.n_clusters)):
                    best = result
        self.labels_, self.cluster_centers_, self.inertia_, self.n_iter_ = best
        return self

    def fit_predict(self, X, y=None):
        return self.fit(X).labels_

    def predict(self, X):
        X = np.ascontiguousarray(X, dtype=np.float64)
        labels, _, _, _ = _assign_only(X, self.cluster_centers_)
        return labels


def _assign_only(X, centres):
    """Labels of X for fixed centres (one E-step)."""
    torch = be.require_cuda()
    lib = be.library()
    T, d = X.shape
    k = centres.shape[0]
    Xd = be.to_device_padded(X)
    Cd = be.to_device_padded(np.ascontiguousarray(centres, dtype=np.float64))
    ldx, ldt = Xd.stride(0), be.round_up(T)
    labels = torch.full((T,), -1, dtype=torch.int32, device='cuda')
    onehot = be.zeros(k, ldt)
    xct = be.zeros(k, ldt)
    cnorm = be.zeros(k)
    counts = torch.zeros(k, dtype=torch.int32, device='cuda')
    changed = torch.zeros(1, dtype=torch.int32, device='cuda')
    ws = be.Workspace(T, d, k)
    be.check(lib.cdr_row_sqnorms(Cd.data_ptr(), ldx, k, d, cnorm.data_ptr(), be.stream_ptr()),
             'cdr_row_sqnorms')
    be.reduce_features(Cd, Xd, T, d, k, xct, ws)
    be.check(lib.cdr_kmeans_labels(xct.data_ptr(), ldt, cnorm.data_ptr(), T, k, labels.data_ptr(),
                                   onehot.data_ptr(), ldt, counts.data_ptr(), changed.data_ptr(),
                                   be.stream_ptr()), 'cdr_kmeans_labels')
    return labels.cpu().numpy(), None, None, None


def _calculate_uniform_reference_wk(X, n_clusters, n_init=10, n_jobs=None, random_state=None):
    """Within-cluster dispersion of one uniform reference data set (reference kmeans.py:18-34):
    same RNG draws in the same order -- the box sample, then the k-means++ seeding of the
    ``n_init`` restarts -- with this module's k-means in place of scikit-learn's.  ``n_jobs``
    is accepted for signature compatibility (the Lloyd iterations run on the GPU)."""
    rng = check_random_state(random_state)
    n_samples, n_features = X.shape
    feature_min = np.broadcast_to(np.min(X, axis=0), (n_samples, n_features))
    feature_max = np.broadcast_to(np.max(X, axis=0), (n_samples, n_features))
    random_data = ((feature_max - feature_min) * rng.uniform(
        size=(n_samples, n_features)) + feature_min)
    kmeans = KMeans(n_clusters=n_clusters, init='k-means++', n_init=n_init,
                    random_state=rng).fit(random_data)
    return kmeans.inertia_


def _calculate_pca_reference_wk(X, n_clusters, n_init=10, n_components=100, n_iter=10,
                                n_jobs=None, random_state=None):
    """As above with the box aligned to the leading right singular vectors of X (reference
    kmeans.py:37-64; the randomised ``TruncatedSVD`` is scikit-learn's, as in the reference)."""
    from sklearn.decomposition import TruncatedSVD
    rng = check_random_state(random_state)
    n_samples = X.shape[0]
    svd = TruncatedSVD(n_components=n_components, n_iter=n_iter, random_state=rng)
    svd.fit(X)
    Vh = svd.components_
    Xp = np.dot(X, np.transpose(Vh))
    feature_min = np.broadcast_to(np.min(Xp, axis=0), (n_samples, n_components))
    feature_max = np.broadcast_to(np.max(Xp, axis=0), (n_samples, n_components))
    random_data = ((feature_max - feature_min) * rng.uniform(
        size=(n_samples, n_components)) + feature_min)
    random_data = np.dot(random_data, Vh)
    kmeans = KMeans(n_clusters=n_clusters, init='k-means++', n_init=n_init,
                    random_state=rng).fit(random_data)
    return kmeans.inertia_


def _calculate_reference_wk(X, n_components, reference='uniform', random_state=None):
    if reference == 'uniform':
        return _calculate_uniform_reference_wk(X, n_components, random_state=random_state)
    if reference == 'pca':
        return _calculate_pca_reference_wk(X, n_components, random_state=random_state)
    raise ValueError("unrecognized reference distribution '%s'" % reference)


def gap_statistic(X, Wk, n_components, n_trials=100, reference='uniform', n_jobs=1,
                  random_state=None):
    """Calculate gap statistic for k-means clustering (reference kmeans.py:81-108).

    ``Wk`` is the inertia of the caller's own fit (the drivers pass ``model.inertia_``,
    bin/run_hadisst_kmeans.py:133); returns ``(gap, sk)`` with
    ``gap = mean(log Wk_ref) - log Wk``.  One seed per trial is drawn first, exactly as the
    reference does, so the reference data sets do not depend on the trial order.  ``n_jobs``
    is accepted for compatibility: the trials run one after the other on the GPU.
    """
    rng = check_random_state(random_state)
    X = np.ascontiguousarray(X, dtype=np.float64)

    random_seeds = []
    for _ in range(n_trials):
        has_seed_already = True
        while has_seed_already:
            seed = rng.randint(np.iinfo(np.int32).max)
            if seed not in random_seeds:
                random_seeds.append(seed)
                has_seed_already = False

    Wk_ref = np.array([_calculate_reference_wk(X, n_components, reference=reference,
                                               random_state=random_seeds[i])
                       for i in range(n_trials)])
    lnWk_ref = np.log(Wk_ref)
    sk = np.std(lnWk_ref) * np.sqrt(1 + 1.0 / n_trials)
    gap = lnWk_ref.mean() - np.log(Wk)
    return gap, sk


def gap_statistic_fit(X, n_clusters, n_trials=100, n_init=10, reference='uniform',
                      random_state=None):
    """Convenience form (not in the reference): fits k-means on ``X`` itself and returns
    ``gap_statistic(X, inertia, n_clusters, ...)``."""
    rng = check_random_state(random_state)
    X = np.ascontiguousarray(X, dtype=np.float64)
    model = KMeans(n_clusters=n_clusters, init='k-means++', n_init=n_init,
                   random_state=rng).fit(X)
    return gap_statistic(X, model.inertia_, n_clusters, n_trials=n_trials, reference=reference,
                         random_state=rng)
