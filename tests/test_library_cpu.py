"""CPU-side checks of the product package (no GPU needed): the shared library
loads, exports every symbol ``include/cdr_b200.h`` declares, the ctypes struct
mirrors match the C layouts, and the host-side logic (validation, RNG draw
order, generic spg) behaves like the reference."""

import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, PKG_DIR

HEADER = os.path.join(ROOT, 'include', 'cdr_b200.h')


@pytest.fixture(scope='module')
def lib():
    from convex_dim_red import _backend as be
    if not os.path.exists(be.LIB_PATH):
        subprocess.check_call(['bash', os.path.join(PKG_DIR, 'csrc', 'build.sh')])
    return be.library()


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    names = re.findall(r'\b(cdr_[a-z0-9_]+)\s*\(', text)
    return sorted(set(names))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from convex_dim_red import _backend as be
    declared = _declared_functions()
    assert len(declared) > 30
    for name in declared:
        assert hasattr(lib, name), 'symbol %s declared in cdr_b200.h is not exported' % name
    missing = [n for n in declared if n not in be.SIGNATURES]
    assert not missing, 'no ctypes signature for %s' % missing
    extra = [n for n in be.SIGNATURES if n not in declared]
    assert not extra, 'bound but not declared in the header: %s' % extra
    assert b'sm_100a' in lib.cdr_version()


def test_struct_layouts_match_the_header(tmp_path):
    from convex_dim_red import _backend as be
    src = tmp_path / 'sizes.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "cdr_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %d %u %zu %zu %zu %zu\\n", '
                   'sizeof(cdr_spg_params),'
                   'sizeof(cdr_loop_state), sizeof(cdr_small_gram_desc), sizeof(cdr_aa_buffers),'
                   'offsetof(cdr_loop_state, tolerance), offsetof(cdr_aa_buffers, grad_scale),'
                   'sizeof(cdr_peer_group), offsetof(cdr_peer_group, inbox_offset),'
                   'CDR_MAX_PEERS, CDR_PEER_HEADER_BYTES, sizeof(cdr_gpnh_problem),'
                   'sizeof(cdr_aa_problem), offsetof(cdr_gpnh_problem, peers),'
                   'offsetof(cdr_aa_problem, peers));'
                   'return 0;}\n')
    exe = tmp_path / 'sizes'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(be.SpgParams), ctypes.sizeof(be.LoopState),
            ctypes.sizeof(be.SmallGramDesc), ctypes.sizeof(be.AaBuffers),
            be.LoopState.tolerance.offset, be.AaBuffers.grad_scale.offset,
            ctypes.sizeof(be.PeerGroupStruct), be.PeerGroupStruct.inbox_offset.offset,
            be.MAX_PEERS, be.PEER_HEADER_BYTES, ctypes.sizeof(be.GpnhProblem),
            ctypes.sizeof(be.AaProblem), be.GpnhProblem.peers.offset, be.AaProblem.peers.offset]
    assert got == want


def test_peer_collective_host_logic(lib):
    """Region layout, argument validation and the eligibility rules of the peer all-reduce,
    without a GPU (no kernel is launched: every call below is rejected on its arguments)."""
    from convex_dim_red import _backend as be, _peer
    inbox, data, total = _peer.region_layout(8, 3 * 1000 + 5, 8 * 44000 * 8)
    # header | world inbox slots + one result slot | data
    assert inbox == be.PEER_HEADER_BYTES and data == inbox + 9 * _peer.round_up(8 * 44000 * 8)
    assert total == data + _peer.round_up(3005) and total % _peer.ALIGN == 0
    assert _peer.peer_collectives_enabled() == (os.environ.get('CDR_PEER_COLLECTIVES', '1') != '0')

    g = be.PeerGroupStruct()
    g.world, g.rank, g.region_bytes = 2, 0, 16 << 20
    g.inbox_offset, g.inbox_slot_bytes = be.PEER_HEADER_BYTES, 1 << 20
    ref = ctypes.byref(g)
    assert lib.cdr_peer_allreduce(ref, be.PEER_HEADER_BYTES, 16, None, None) == -1   # no regions
    g.region[0], g.region[1] = 0x1000, 0x2000            # never dereferenced: rejected earlier
    assert lib.cdr_peer_allreduce(ref, be.PEER_HEADER_BYTES, 15, None, None) == -1   # odd count
    assert lib.cdr_peer_allreduce(ref, 64, 16, None, None) == -1                     # in header
    assert lib.cdr_peer_allreduce(ref, be.PEER_HEADER_BYTES + 8, 16, None, None) == -1
    assert lib.cdr_peer_allreduce(ref, (16 << 20) - 64, 16, None, None) == -1        # overruns
    g.world = 9
    assert lib.cdr_peer_allreduce(ref, be.PEER_HEADER_BYTES, 16, None, None) == -1
    g.world = 2
    assert lib.cdr_peer_allgather_columns(ref, 0x3000, 10, 3 << 20, 4, 8, 0, 10, 10, None,
                                          None) == -1                               # ldd < cols
    # fused reduce + all-reduce: output must fit an inbox slot, shapes the strip kernel does
    # not cover are reported as "not applicable" (the caller then runs the unfused pair)
    out_off = 4 << 20
    args = (None, 1, 8, 0x3000, 44000, 810, 810, 44000, 8, None)
    assert lib.cdr_reduce_samples_allreduce(ref, *args, out_off, 44000, None, None) == \
        be.ERR_NOT_APPLICABLE                                                        # slot too small
    small = (None, 1, 8, 0x3000, 320, 100, 100, 300, 8, None)
    assert lib.cdr_reduce_samples_allreduce(ref, *small, out_off, 320, None, None) == \
        be.ERR_NOT_APPLICABLE                                                        # too few strips
    assert lib.cdr_reduce_samples_allreduce(ref, *small, 8, 320, None, None) == -1   # offset in header
    # on a machine without a CUDA device the allocation reports the CUDA error, it does not crash
    base = ctypes.c_void_p()
    rc = lib.cdr_peer_region_alloc(1 << 10, ctypes.byref(base))
    assert rc == -1                                                                  # below header size


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    import convex_dim_red as cdr
    from convex_dim_red import _backend as be
    with pytest.raises(be.BackendError):
        cdr.simplex_project_rows(np.ones((2, 3)))
    with pytest.raises(be.BackendError):
        cdr.GPNHConvexCoding(n_components=2, random_state=0).fit(np.ones((6, 4)))
    with pytest.raises(be.BackendError):
        cdr.furthest_sum(np.zeros((4, 4)), 2, 0)


def test_product_package_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(PKG_DIR, 'convex_dim_red')):
        for f in files:
            if f.endswith('.py'):
                text = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in text.replace('convex_oracle', 'oracle') or \
                    not re.search(r'^\s*(from|import)\s+oracle', text, flags=re.M), f


def test_exports_match_reference_package():
    import convex_dim_red as cdr
    for name in ('ArchetypalAnalysis', 'KernelAA', 'GPNHConvexCoding', 'furthest_sum',
                 'gap_statistic', 'simplex_project_rows', 'simplex_project_columns', 'spg',
                 'left_stochastic_matrix', 'right_stochastic_matrix'):
        assert hasattr(cdr, name)
    from convex_dim_red.simplex_projection import simplex_project_vector      # noqa: F401
    from convex_dim_red.archetypal_analysis import (                          # noqa: F401
        _iterate_kernel_aa, _kernel_aa_cost, _update_kernel_aa_dictionary,
        _update_kernel_aa_weights, _iterate_aa, _update_aa_dictionary)
    from convex_dim_red.gpnh_convex_coding import (                           # noqa: F401
        _gpnh_cost, _iterate_gpnh_convex_coding, _update_gpnh_dictionary, _update_gpnh_weights)


def test_stochastic_matrices_draw_order():
    import convex_dim_red as cdr
    rs = np.random.RandomState(0)
    ref = np.random.RandomState(0)
    a = cdr.right_stochastic_matrix((4, 7), random_state=rs)
    b = cdr.left_stochastic_matrix((5, 3), random_state=rs)
    u = ref.uniform(size=(4, 7))
    np.testing.assert_array_equal(a, u / u.sum(axis=1)[:, None])
    v = ref.uniform(size=(5, 3))
    np.testing.assert_array_equal(b, v / v.sum(axis=0)[None, :])
    np.testing.assert_allclose(a.sum(axis=1), 1.0, atol=1e-15)
    np.testing.assert_allclose(b.sum(axis=0), 1.0, atol=1e-15)


def test_generic_spg_matches_golden(golden):
    # reference tests/test_spg.py:37-90 and the golden box-constrained problem
    import convex_dim_red as cdr
    x, fx, n_it, n_fe = cdr.spg(lambda x: x ** 4 + 2 * x ** 2 + 1, lambda x: 4 * x ** 3 + 4 * x,
                                0.4, project=lambda x: min(max(x, -1.0), 0.5))
    np.testing.assert_allclose([x, fx, n_it, n_fe], golden['spg/quartic/out'], rtol=1e-12,
                               atol=1e-14)
    assert abs(x) < 1e-6 and abs(fx - 1) < 1e-6
    M, y, x0 = golden['spg/box/M'], golden['spg/box/y'], golden['spg/box/x0']
    x, fx, n_it, n_fe = cdr.spg(lambda v: 0.5 * (M.dot(v) - y).dot(M.dot(v) - y),
                                lambda v: M.T.dot(M.dot(v) - y), x0,
                                project=lambda v: np.fmin(np.fmax(v, 0.0), 0.3))
    np.testing.assert_allclose(x, golden['spg/box/x'], rtol=1e-9)
    np.testing.assert_allclose([fx, n_it, n_fe], golden['spg/box/stats'], rtol=1e-12)
    # unconstrained scalar quadratic (tests/test_spg.py:13-35 style)
    x, fx, _, _ = cdr.spg(lambda x: (x - 2.0) ** 2, lambda x: 2 * (x - 2.0), 10.0)
    assert abs(x - 2.0) < 1e-6 and fx < 1e-10


def test_validation_utils():
    from convex_dim_red.validation_utils import (check_array_shape, check_stochastic_matrix,
                                                 check_unit_axis_sums)
    good = np.full((3, 4), 0.25)
    check_stochastic_matrix(good, (3, 4), 'test', axis=1)
    with pytest.raises(ValueError):
        check_array_shape(good, (4, 3), 'test')
    with pytest.raises(ValueError):
        check_unit_axis_sums(good, 'test', axis=0)


def test_solver_option_defaults():
    from convex_dim_red import _backend as be
    p = be.make_spg_params({})
    assert (p.gamma, p.memory, p.sigma_one, p.sigma_two) == (1e-4, 1, 0.1, 0.9)
    assert (p.max_iterations, p.max_feval, p.alpha0) == (1000, 2000, -1.0)
    from convex_dim_red.archetypal_analysis import _dictionary_params
    p = _dictionary_params({'max_iterations': 1})
    assert (p.max_iterations, p.max_feval) == (1, 1000000)
    with pytest.raises(ValueError):
        be.make_spg_params({'memory': 99})


def test_workspace_size_queries_are_consistent(lib):
    """Host-only size queries: the Gram workspace must cover every slab the Gram routine
    hands to the feature reduction (64-row slabs and the short last slab)."""
    for T, d in ((18000, 44000), (1620, 44000), (700, 41800), (50, 33), (1000, 100000), (65, 5000)):
        need = lib.cdr_gram_workspace_bytes(T, d)
        assert need >= lib.cdr_reduce_features_workspace_bytes(T, d, 64)
        if T % 64:
            assert need >= lib.cdr_reduce_features_workspace_bytes(T, d, T % 64)
        for k in (1, 8, 9, 16, 17, 64):
            assert lib.cdr_reduce_samples_workspace_bytes(T, d, k) >= 0
            assert lib.cdr_reduce_features_workspace_bytes(T, d, k) > 0
    assert lib.cdr_small_gram_workspace_bytes() >= 4 * 64 * 64 * 8
    assert lib.cdr_furthest_sum_workspace_bytes(1620, 8, 10) >= 1620 * 8 * (1 + 17)
    assert lib.cdr_launch_count() == 0 or lib.cdr_launch_count() > 0


def test_stream_plan_invariants(lib):
    """The host-side dispatch of the two streaming passes (which kernel, strip width, ring depth,
    shared memory) checked over many shapes without a GPU: strips tile the padded feature axis
    exactly, every strip is non-empty, the ring fits in shared memory, and the workspace the
    size queries report covers the per-strip partials the strip kernel writes."""
    rng = np.random.RandomState(11)
    shapes = [(1620, 44000), (18000, 44000), (700, 41800), (400, 6000), (64, 9472), (63, 9472),
              (50000, 100000), (1, 1), (7, 33), (1620, 1620)]
    shapes += [(int(rng.randint(1, 30000)), int(rng.randint(1, 120000))) for _ in range(150)]
    out = (ctypes.c_int * 12)()
    used = {'samples': 0, 'features': 0}
    for T, d in shapes:
        dpad = (d + 31) // 32 * 32
        for k in (1, 3, 8, 9, 16, 17, 24, 32, 33, 64):
            kp = (k + 7) // 8 * 8
            for epi in (0, 1):
                assert lib.cdr_debug_stream_plan(T, d, k, epi, out) == 0
                s_on, s_tc, s_n, s_st, s_smem, f_on, f_tc, f_n, f_st, f_smem, nsplit, nchunk = out
                assert nsplit >= 1 and nchunk >= 1
                assert lib.cdr_reduce_samples_workspace_bytes(T, d, k) >= (nsplit > 1) * nsplit * k * dpad * 8
                assert lib.cdr_reduce_features_workspace_bytes(T, d, k) >= nchunk * T * kp * 8
                for on, tc, n, st, smem, tcmax, name in (
                        (s_on, s_tc, s_n, s_st, s_smem, 512 if k <= 16 else 256, 'samples'),
                        (f_on, f_tc, f_n, f_st, f_smem, 384, 'features')):
                    if not on:
                        continue
                    used[name] += 1
                    assert T >= 64 and 64 <= tc <= tcmax and tc % 16 == 0
                    assert n * tc >= dpad > (n - 1) * tc          # exact cover, no empty strip
                    assert n >= 148 // 2
                    assert 2 <= st <= 8 and smem <= 227 * 1024
                    assert st * 16 * tc * 8 >= 32 * 1024           # enough bytes in flight per SM
                if s_on:
                    assert k <= 32 and (not epi or k <= 16)
                if f_on:
                    assert k <= 16
                    assert lib.cdr_reduce_features_workspace_bytes(T, d, k) >= f_n * T * kp * 8
                    assert f_n * kp * 10 <= dpad                   # partials <= 10 % of X
        if T % 64 or True:
            for rows in {64, T % 64} - {0}:
                lib.cdr_debug_stream_plan(T, d, min(rows, 64), 0, out)
                if out[5]:
                    assert lib.cdr_gram_workspace_bytes(T, d) >= out[7] * T * ((rows + 7) // 8 * 8) * 8
    assert used['samples'] > 100 and used['features'] > 100
    assert lib.cdr_debug_stream_plan(10, 10, 65, 0, out) != 0


def test_model_selection_host_helpers():
    from convex_dim_red import model_selection as ms
    X = np.arange(40, dtype=float).reshape(20, 2)
    train, val = ms.train_validation_split(X, 0.1)
    assert train.shape == (18, 2) and val.shape == (2, 2)          # ceil(0.9 * 20) = 18
    a = np.array([[1.0, 2.0], [3.0, 6.0]])
    b = np.array([[1.0, 0.0], [1.0, 2.0]])
    # per-column RMSE, uniformly averaged: sqrt((0 + 4) / 2), sqrt((4 + 16) / 2)
    np.testing.assert_allclose(ms.root_mean_squared_error(a, b),
                               0.5 * (np.sqrt(2.0) + np.sqrt(10.0)))


def test_preprocessing_host_helpers():
    from convex_dim_red import preprocessing as pp
    lat = np.array([-60.0, 0.0, 60.0, 90.0])
    np.testing.assert_allclose(pp.latitude_weights(lat, 'cos'), [0.5, 1.0, 0.5, 0.0], atol=1e-15)
    np.testing.assert_allclose(pp.latitude_weights(lat, 'scos') ** 2, [0.5, 1.0, 0.5, 0.0], atol=1e-15)
    np.testing.assert_array_equal(pp.latitude_weights(lat, 'none'), np.ones(4))
    with pytest.raises(ValueError):
        pp.latitude_weights(lat, 'bogus')
    rs = np.random.RandomState(0)
    field = rs.standard_normal((10, 4, 3))
    field[:, 1, 2] = np.nan                      # a land cell
    field[3, 0, 0] = np.nan                      # a single missing value also drops the column
    w = pp.latitude_weights(lat, 'scos')
    flat = pp.weight_and_flatten(field, w)
    assert flat.shape == (10, 12) and flat.flags.c_contiguous
    np.testing.assert_allclose(flat[:, 0 * 3 + 1], field[:, 0, 1] * w[0])
    valid, missing = pp.drop_missing_features(flat)
    assert valid.shape == (10, 10) and missing.sum() == 2 and not np.isnan(valid).any()
    back = pp.restore_features(valid[:2], missing, grid_shape=(4, 3))
    assert back.shape == (2, 4, 3) and np.isnan(back[:, 1, 2]).all() and np.isnan(back[:, 0, 0]).all()
    np.testing.assert_allclose(back[:, 2, 1], valid[:2, list(np.where(~missing)[0]).index(2 * 3 + 1)])
    train, val, missing2, w2 = pp.prepare_field(field, lat, 'scos', validation_frac=0.1)
    assert train.shape == (9, 10) and val.shape == (1, 10) and np.array_equal(missing, missing2)
    with pytest.raises(ValueError):
        pp.weight_and_flatten(field[0], w)


def test_graph_capture_policy(monkeypatch):
    """Engines on the whole-iteration C entry points launch 32 iterations eagerly before they
    capture the iteration graph (a capture costs more than a short call's launches); the
    general kernel sequences capture right after the first iteration; CDR_GRAPH_AFTER
    overrides both; CDR_LIBRARY selects another build of the library."""
    from convex_dim_red import _backend as be
    monkeypatch.delenv('CDR_GRAPH_AFTER', raising=False)
    assert be.graph_after(True) == 32 and be.graph_after(False) == 1
    monkeypatch.setenv('CDR_GRAPH_AFTER', '5')
    assert be.graph_after(True) == 5 and be.graph_after(False) == 5
    monkeypatch.setenv('CDR_GRAPH_AFTER', '0')
    assert be.graph_after(True) == 1
    assert be.LIB_PATH.endswith('libcdr_b200.so')


def test_kmeans_device_loop_covers_sixteen_clusters(lib):
    # shape-only query (no GPU needed): the strip plan of BASELINE configs[2]
    assert lib.cdr_kmeans_fused_applicable(700, 41800, 8) == 1
    assert lib.cdr_kmeans_fused_applicable(700, 41800, 16) == 1
    assert lib.cdr_kmeans_fused_applicable(700, 41800, 17) == 0
    assert lib.cdr_kmeans_fused_applicable(60, 200, 4) == 0
