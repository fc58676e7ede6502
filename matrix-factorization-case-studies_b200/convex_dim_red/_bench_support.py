"""Timing harness used by ``bench.py`` (kept in the package so the engines'
internals stay private to it)."""

import time

import numpy as np

from . import _backend as be


def _make_engine(args, X, Z0, F0, Xd=None, comm=None):
    from .archetypal_analysis import _AaEngine
    from .gpnh_convex_coding import _GpnhEngine
    big = 10 ** 6
    if args.workload == 'gpnh':
        return _GpnhEngine(X, Z0, F0, lambda_W=0.0, tolerance=0.0, max_iterations=big,
                           require_monotonic_cost_decrease=False, X_device=Xd, comm=comm)
    if getattr(args, 'formulation', 'stream') == 'gram':
        from .archetypal_analysis import _ShapeOnly
        T, d = X.shape
        trace = float(be.frobenius_sq(Xd, T, d).item())
        return _AaEngine(_ShapeOnly((T, T)), Z0, F0, np.ones(F0.shape[0]), 'kernel', tolerance=0.0,
                         max_iterations=big, require_monotonic_cost_decrease=False,
                         dictionary_solver_kwargs=dict(max_iterations=1),
                         data_device=be.gram(Xd, T, d), trace_data=trace, grad_scale=1.0 / T)
    return _AaEngine(X, Z0, F0, np.ones(F0.shape[0]), 'feature', tolerance=0.0,
                     max_iterations=big, require_monotonic_cost_decrease=False,
                     dictionary_solver_kwargs=dict(max_iterations=1), data_device=Xd, comm=comm)


def _time_launches(fn, reps=10):
    torch = be.torch_mod()
    fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run_benchmark(args, X, Z0, F0, rank, world, sampler):
    torch = be.require_cuda()
    import torch.distributed as dist
    lib = be.library()
    T, d = X.shape
    k = args.components
    from ._dist import Comm
    comm = Comm() if world > 1 else None
    Xd = be.to_device_padded(X)
    eng = _make_engine(args, X, Z0, F0, Xd, comm)
    eng.initial_cost()
    n0 = lib.cdr_launch_count()
    eng.iteration()                                  # eager: first warm-up step
    launches_per_step = lib.cdr_launch_count() - n0
    graph = None if (be.graphs_disabled() or (world > 1 and not be.graph_collectives())) else be.capture_graph(eng.iteration)
    step = graph.replay if graph is not None else eng.iteration
    # clocks / throttle reasons are sampled from the warm-up through the timed region and the
    # per-kernel timings (the timed region alone lasts a few milliseconds, shorter than one
    # nvidia-smi query)
    sampler.__enter__()
    for _ in range(max(args.warmup - 1, 0)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    st = eng.state.read()
    if args.workload == 'aa':
        # the first (eager) AA step also rebuilds C K for the projected start: count a
        # steady-state step instead
        n1 = lib.cdr_launch_count()
        eng.iteration()
        launches_per_step = lib.cdr_launch_count() - n1

    # ---- the streaming passes on their own (roofline of the dominant kernel)
    ws = eng.ws
    gram_mode = args.workload == 'aa' and getattr(args, 'formulation', 'stream') == 'gram'
    if gram_mode:
        t_gram = _time_launches(lambda: be.gram(Xd, T, d), reps=3)
        roofline = None
        kernels = {'gram_build_ms': t_gram, 'gram_build_tflops': 2.0 * T * T * d / (t_gram * 1e-3) / 1e12,
                   'step_ms': ms / args.steps}
    elif args.workload == 'gpnh':
        t_samples = _time_launches(lambda: be.reduce_samples(
            eng.Z, 1, k, eng.X, T, d, k, eng.WT, ws, E=eng.P))
        t_features = _time_launches(lambda: be.reduce_features(eng.WT, eng.X, T, d, k, eng.XWt, ws))
        zsave = eng.Z.clone()
        t_qp = _time_launches(lambda: (eng.Z.copy_(zsave), be.quad_simplex_spg_batched(
            eng.WtW, None, eng.XWt, 1, eng.ldt, eng.Z, T, k, eng.params)), reps=5)
        passes = 2
    else:
        t_samples = _time_launches(lambda: be.reduce_samples(
            eng.D, eng.ldt, 1, eng.X, T, d, k, eng.tmp_kd, ws))
        t_features = _time_launches(lambda: be.reduce_features(eng.tmp_kd, eng.X, T, d, k, eng.DK, ws))
        zsave = eng.Z.clone()
        t_qp = _time_launches(lambda: (eng.Z.copy_(zsave), be.quad_simplex_spg_batched(
            eng.CKCt, eng.alpha, eng.CK, 1, eng.ldt, eng.Z, T, k, eng.w_params)), reps=5)
        passes = 4
    pass_bytes = 8.0 * T * d
    if not gram_mode:
        slow, name = max((t_samples, 'reduce_samples_tma_kernel'),
                         (t_features, 'reduce_features_strip_kernel'))
        roofline = {'kernel': name, 'achieved': pass_bytes / (slow * 1e-3) / 1e9,
                    'algorithmic_bytes_per_launch': pass_bytes, 'ms_per_launch': slow,
                    'traffic': None}
    kernels = kernels if gram_mode else {'reduce_samples_ms': t_samples, 'reduce_features_ms': t_features,
               'reduce_samples_gbs': pass_bytes / (t_samples * 1e-3) / 1e9,
               'reduce_features_gbs': pass_bytes / (t_features * 1e-3) / 1e9,
               'qp_batched_ms': t_qp, 'passes_per_step': passes,
               'step_ms': ms / args.steps,
               'streaming_share_of_step': passes / 2.0 * (t_samples + t_features) /
               (ms / args.steps)}

    # keep the GPU busy with 1000 more steps (a fixed count: every rank must issue the same
    # collectives) so that several clock samples land under load, then stop sampling
    for _ in range(1000):
        step()
    torch.cuda.synchronize()
    sampler.__exit__(None, None, None)

    # ---- end to end through the public NumPy API (host buffers, pinned)
    e2e = run_e2e(args, X, Z0, F0, world, comm)
    converge = run_to_convergence(args, X, Z0, F0, comm) if world == 1 else None

    return {'value': world * args.steps / (ms * 1e-3), 'ms_per_step': ms / args.steps,
            'clocks': sampler.summary(), 'gpu_launches': int(launches_per_step * args.steps),
            'roofline': roofline, 'kernels': kernels, 'e2e': e2e, 'final_cost': st.cost,
            'time_to_converge': converge}


def run_e2e(args, X, Z0, F0, world, comm=None):
    """One public-API call of `steps` outer iterations: X (pinned host memory) is uploaded
    inside the timed region, the factors and the cost history come back as NumPy arrays."""
    torch = be.torch_mod()
    import torch.distributed as dist
    from . import archetypal_analysis as aa
    from . import gpnh_convex_coding as gp
    Xp = torch.from_numpy(X).pin_memory()
    Xn = Xp.numpy()
    K = args.steps

    def call():
        if args.workload == 'gpnh':
            out = gp._iterate_gpnh_convex_coding(
                Xn, Z0, F0, lambda_W=0.0, tolerance=0.0, max_iterations=K,
                require_monotonic_cost_decrease=False, comm=comm)
        else:
            out = aa._iterate_aa(
                Xn, Z0, F0, np.ones(F0.shape[0]), tolerance=0.0, max_iterations=K,
                require_monotonic_cost_decrease=False,
                dictionary_solver_kwargs=dict(max_iterations=1), comm=comm)
        return out[0].nbytes + out[1].nbytes + 8 * K

    call()                                   # untimed warm-up call (allocator, module load)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    d2h = call()
    torch.cuda.synchronize()
    elapsed = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([elapsed], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    h2d = X.nbytes + Z0.nbytes + F0.nbytes
    return {'value': world * K / elapsed, 'unit': 'iterations/s',
            'h2d_bytes_per_step': h2d / K, 'd2h_bytes_per_step': d2h / K,
            'call': 'one _iterate_%s call of %d outer iterations (the body of fit_transform): '
                    'X uploaded once from pinned host memory, factors read back'
                    % ('gpnh_convex_coding' if args.workload == 'gpnh' else 'aa', K),
            'seconds': elapsed}


def run_to_convergence(args, X, Z0, F0, comm=None, tolerance=1e-4, max_iterations=10000):
    """Wall time of one public-API fit to the drivers' stopping rule (abs_delta_f, tolerance
    1e-4 as in bin/run_hadisst_aa_wrapper.sh:44, max_iterations 10 000), host buffers in,
    factors out."""
    torch = be.torch_mod()
    from . import archetypal_analysis as aa
    from . import gpnh_convex_coding as gp
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if args.workload == 'gpnh':
        out = gp._iterate_gpnh_convex_coding(X, Z0, F0, lambda_W=0.0, tolerance=tolerance,
                                             max_iterations=max_iterations, comm=comm)
        cost, n_iter = out[2], out[3]
    else:
        out = aa._iterate_aa(X, Z0, F0, np.ones(F0.shape[0]), tolerance=tolerance,
                             max_iterations=max_iterations,
                             dictionary_solver_kwargs=dict(max_iterations=1), comm=comm,
                             formulation=getattr(args, 'formulation', 'stream'))
        cost, n_iter = out[3], out[4]
    torch.cuda.synchronize()
    return {'seconds': time.perf_counter() - t0, 'iterations': int(n_iter) + 1,
            'tolerance': tolerance, 'stopping_criterion': 'abs_delta_f', 'cost': float(cost)}
