import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'matrix-factorization-case-studies_b200'))
import numpy as np, torch
from convex_dim_red import _backend as be
from convex_dim_red.archetypal_analysis import _AaEngine
from convex_dim_red.gpnh_convex_coding import _GpnhEngine
from convex_dim_red.datasets import synthetic_field
from convex_dim_red import stochastic_matrices as orc
T, d, k = 1620, 44000, 8
X = synthetic_field(T, d, seed=0)
rs = np.random.RandomState(1000)
def tick(msg, t0):
    torch.cuda.synchronize(); t = time.perf_counter(); print('%-28s %.2f ms' % (msg, (t - t0) * 1e3)); return t
for workload in ('gpnh', 'aa', 'gpnh', 'aa'):
    print('==', workload)
    if workload == 'gpnh':
        F0 = np.sqrt(np.abs(X).mean() / k) * np.random.RandomState(0).randn(d, k)
    else:
        F0 = orc.right_stochastic_matrix((k, T), rs)
    Z0 = orc.right_stochastic_matrix((T, k), rs)
    torch.cuda.synchronize(); t = time.perf_counter()
    Xd = be.to_device_padded(X); t = tick('upload X', t)
    if workload == 'gpnh':
        eng = _GpnhEngine(X, Z0, F0, tolerance=0.0, max_iterations=20, require_monotonic_cost_decrease=False, X_device=Xd)
    else:
        eng = _AaEngine(X, Z0, F0, np.ones(k), 'feature', tolerance=0.0, max_iterations=20, require_monotonic_cost_decrease=False, dictionary_solver_kwargs=dict(max_iterations=1), data_device=Xd)
    t = tick('engine ctor', t)
    eng.initial_cost(); t = tick('initial cost', t)
    for i in range(3):
        eng.iteration(); t = tick('eager iteration %d' % i, t)
    g = be.capture_graph(eng.iteration); t = tick('graph capture', t)
    for i in range(3):
        g.replay(); t = tick('replay %d' % i, t)
    st = eng.state.read(); t = tick('state read', t)
    print(st.n_iter, st.cost)
