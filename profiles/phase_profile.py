"""Phase time stamps of gpnh_weights_fused_kernel (profiling build only:
CDR_NVCC_EXTRA=-DCDR_PROFILE_PHASES bash .../csrc/build.sh).  Runs the GPNH bench workload
eagerly for a few iterations and prints, for the last one, when each warp passed the phase
marks relative to the first warp's start."""
import ctypes
import os
import sys

import numpy as np

os.environ['CDR_NO_CUDA_GRAPH'] = '1'
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', 'matrix-factorization-case-studies_b200'))
sys.path.insert(0, os.path.join(HERE, '..'))
import torch                                                  # noqa: E402
from convex_dim_red import _backend as be                     # noqa: E402
from convex_dim_red.datasets import synthetic_field           # noqa: E402
from convex_dim_red.stochastic_matrices import right_stochastic_matrix   # noqa: E402
import bench_harness as bh                                    # noqa: E402

T, d, k = 1620, 44000, 8
X = synthetic_field(T, d, seed=0)
Z0 = right_stochastic_matrix((T, k), np.random.RandomState(1000))
W0 = np.sqrt(0.4 / k) * np.random.RandomState(0).randn(d, k)
Xd = be.to_device_padded(X)
eng = bh.make_engine('gpnh', X, Z0, W0, Xd=Xd)
lib = be.library()
lib.cdr_debug_phase_read.argtypes = [ctypes.c_void_p]
lib.cdr_debug_phase_read.restype = ctypes.c_int
eng.initial_cost()
done = 0
for n_it in (6, 26):
    while done < n_it:
        eng.iteration()
        done += 1
    torch.cuda.synchronize()
    buf = np.zeros(8 * 4096, dtype=np.uint64)
    assert lib.cdr_debug_phase_read(buf.ctypes.data) == 0
    ph = buf.reshape(8, 4096)
    nw = T
    t0 = ph[0, :nw].min()
    rel = lambda a: (a.astype(np.int64) - np.int64(t0)) / 1e3
    print('--- outer iteration', done)
    for i, name in enumerate(('start', 'after partial sums', 'after qp', 'after statistics')):
        r = rel(ph[i, :nw])
        print('%-20s min %7.2f  p50 %7.2f  p90 %7.2f  p99 %7.2f  max %7.2f us' %
              (name, r.min(), np.percentile(r, 50), np.percentile(r, 90), np.percentile(r, 99), r.max()))
    last = ph[5, :].argmax()
    print('last CTA: final sum done %.2f, end %.2f us' % (rel(ph[4, last:last + 1])[0], rel(ph[5, last:last + 1])[0]))
    tm = rel(ph[4, 4000:4005])
    print('tail marks: enter %.2f, partials summed %.2f, cost checks done %.2f, before solve %.2f, end %.2f us' % tuple(tm))
    qp = (ph[2, :nw].astype(np.int64) - ph[1, :nw].astype(np.int64)) / 1e3
    ni, nf = ph[6, :nw].astype(np.int64), ph[7, :nw].astype(np.int64)
    print('iterations per sample: mean %.1f, max %d; function evaluations: mean %.1f, max %d'
          % ((ni + 1).mean(), (ni + 1).max(), nf.mean(), nf.max()))
    order = np.argsort(qp)[::-1][:8]
    print('slowest samples: qp us', np.round(qp[order], 2), 'n_iter', ni[order], 'n_feval', nf[order])
    A = np.stack([ni + 1, nf, np.ones(nw)], axis=1).astype(float)
    coef = np.linalg.lstsq(A, qp, rcond=None)[0]
    print('least squares: qp_us = %.3f * iterations + %.4f * fevals + %.2f' % tuple(coef))
