"""Phase time stamps of aa_head_kernel (profiling build only, see phase_profile.py)."""
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

T, d, k = int(os.environ.get('PROFILE_T', '1620')), 44000, 8
X = synthetic_field(T, d, seed=0)
Z0 = right_stochastic_matrix((T, k), np.random.RandomState(1000))
C0 = right_stochastic_matrix((k, T), np.random.RandomState(7))
Xd = be.to_device_padded(X)
eng = bh.make_engine('aa', X, Z0, C0, Xd=Xd)
lib = be.library()
lib.cdr_debug_head_read.argtypes = [ctypes.c_void_p]
lib.cdr_debug_head_read.restype = ctypes.c_int
eng.initial_cost()
names = ['start', 'row loaded', 'threshold 1', 'x, a0', 'gradient', 'threshold 2', 'row max',
         'grid barrier', 'x - alpha g', 'threshold 3', 'end']
done = 0
for n_it in (6, 26):
    while done < n_it:
        eng.iteration()
        done += 1
    torch.cuda.synchronize()
    buf = np.zeros(16 * 16, dtype=np.uint64)
    assert lib.cdr_debug_head_read(buf.ctypes.data) == 0
    ph = buf.reshape(16, 16)[:k, :11].astype(np.int64)
    rel = (ph - ph[:, 0].min()) / 1e3
    print('--- outer iteration', done, '(us since the first CTA started; min / max over the row CTAs)')
    for i, name in enumerate(names):
        print('%-14s %7.2f %7.2f' % (name, rel[:, i].min(), rel[:, i].max()))
