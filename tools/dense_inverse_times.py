"""Set-up cost of the dense SPD inversions (DIRE_SOLV.dense -> blocked Gauss-Jordan, kernels.cuh k_bgj_*) on banded
test matrices of the sizes the bench workload inverts.  Run with DDPCA_VERBOSE=1 for the per-kernel-class device times."""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ddpca-admm_b200"))
import ddpca_b200 as dd  # noqa: E402
from ddpca_b200 import ddpk  # noqa: E402

for n in [int(a) for a in sys.argv[1:]] or [1683, 6579, 12288]:
    k = 12
    M = sp.diags([np.full(n - abs(o), 1.0 / (1 + abs(o))) for o in range(-k, k + 1)], list(range(-k, k + 1)), format="csr") + 4.0 * sp.identity(n, format="csr")
    M = M.tocsr()
    M.sort_indices()
    A = ddpk.Csr.from_scipy(M)
    for rep in range(2):
        t0 = time.time()
        s = dd.DIRE_SOLV.dense(A)
        t1 = time.time()
        b = np.ones(n)
        x = s.solve(b)
        err = np.linalg.norm(M @ x - b) / np.linalg.norm(b)
        s.close()
        print(f"n {n} rep {rep}: create {1e3 * (t1 - t0):.1f} ms  ({2.0 * n ** 3 / (t1 - t0) / 1e12:.2f} TFLOP/s incl. everything)  residual {err:.1e}", flush=True)
