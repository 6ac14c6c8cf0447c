"""The non-symmetric variant of the condensed
dual-mortar system in tests/golden/block_lagrange.ddpk.gz (oracle/ref_drivers/lagrange_tap.h, SKEW_VARIANT; pinned on
the CPU by tests/test_oracle_golden.py::test_non_symmetric_system_follows_the_reference) on the device:

  python tools/skew_check.py

Products, sweeps and transfers do not assume symmetry, so BiCGSTAB_SOLV converges to the reference's solution; the
V-cycle differs from the reference's in the level-0 solve only (the reference factorises the lower triangle of
consStif[0], the device inverts the whole block and symmetrises the inverse).  Run once on a B200
(profiles/r2/r2_skew_check.log): LEX 10 iterations = reference, solution 5e-15; MC 11 iterations, 1.4e-15; V-cycle
output 1.55 away from the reference's in relative norm -- exactly what the CPU oracle gives when its level 0 is
replaced by the symmetrised inverse of the whole block (1.545, same 10 iterations).
Prints what it finds; exits non-zero if the solution is off."""
import os
import sys

import numpy as np  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ddpca-admm_b200"))
import ddpca_b200 as dd  # noqa: E402
from tests.helpers import load_golden, rel  # noqa: E402
from tests.test_oracle_golden import skewed_hierarchy  # noqa: E402

d, meta, A, P = load_golden("block_lagrange")
As = skewed_hierarchy(d, A)
bad = False
for mode, name in ((dd.SMOOTH_LEX, "LEX"), (dd.SMOOTH_MC, "MC")):
    mg = dd.MGPIS.from_hierarchy(As, P, smoother=mode)
    z = mg.MULT_VCYC(len(As) - 1, d["F"])
    x = mg.BiCGSTAB_SOLV(1, d["F"])
    err = rel(x, d["skew.U"])
    print(f"{name}: V-cycle vs reference {rel(z, d['skew.vcyc_of_F']):.2e}; BiCGSTAB iterations {mg.last_iterNumb} "
          f"(reference {int(d['skew.bicgstab_iters'][0])}), residual {mg.last_resid:.2e} / {mg.last_tol:.2e}, solution vs reference {err:.2e}")
    bad |= not (err < 1e-8)
    mg.close()
sys.exit(1 if bad else 0)
