import os, sys, json, numpy as np
ROOT="/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ddpca-admm_b200"))
import ddpca_b200 as dd
from ddpca_b200 import ddpk
from tests.helpers import dense_ldlt_factor
d = ddpk.load(os.path.join(ROOT, "tests/golden/block_small.ddpk.gz"))
outs = []
for rep in range(3):
    mc = dd.MCONTACT.from_ddpk(d, factorize=dense_ldlt_factor)
    rows = [mc.step(tc).copy() for tc in range(4)]
    outs.append((np.concatenate(mc.resuDisp), np.array(rows)))
    mc.close()
same = all(np.array_equal(outs[0][0], o[0]) and np.array_equal(outs[0][1], o[1]) for o in outs[1:])
print("ADMM bit-reproducible:", same, "max diff", max(float(np.abs(outs[0][0]-o[0]).max()) for o in outs[1:]))
