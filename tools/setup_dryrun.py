"""Host half of the batched hierarchy set-up on a DDPK dump, without a GPU (ddpca_mg_setup_dryrun): where the host
time of `finalize` goes and how much HBM the hierarchy takes.  usage: python tools/setup_dryrun.py dump.ddpk [repeat]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ddpca-admm_b200"))
import ddpca_b200 as dd  # noqa: E402
from ddpca_b200 import ddpk  # noqa: E402

d = ddpk.load(sys.argv[1])
nb = int(d["nbody"][0])
hiers = []
for v in range(nb):
    p = f"body{v}."
    L = int(d[p + "maxiLeve"][0])
    A = [ddpk.get_csr(d, p + f"consStif{l}") for l in range(L + 1)]
    P = [ddpk.get_csr(d, p + f"realProl{l}") for l in range(L)]
    hiers.append((A, P))
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    t0 = time.time()
    out = dd.setup_dryrun(hiers)
    out["wall_s"] = round(time.time() - t0, 3)
    out["seconds"] = {k: round(v, 3) for k, v in out["seconds"].items()}
    out["rows"] = sum(A[-1].shape[0] for A, _ in hiers)
    print(json.dumps(out))
