"""Builds tests/golden/block_small.ddpk.gz from the UNTOUCHED reference (oracle/_ref/block_admm).

BLOCK (examples/BLOCK.h): 9 bodies, 6 tied + 2 frictionless contact interfaces, coarsest mesh
2x2x2 per block, globLeve=1.  Holds every operator after MCONTACT::ESTABLISH, the converged
reference run with the macroscopic problem (muscSett=1) and the first 30 monitor rows of the
reference loop without coarse-space correction (muscSett=0).  Run in the build container."""
import gzip
import json
import os
import subprocess
import sys
import tempfile

here = os.path.dirname(os.path.abspath(__file__))
root = os.path.dirname(os.path.dirname(here))
sys.path.insert(0, os.path.join(root, "ddpca-admm_b200"))
from ddpca_b200 import ddpk  # noqa: E402

ref = os.path.join(root, "oracle", "_ref", "block_admm")
tmp = tempfile.mkdtemp()


def run(args, out):
    # the K-iterations mode ends the process from a watcher thread while the reference loop is still
    # running; on rare occasions that races with the OpenMP runtime's teardown -> retry
    for attempt in range(5):
        try:
            txt = subprocess.check_output([ref, "--glob", "1", "--divi", "2,2,2", "--out", out] + args, cwd=tmp).decode()
            return json.loads(txt.strip().splitlines()[-1])
        except subprocess.CalledProcessError:
            if attempt == 4:
                raise


m1 = run(["--musc", "1"], os.path.join(tmp, "m1.ddpk"))
m0 = run(["--musc", "0", "--ref-iters", "30"], os.path.join(tmp, "m0.ddpk"))
d1 = ddpk.load(os.path.join(tmp, "m1.ddpk"))
d0 = ddpk.load(os.path.join(tmp, "m0.ddpk"))
# the LDLT factors are rebuilt inside the tests (tests/helpers.py) to keep the fixture small
keep = {k: v for k, v in d1.items() if not any(t in k for t in ("inteDiso", "coarSolv_D"))}
keep["ref0.resuMoni"] = d0["ref.resuMoni"]
keep["ref0.resuMoni.shape"] = d0["ref.resuMoni.shape"]
ddpk.save(os.path.join(tmp, "block_small.ddpk"), keep)
with open(os.path.join(tmp, "block_small.ddpk"), "rb") as f, open(os.path.join(here, "block_small.ddpk.gz"), "wb") as g:
    g.write(gzip.compress(f.read(), 9, mtime=0))
for m in (m1, m0):
    for k in list(m):
        if k.endswith("_s"):
            m.pop(k)
json.dump({"musc1": m1, "musc0": m0}, open(os.path.join(here, "block_small.json"), "w"), indent=1)
print(os.path.getsize(os.path.join(here, "block_small.ddpk.gz")), "bytes")
