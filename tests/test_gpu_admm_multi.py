"""Multi-GPU ADMM (one process per GPU, NCCL): the partitioned loop must reproduce the single-GPU
loop and the reference -- same iteration count, monitor rows to 1e-8.  Needs >= 2 GPUs."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import ddpca_b200 as dd
from ddpca_b200 import ddpk
from tests.helpers import GOLDEN, have_ref_binary, run_ref_beam_dd

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        return dd.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("musc,iters", [(1, 3000), (0, 15)])
def test_two_gpu_admm_matches_reference(musc, iters):
    path = os.path.join(GOLDEN, "block_small.ddpk.gz")
    out = tempfile.mkdtemp(prefix="ddpca_gpu_dist_")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(ROOT, "tests", "gpu_dist_worker.py"), path, out, str(musc), str(iters)]
    subprocess.check_call(cmd, timeout=900)
    res = [json.load(open(os.path.join(out, f"rank{r}.json"))) for r in range(2)]
    d = ddpk.load(path)
    key = "ref.resuMoni" if musc else "ref0.resuMoni"
    ref = d[key].reshape(tuple(int(v) for v in d[key + ".shape"]))[:iters]
    assert res[0]["iterNumbReco"] == res[1]["iterNumbReco"]
    if musc:
        assert res[0]["iterNumbReco"] == int(d["ref.iterNumbReco"][0])
    assert len(set(res[0]["body_rank"])) == 2
    for r in res:
        rows = np.array(r["rows"])
        assert rows.shape == ref.shape
        scale = np.abs(ref).max(axis=0, keepdims=True)
        sig = np.abs(ref) > 1e-14 * scale
        if musc == 0:
            assert np.max(np.abs(rows - ref)[sig] / np.abs(ref)[sig]) < 1e-6
        for v, nrm in r["disp_norm"].items():
            if musc:
                assert abs(nrm - np.linalg.norm(d[f"ref.resuDisp{v}"])) <= 1e-8 * nrm
        assert r["launches"] > 0


@pytest.mark.skipif(_ngpu() < 2 or not have_ref_binary("beam_admm"), reason="needs 2 GPUs and oracle/_ref/beam_admm")
@pytest.mark.parametrize("musc", [1, 2, 3])
def test_two_gpu_beam_dd_8_subdomains(musc):
    """8 BEAM subdomains split 4 + 4 over two GPUs (NCCL: all-reduce of the coarse right-hand sides and MONITOR sums,
    pairwise swap of the interface traces), with the macroscopic problem, the interface-eliminated coarse problem and
    both: same iteration count and displacements as the reference."""
    d, meta = run_ref_beam_dd(1, doma=(8, 1, 1), musc=musc, keep_file=True)
    out = tempfile.mkdtemp(prefix="ddpca_gpu_dist_")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29619", os.path.join(ROOT, "tests", "gpu_dist_worker.py"), meta["path"], out, str(musc), "3000"]
    subprocess.check_call(cmd, timeout=900)
    res = [json.load(open(os.path.join(out, f"rank{r}.json"))) for r in range(2)]
    assert res[0]["iterNumbReco"] == res[1]["iterNumbReco"] == meta["ref_iterNumbReco"]
    assert sorted(res[0]["body_rank"].count(r) for r in (0, 1)) == [4, 4]
    for r in res:
        for v, nrm in r["disp_norm"].items():
            assert abs(nrm - np.linalg.norm(d[f"ref.resuDisp{v}"])) <= 1e-8 * nrm
