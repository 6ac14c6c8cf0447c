"""N>1 host logic on CPU (gloo, world_size 2): subdomain partitioning and the three per-iteration
exchanges of the multi-GPU ADMM loop, exercised with the partitioned CPU oracle."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from ddpca_b200 import ddpk
from ddpca_b200.partition import cross_interfaces, partition_bodies
from tests.helpers import have_ref_binary, run_ref_beam_dd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, musc, iters, dump=None):
    out = tempfile.mkdtemp(prefix="ddpca_dist_")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29613", os.path.join(ROOT, "tests", "dist_worker.py"), out, str(musc), str(iters)] + ([dump] if dump else [])
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)
    return [json.load(open(os.path.join(out, f"rank{r}.json"))) for r in range(world)]


def test_partition_balances_and_covers():
    w = [100, 100, 100, 10, 12, 12, 12, 12, 12]
    cb = [[0, 3], [0, 4], [1, 5], [1, 6], [2, 7], [2, 8], [4, 5], [6, 7]]
    for n in (1, 2, 3, 4, 8):
        r = partition_bodies(w, cb, n)
        assert len(r) == len(w) and set(r) <= set(range(n))
        load = [sum(wi for wi, ri in zip(w, r) if ri == k) for k in range(n)]
        assert max(load) <= max(w) + sum(w) / n * 1.1
    assert cross_interfaces(cb, [0] * 9) == []
    assert cross_interfaces(cb, [0, 1, 1, 0, 0, 1, 1, 1, 1]) == [6]


@pytest.mark.parametrize("musc,iters", [(1, 3000), (0, 12)])
def test_two_rank_admm_equals_single_process_oracle(golden_dir, musc, iters):
    from oracle.admm_oracle import AdmmOracle

    res = _run(2, musc, iters)
    d = ddpk.load(os.path.join(golden_dir, "block_small.ddpk.gz"))
    o = AdmmOracle(d)
    o.muscSett = musc
    it = o.run(max_iter=iters)
    assert res[0]["iterNumbReco"] == res[1]["iterNumbReco"] == it      # same stopping decision on every rank
    assert len(set(res[0]["body_rank"])) == 2 and res[0]["cross"]       # the split really is across ranks
    rows = np.array(o.rows)
    for r in res:
        mine = np.array(r["rows"])
        assert mine.shape == rows.shape
        scale = np.abs(rows).max(axis=0, keepdims=True)
        sig = np.abs(rows) > 1e-14 * scale
        assert np.max(np.abs(mine - rows)[sig] / np.abs(rows)[sig]) < 1e-8
        for v, nrm in r["disp_norm"].items():
            assert abs(nrm - np.linalg.norm(o.resuDisp[int(v)])) <= 1e-9 * nrm
    # three all-reduces per iteration with the macroscopic problem, two without
    n_it = len(o.rows)
    assert res[0]["allreduces"] == (3 if musc else 2) * n_it


@pytest.mark.skipif(not have_ref_binary("beam_admm"), reason="oracle/_ref/beam_admm not built")
@pytest.mark.parametrize("musc", [2, 3])
def test_two_rank_admm_with_the_interface_eliminated_coarse_problem(musc):
    """muscSett bit 1 (MCONTACT.h:2575-2607) on the reference's BEAM with 8 subdomains: one more
    all-reduce of a coarse right-hand side per iteration; two ranks == one process == the reference."""
    from oracle.admm_oracle import AdmmOracle

    d, meta = run_ref_beam_dd(1, doma=(8, 1, 1), musc=musc, keep_file=True)
    res = _run(2, musc, 3000, dump=meta["path"])
    o = AdmmOracle(d)
    it = o.run()
    assert it == meta["ref_iterNumbReco"] == res[0]["iterNumbReco"] == res[1]["iterNumbReco"]
    for r in res:
        for v, nrm in r["disp_norm"].items():
            assert abs(nrm - np.linalg.norm(d[f"ref.resuDisp{v}"])) <= 1e-8 * nrm
    n_it = len(o.rows)
    assert res[0]["allreduces"] == (2 + bin(musc).count("1")) * n_it



@pytest.mark.parametrize("world", [2, 3])
def test_pairwise_trace_swap_and_allreduce_over_gloo(world, tmp_path):
    """ddpca_b200.comm.TorchComm (the exchange layer of the multi-GPU loop): ranges of the send buffer go to their peers
    and the same ranges of the receive buffer come back (grouped isend/irecv, no reduction), all-reduce sums."""
    out = str(tmp_path)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(ROOT, "tests", "dist_comm_worker.py"), out]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300)
    for r in range(world):
        assert json.load(open(os.path.join(out, f"comm{r}.json")))["ok"]
