"""Host logic (no GPU): the stopping logic of the device ADMM loop lives on the host --
`ddpca_b200.MCONTACT.MONITOR` restates MCONTACT::MONITOR (MCONTACT.h:2725-2845) on the sums a device
iteration returns.  Replaying the reference's own resuMoni.txt rows (tests/golden/block_small) through it
must reproduce the reference's decisions: stop exactly at its last iteration, and the same MULT_MAXI
switch (oscillation test) as the pinned oracle takes on the non-converged trajectory."""
import json
import os

import numpy as np

import ddpca_b200 as dd
from ddpca_b200 import ddpk

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rows(d, key):
    return d[key].reshape(tuple(int(v) for v in d[key + ".shape"]))


def _mirror(d):
    mc = dd.MCONTACT()   # no device handle: only the host-side state of the loop
    mc.nb, mc.ni = int(d["nbody"][0]), int(d["niface"][0])
    mc.moniReco = [[0.0] * 10 for _ in range(mc.nb + 4 * mc.ni)]
    return mc


def test_monitor_mirror_stops_where_the_reference_stops():
    d = ddpk.load(os.path.join(GOLDEN, "block_small.ddpk.gz"))
    ref = _rows(d, "ref.resuMoni")
    assert ref.shape[1] == 2 * int(d["nbody"][0]) + 8 * int(d["niface"][0]) + 2
    mc = _mirror(d)
    decisions = [mc.MONITOR(tc, ref[tc]) for tc in range(ref.shape[0])]
    assert decisions == [-1] * (ref.shape[0] - 1) + [1]
    assert ref.shape[0] - 1 == int(d["ref.iterNumbReco"][0])


def test_monitor_mirror_follows_the_oracle_on_a_non_converged_trajectory():
    from oracle.admm_oracle import AdmmOracle

    d = ddpk.load(os.path.join(GOLDEN, "block_small.ddpk.gz"))
    ref0 = _rows(d, "ref0.resuMoni")   # muscSett = 0: 30 rows, far from converged
    mc = _mirror(d)
    assert all(mc.MONITOR(tc, ref0[tc]) == -1 for tc in range(ref0.shape[0]))
    o = AdmmOracle(d)
    o.muscSett = 0
    o.run(max_iter=ref0.shape[0])
    assert mc.MULT_MAXI == o.MULT_MAXI
    # the ring buffers hold the last ten squared increments of every monitored quantity
    nb = mc.nb
    for v in range(nb):
        assert np.allclose(sorted(mc.moniReco[v]), sorted(ref0[-10:, 2 * v]), rtol=0, atol=0)
