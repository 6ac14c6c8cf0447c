"""GPU parity tests of the ADMM loop (MCONTACT::CONTACT_ANALYSIS) and the device LDLT solve,
against the reference's golden vectors and the CPU oracle (oracle/admm_oracle.py)."""
import json
import os

import numpy as np
import pytest

import ddpca_b200 as dd
from ddpca_b200 import ddpk
from tests.helpers import GOLDEN, dense_ldlt_factor, have_ref_binary, rel, run_ref_beam_dd, run_ref_block

pytestmark = pytest.mark.gpu


def _moni(d, key):
    return d[key].reshape(tuple(int(v) for v in d[key + ".shape"]))


def _rows_close(mine, ref, rtol):
    mine, ref = np.asarray(mine), np.asarray(ref)
    scale = np.abs(ref).max(axis=0, keepdims=True)
    sig = np.abs(ref) > 1e-16 * scale
    return float(np.max(np.abs(mine - ref)[sig] / np.abs(ref)[sig])) < rtol


@pytest.fixture(scope="module")
def block_small():
    d = ddpk.load(os.path.join(GOLDEN, "block_small.ddpk.gz"))
    meta = json.load(open(os.path.join(GOLDEN, "block_small.json")))
    return d, meta


def test_ldlt_solve_matches_direct_solve(block_small):
    d, meta = block_small
    rng = np.random.default_rng(0)
    for name in ("if0.s0.inteMass", "if7.s1.inteMass_pena", "globCoup"):
        m = ddpk.get_csr(d, name)
        perm, L, D = dense_ldlt_factor(m)
        s = dd.DIRE_SOLV(perm, L, D)
        b = rng.standard_normal(m.shape[0])
        x = s.solve(b)
        assert rel(m.to_scipy() @ x, b) < 1e-10
        info = s.info()
        assert info["n"] == m.shape[0] and info["stages_fwd"] >= 1
        s.close()


@pytest.mark.skipif(not have_ref_binary("block_admm"), reason="oracle/_ref/block_admm not built")
def test_ldlt_solve_with_eigen_amd_factor():
    """The factor exactly as Eigen::SimplicialLDLT produced it (AMD ordering) in the reference run."""
    d, meta = run_ref_block(1, divi=(2, 2, 2), musc=1, ref_iters=-1)
    rng = np.random.default_rng(1)
    for name, mat in (("coarSolv_D", "globCoup"), ("if6.s0.inteDiso", "if6.s0.inteMass"), ("if0.s1.inteDiso_pena", "if0.s1.inteMass_pena")):
        s = dd.DIRE_SOLV(d[name + ".perm"], ddpk.get_csr(d, name + ".L"), d[name + ".D"])
        m = ddpk.get_csr(d, mat).to_scipy()
        b = rng.standard_normal(m.shape[0])
        assert rel(m @ s.solve(b), b) < 1e-10
        s.close()


@pytest.mark.parametrize("smoother", [dd.SMOOTH_MC, dd.SMOOTH_LEX])
def test_admm_macroscopic_run_matches_reference(block_small, smoother):
    d, meta = block_small
    mc = dd.MCONTACT.from_ddpk(d, smoother=smoother, factorize=dense_ldlt_factor)
    mc.CONTACT_ANALYSIS()
    assert mc.iterNumbReco == meta["musc1"]["ref_iterNumbReco"] == int(d["ref.iterNumbReco"][0])   # bit-exact count
    disp = mc.resuDisp
    for v in range(mc.nb):
        assert rel(disp[v], d[f"ref.resuDisp{v}"]) < 1e-8
    aux, lagr = mc.inteAuxi, mc.inteLagr
    for ts in range(mc.ni):
        for tv in range(2):
            assert rel(aux[ts][tv], d[f"ref.if{ts}.s{tv}.inteAuxi"]) < 1e-8
            assert rel(lagr[ts][tv], d[f"ref.if{ts}.s{tv}.inteLagr"]) < 1e-6
    for ts in range(mc.ni):
        if mc.fricCoef[ts] == 0.0:
            g, st = mc.inpoGamm(ts)
            p_ref = _moni(d, f"ref.resuCont{ts}")[:, 0]
            assert rel(g, p_ref) < 1e-8                      # contact pressure
            assert ((g > 0) == (p_ref > 0)).all()            # active contact set, bit-exact
    assert mc.launch_count() > 0 and mc.cg_iters > 0
    mc.close()


def test_admm_trajectory_without_coarse_space_matches_reference(block_small):
    d, meta = block_small
    ref = _moni(d, "ref0.resuMoni")
    mc = dd.MCONTACT.from_ddpk(d, muscSett=0, factorize=dense_ldlt_factor)
    mc.CONTACT_ANALYSIS(maxiIter=ref.shape[0])
    assert len(mc.resuMoni) == ref.shape[0]
    assert _rows_close(mc.resuMoni, ref, 1e-6)
    mc.close()


def test_admm_matches_cpu_oracle_state_by_state(block_small):
    from oracle.admm_oracle import AdmmOracle

    d, meta = block_small
    o = AdmmOracle(d)
    o.muscSett = 0
    mc = dd.MCONTACT.from_ddpk(d, muscSett=0, factorize=dense_ldlt_factor)
    for tc in range(5):
        o.step(tc)
        mc.step(tc)
    disp = mc.resuDisp
    for v in range(mc.nb):
        assert rel(disp[v], o.resuDisp[v]) < 1e-9
    lagr = mc.inteLagr
    for ts in range(mc.ni):
        for tv in range(2):
            assert rel(lagr[ts][tv], o.inteLagr[ts][tv]) < 1e-8
        g, st = mc.inpoGamm(ts)
        assert rel(g, o.inpoGamm[ts]) < 1e-8
    mc.close()


@pytest.mark.skipif(not have_ref_binary("block_admm"), reason="oracle/_ref/block_admm not built")
def test_block_g2_small_25_iterations_against_reference_run_here():
    """BLOCK, coarsest mesh 2x2x2, globLeve=2, macroscopic problem: the reference needs 25 ADMM
    iterations; counts must match exactly, states to 1e-8 (uses Eigen's own LDLT factors)."""
    d, meta = run_ref_block(2, divi=(2, 2, 2), musc=1, ref_iters=0)
    mc = dd.MCONTACT.from_ddpk(d)
    mc.CONTACT_ANALYSIS()
    assert mc.iterNumbReco == meta["ref_iterNumbReco"]
    disp = mc.resuDisp
    for v in range(mc.nb):
        assert rel(disp[v], d[f"ref.resuDisp{v}"]) < 1e-8
    ref = _moni(d, "ref.resuMoni")
    assert len(mc.resuMoni) == ref.shape[0]
    assert _rows_close(np.array(mc.resuMoni)[:, -2:], ref[:, -2:], 1e-5)   # Cvalu, Ccrit per iteration
    for ts in range(mc.ni):
        if mc.fricCoef[ts] == 0.0:
            g, st = mc.inpoGamm(ts)
            p_ref = _moni(d, f"ref.resuCont{ts}")[:, 0]
            assert rel(g, p_ref) < 1e-8 and ((g > 0) == (p_ref > 0)).all()
    mc.close()


@pytest.mark.skipif(not have_ref_binary("beam_admm"), reason="oracle/_ref/beam_admm not built")
@pytest.mark.parametrize("musc", [2, 3])
def test_beam_dd_with_the_interface_eliminated_coarse_problem(musc):
    """muscSett bit 1 (operators of MCONTACT::MULTISCALE_1, loop block MCONTACT.h:2575-2607), alone and with
    the macroscopic problem, on the reference's BEAM with 8 subdomains: the reference runs on this box's
    CPU, the device loop must stop at the same iteration with the same state."""
    d, meta = run_ref_beam_dd(1, doma=(8, 1, 1), musc=musc)
    mc = dd.MCONTACT.from_ddpk(d)
    assert mc.muscSett == musc
    mc.CONTACT_ANALYSIS()
    assert mc.iterNumbReco == meta["ref_iterNumbReco"]
    disp = mc.resuDisp
    for v in range(mc.nb):
        assert rel(disp[v], d[f"ref.resuDisp{v}"]) < 1e-8
    mc.close()


def test_admm_loop_is_bit_reproducible(block_small):
    """Fixed-order reductions everywhere, body solves on their own streams, the macroscopic solve with its dense
    tail: three fresh handles must produce identical monitor rows and displacements, bit for bit."""
    d, meta = block_small
    outs = []
    for rep in range(3):
        mc = dd.MCONTACT.from_ddpk(d, factorize=dense_ldlt_factor)
        rows = [mc.step(tc).copy() for tc in range(4)]
        outs.append((np.concatenate(mc.resuDisp), np.array(rows)))
        mc.close()
    for o in outs[1:]:
        assert np.array_equal(outs[0][0], o[0]) and np.array_equal(outs[0][1], o[1])


def test_macroscopic_problem_through_its_own_mgpis_hierarchy(block_small):
    """MCONTACT.h:2553-2562: beyond DIRE_MAXI rows the reference solves the macroscopic problem with
    MCONTACT's own hierarchy, mgpi.CG_SOLV(1, globForc, globSolu), instead of the factor coarSolv_D.
    Here the fixture's globCoup is wrapped in a one-level MGPIS (its level-0 direct solve makes the
    preconditioner exact), so the ADMM iterates must agree with those of the factor path."""
    d, meta = block_small
    mg = dd.MGPIS.from_hierarchy([ddpk.get_csr(d, "globCoup")], [])
    a = dd.MCONTACT.from_ddpk(d, factorize=dense_ldlt_factor)
    b = dd.MCONTACT.from_ddpk(d, factorize=dense_ldlt_factor, macro_mgpis=mg)
    for tc in range(5):
        ra, rb = a.step(tc), b.step(tc)
    da, db = a.resuDisp, b.resuDisp
    for v in range(a.nb):
        assert rel(db[v], da[v]) < 1e-8
    la, lb = a.inteLagr, b.inteLagr
    for ts in range(a.ni):
        for tv in range(2):
            assert rel(lb[ts][tv], la[ts][tv]) < 1e-8
    a.close()
    b.close()


def test_coulomb_friction_projection_matches_oracle(block_small):
    """The reference's frictional branch (MCONTACT.h:2648-2668: normal clamp, Coulomb cone, status
    0 open / 1 slide / 2 stick) is only reached by DEHW.  Here two vector-valued (d = 3) interfaces of
    the BLOCK fixture are declared frictional in BOTH the oracle and the device loop: same operators,
    same algebra, so states and status codes must agree."""
    from oracle.admm_oracle import AdmmOracle

    d, meta = block_small
    d = dict(d)
    for ts, mu in ((0, 0.3), (3, 0.05)):
        assert float(d[f"if{ts}.fricCoef"][0]) < 0.0          # tied: three components per point
        d[f"if{ts}.fricCoef"] = np.array([mu])
    o = AdmmOracle(d)
    o.muscSett = 0
    mc = dd.MCONTACT.from_ddpk(d, muscSett=0, factorize=dense_ldlt_factor)
    for tc in range(6):
        o.step(tc)
        mc.step(tc)
    seen = set()
    for ts in (0, 3):
        g, st = mc.inpoGamm(ts)
        assert rel(g, o.inpoGamm[ts]) < 1e-8
        assert np.array_equal(st[1::3], o.fricStat[ts][1::3])  # friction status per integration point
        seen |= set(st[1::3].tolist())
    assert len(seen) >= 2                                      # more than one branch was exercised
    disp = mc.resuDisp
    for v in range(mc.nb):
        assert rel(disp[v], o.resuDisp[v]) < 1e-8
    mc.close()


@pytest.mark.skipif(not have_ref_binary("beam_admm"), reason="oracle/_ref/beam_admm not built")
def test_beam_dd_8_subdomains_against_reference_run_here():
    """BASELINE.json config "BEAM ... 8 subdomains": examples/BEAM.h with domaNumb = 8x1x1, tied
    (vector-valued, d = 3) interfaces only, macroscopic problem; the reference runs on this box's CPU."""
    d, meta = run_ref_beam_dd(1, doma=(8, 1, 1), musc=1)
    assert meta["bodies"] == 8 and meta["interfaces"] == 7
    mc = dd.MCONTACT.from_ddpk(d)
    mc.CONTACT_ANALYSIS()
    assert mc.iterNumbReco == meta["ref_iterNumbReco"]
    disp = mc.resuDisp
    for v in range(mc.nb):
        assert rel(disp[v], d[f"ref.resuDisp{v}"]) < 1e-8
    ref = _moni(d, "ref.resuMoni")
    mine = np.array(mc.resuMoni)
    assert np.allclose(mine[:, -1], ref[:, -1], rtol=1e-8, atol=0)          # Ccrit: squared norms of the state
    big = ref[:, -2] > 1e-9 * ref[:, -1]                                     # Cvalu above round-off level
    assert np.allclose(mine[big, -2], ref[big, -2], rtol=1e-5, atol=0)
    mc.close()


def test_interface_mass_solves_by_batched_jacobi_pcg_match_the_factor_path(block_small):
    """Sides without a factor (the reference: inteMass of DIRE_MAXI rows or more, Eigen CG with the diagonal
    preconditioner, MCONTACT.h:2678-2683, :2698-2703) are solved by ONE batched Jacobi-PCG per update.  Forced here
    for every side of the BLOCK fixture: same ADMM iteration count as the reference, state to 1e-8."""
    d, meta = block_small
    mc = dd.MCONTACT.from_ddpk(d, factorize=dense_ldlt_factor, iterative_above=0)
    mc.CONTACT_ANALYSIS()
    assert mc.iterNumbReco == int(d["ref.iterNumbReco"][0])
    disp = mc.resuDisp
    for v in range(mc.nb):
        assert rel(disp[v], d[f"ref.resuDisp{v}"]) < 1e-8
    aux = mc.inteAuxi
    for ts in range(mc.ni):
        for tv in range(2):
            assert rel(aux[ts][tv], d[f"ref.if{ts}.s{tv}.inteAuxi"]) < 1e-8
    mc.close()


def test_interface_eliminated_problem_through_its_own_mgpis_hierarchy():
    """MCONTACT.h:2590-2595: beyond DIRE_MAXI rows the interface-eliminated coarse problem is solved by mgpi_1.CG_SOLV.
    As for the macroscopic problem, globCoup_1 of the BEAM DD run is wrapped in a one-level MGPIS (exact level-0 solve):
    the iterates must agree with the factor path."""
    import ctypes as C

    from ddpca_b200.lib import check, load_library

    if not have_ref_binary("beam_admm"):
        pytest.skip("oracle/_ref/beam_admm not built")
    d, meta = run_ref_beam_dd(1, doma=(8, 1, 1), musc=2)
    a = dd.MCONTACT.from_ddpk(d)
    a.CONTACT_ANALYSIS()
    assert a.iterNumbReco == meta["ref_iterNumbReco"]
    da = a.resuDisp
    a.close()
    # second handle: same upload, then the coarse solver is replaced before finalize is not possible from Python's
    # from_ddpk, so the C entry is exercised directly on a handle built step by step below
    mg = dd.MGPIS.from_hierarchy([ddpk.get_csr(d, "globCoup_1")], [])
    b = dd.MCONTACT.from_ddpk(d, macro1_mgpis=mg)
    b.CONTACT_ANALYSIS()
    assert b.iterNumbReco == meta["ref_iterNumbReco"]
    for v in range(b.nb):
        assert rel(b.resuDisp[v], da[v]) < 1e-8
    b.close()
