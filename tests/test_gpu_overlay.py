"""The drop-in itself: reference sources compiled UNCHANGED against the C++ overlay
(ddpca-admm_b200/host/MGPIS.h, MCONTACT_B200.h; built by ddpca-admm_b200/host/Makefile in the build
container) run here on the GPU and are compared with the pure-reference binaries (oracle/_ref)."""
import json
import os
import subprocess
import tempfile

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "ddpca-admm_b200", "host", "_bin")
REF = os.path.join(ROOT, "oracle", "_ref")


def _run(exe, args, env=None):
    tmp = tempfile.mkdtemp(prefix="ddpca_overlay_")
    e = dict(os.environ)
    e.update(env or {})
    txt = subprocess.check_output([exe] + args, cwd=tmp, timeout=900, env=e).decode()
    out = json.loads(txt.strip().splitlines()[-1])
    moni = os.path.join(tmp, "Beam", "resuMoni.txt")
    if os.path.exists(moni):
        out["resuMoni"] = open(moni).read()
    out["cwd"] = tmp
    return out


def _ngpu():
    try:
        import ddpca_b200 as dd

        return dd.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(not (os.access(os.path.join(BIN, "beam_nodd_b200"), os.X_OK) and os.access(os.path.join(REF, "beam_nodd"), os.X_OK)),
                    reason="overlay / reference binaries not built")
def test_beam_example_through_the_mgpis_overlay():
    """examples/BEAM.h:403-410 (SOLVE_NODD path): same driver source, MGPIS swapped for the overlay."""
    args = ["--glob", "2", "--divi", "16,2,2"]
    ref = _run(os.path.join(REF, "beam_nodd"), args)
    gpu = _run(os.path.join(BIN, "beam_nodd_b200"), args)
    assert gpu["levels"] == ref["levels"]
    assert abs(gpu["x_norm"] - ref["x_norm"]) <= 1e-8 * ref["x_norm"]
    assert abs(gpu["x_maxabs"] - ref["x_maxabs"]) <= 1e-8 * ref["x_maxabs"]
    assert gpu["cg_mg_iters"] <= 2 * ref["cg_mg_iters"]          # multicolour ordering: other inner count
    assert gpu["true_resid"] <= 10 * ref["true_resid"]


@pytest.mark.skipif(not (os.access(os.path.join(BIN, "block_b200"), os.X_OK) and os.access(os.path.join(REF, "block_admm"), os.X_OK)),
                    reason="overlay / reference binaries not built")
def test_block_example_through_the_contact_analysis_overlay():
    """examples/BLOCK.h:510-722 unchanged; MCONTACT::CONTACT_ANALYSIS replaced by DDPCA_CONTACT_ANALYSIS."""
    args = ["--glob", "2", "--divi", "2,2,2", "--musc", "1"]
    ref = _run(os.path.join(REF, "block_admm"), args)
    gpu = _run(os.path.join(BIN, "block_b200"), args)
    assert gpu["error"] is False
    assert gpu["iterNumbReco"] == ref["ref_iterNumbReco"]        # ADMM iteration count: bit-exact
    for a, b in zip(gpu["disp_norm"], ref["ref_disp_norm"]):
        assert abs(a - b) <= 1e-8 * b


@pytest.mark.skipif(not (os.access(os.path.join(BIN, "beam_dd_b200"), os.X_OK) and os.access(os.path.join(REF, "beam_admm"), os.X_OK)),
                    reason="overlay / reference binaries not built")
@pytest.mark.parametrize("musc", ["1", "2", "3"])
def test_beam_dd_example_through_the_overlay(musc):
    """examples/BEAM.h:424-609 (SOLVE_DD, 8 subdomains) unchanged, ADMM loop and MG-PCG on the GPU; with the
    macroscopic problem (muscSett 1), the interface-eliminated coarse problem (2, MCONTACT.h:2575-2607) and both."""
    args = ["--glob", "1", "--doma", "8,1,1", "--musc", musc]
    ref = _run(os.path.join(REF, "beam_admm"), args)
    gpu = _run(os.path.join(BIN, "beam_dd_b200"), args)
    assert gpu["error"] is False
    assert gpu["iterNumbReco"] == ref["ref_iterNumbReco"]
    for a, b in zip(gpu["disp_norm"], ref["ref_disp_norm"]):
        assert abs(a - b) <= 1e-8 * b


@pytest.mark.skipif(not (os.access(os.path.join(BIN, "block_lagrange_b200"), os.X_OK) and os.access(os.path.join(REF, "block_lagrange"), os.X_OK)),
                    reason="overlay / reference binaries not built")
def test_block_example_on_the_dual_mortar_path_through_the_overlay():
    """examples/BLOCK.cpp:96-102 (menu 3) -> SOLVE(2) -> MCONTACT::LAGRANGE(1)
    (MCONTACT.h:2847-3701), the reference's host code unchanged: boundary-consistent dual mortar operators, active-set
    loop, condensation and the hierarchy rebuilt for the condensed system stay on the host; the solve of every
    active-set step, `mgpi.ESTABLISH(); mgpi.BiCGSTAB_SOLV(1, F, U_1)` (:3561-3562), runs on the device through the
    MGPIS overlay.  Compared with the pure-reference binary of the same source: active-set steps, displacements of all
    nine bodies, and the multipliers the reference writes to resuLagr_*.txt (:3618-3635), 1e-8."""
    import numpy as np

    args = ["--glob", "2", "--divi", "2,2,2"]
    ref = _run(os.path.join(REF, "block_lagrange"), args)
    gpu = _run(os.path.join(BIN, "block_lagrange_b200"), args)
    assert gpu["impl"] == "b200" and ref["impl"] == "reference"
    assert gpu["error"] is False
    assert gpu["active_set_steps"] == ref["active_set_steps"] >= 1
    for it_g, it_r in zip(gpu["bicgstab_iters"], ref["bicgstab_iters"]):
        assert 1 <= it_g <= 2 * it_r                       # multicolour ordering: other inner count
    for a, b in zip(gpu["disp_norm"], ref["disp_norm"]):
        assert abs(a - b) <= 1e-8 * b
    nlagr = 0
    for name in sorted(os.listdir(os.path.join(ref["cwd"], "Block"))):
        if not (name.startswith("resuLagr_") or name.startswith("resuDisp_")):
            continue
        a = np.loadtxt(os.path.join(ref["cwd"], "Block", name), ndmin=2)
        b = np.loadtxt(os.path.join(gpu["cwd"], "Block", name), ndmin=2)
        assert a.shape == b.shape
        if a.size:
            assert np.linalg.norm(a - b) <= 1e-8 * np.linalg.norm(a), name
            nlagr += name.startswith("resuLagr_")
    assert nlagr >= 2                                      # the two contact interfaces carry pressure
    # the patch test's analytic answer: uniform contact pressure = the applied 1e7 Pa (examples/BLOCK.h:46)
    for ts in (0, 1):
        lam = np.loadtxt(os.path.join(gpu["cwd"], "Block", f"resuLagr_{ts}.txt"), ndmin=2)
        assert np.all(lam[:, 1] == 2) and np.abs(lam[:, 2] - 1.0e7).max() <= 1e-8 * 1.0e7


@pytest.mark.skipif(os.environ.get("DDPCA_SLOW_TESTS") != "1", reason="100 s of reference host code per run: DDPCA_SLOW_TESTS=1 (tools/lagrange_bench.py runs the same check)")
@pytest.mark.skipif(not os.access(os.path.join(BIN, "cylinder_lagrange_b200"), os.X_OK), reason="overlay binaries not built")
def test_cylinder_example_on_the_dual_mortar_path_through_the_overlay():
    """examples/CYLINDER.cpp:85-90 (menu 3): CYLINDER_1, Hertzian contact, MCONTACT::LAGRANGE(1) with a changing active
    set (two steps at locaLeve 5, 504 036 rows, 1 026 constraints change status after the first), every step's
    BiCGSTAB_SOLV on the device; against the reference's own run (tests/golden/cylinder_lagrange.json, made by
    tests/golden/make_lagrange_golden.py): same active-set history, displacements and contact tractions to 1e-8."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import lagrange_bench as lb

    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "cylinder_lagrange.json")))
    gpu = lb.run(os.path.join(BIN, "cylinder_lagrange_b200"), ["--loca", "5"], "Cylinder")
    assert lb.compare(gpu, ref)["failures"] == []


@pytest.mark.skipif(_ngpu() < 2, reason="needs at least 2 GPUs")
@pytest.mark.skipif(not os.access(os.path.join(BIN, "beam_dd_b200"), os.X_OK), reason="overlay binaries not built")
@pytest.mark.parametrize("musc", ["1", "3"])
def test_beam_dd_example_on_several_gpus_in_one_process(musc):
    """DDPCA_DEVICES=0,1[,2,3]: the same unchanged example, its 8 subdomains bin-packed over the devices of ONE
    process (ddpca_admm_group_*: peer-copy exchanges ordered by events).  Same iteration count as on one device;
    per-body arithmetic does not depend on the batch a body runs in, so the displacements agree to round-off of the
    coarse right-hand side's partial sums."""
    args = ["--glob", "1", "--doma", "8,1,1", "--musc", musc]
    one = _run(os.path.join(BIN, "beam_dd_b200"), args, {"DDPCA_DEVICES": "0"})
    n = min(_ngpu(), 4)
    many = _run(os.path.join(BIN, "beam_dd_b200"), args, {"DDPCA_DEVICES": ",".join(str(k) for k in range(n))})
    assert one["error"] is False and many["error"] is False
    assert many["iterNumbReco"] == one["iterNumbReco"]
    for a, b in zip(many["disp_norm"], one["disp_norm"]):
        assert abs(a - b) <= 1e-10 * b
