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
