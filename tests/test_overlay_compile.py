"""Host logic (no GPU): every example program of the reference compiles UNCHANGED against the drop-in
headers -- `class MGPIS` replaced by the overlay (force-included host/MGPIS.h, same include guard) and
every `CONTACT_ANALYSIS()` call redirected to the device loop (host/MCONTACT_B200.h).  A missing member,
a changed signature or a different return type of the overlay shows up here as a compile error.
Needs /root/reference (the build container); skipped on the GPU box."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
HOST = os.path.join(ROOT, "ddpca-admm_b200", "host")
EXAMPLES = ["BEAM", "BLOCK", "CYLINDER", "DEHW", "TORSION"]
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "examples")) or CXX is None, reason="reference sources or g++ not available")
def test_reference_examples_compile_against_the_overlay():
    base = [CXX, "-O0", "-fsyntax-only", "-std=c++17", "-fopenmp", "-include", "MGPIS.h", "-include", "MCONTACT.h",
            "-DDDPCA_HOOK_CONTACT_ANALYSIS", "-include", "MCONTACT_B200.h", "-I" + HOST, "-I" + os.path.join(ROOT, "include"),
            "-I" + REF, "-I" + os.path.join(ROOT, "oracle", "ref_drivers")]
    procs = {e: subprocess.Popen(base + [os.path.join(REF, "examples", e + ".cpp")], cwd=HOST, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for e in EXAMPLES}
    failed = {}
    for e, p in procs.items():
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            failed[e] = out[-2000:]
    assert not failed, failed


def _has_device():
    try:
        import ddpca_b200 as dd

        return dd.device_count() > 0
    except Exception:
        return False


@pytest.mark.skipif(not os.access(os.path.join(HOST, "_bin", "block_lagrange_b200"), os.X_OK), reason="overlay binaries not built")
@pytest.mark.skipif(_has_device(), reason="a device is present: the solve succeeds (tests/test_gpu_overlay.py)")
def test_overlay_solver_without_a_device_ends_the_process_loudly():
    """No CPU fallback behind the class surface either.  The reference's callers ignore return values:
    MCONTACT::LAGRANGE would take the untouched zero vector for a solution and its active-set loop
    (MCONTACT.h:3690-3698) would never end.  The overlay therefore prints the error and terminates with status 3."""
    import tempfile

    tmp = tempfile.mkdtemp(prefix="ddpca_nodev_")
    p = subprocess.run([os.path.join(HOST, "_bin", "block_lagrange_b200"), "--glob", "1", "--divi", "2,2,2"], cwd=tmp,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert p.returncode == 3
    assert b"MGPIS::BiCGSTAB_SOLV (B200): ERROR" in p.stderr and b"no CPU fallback" in p.stderr
    # the ADMM loop's overlay follows the same policy (an example would otherwise write result files of a zero state)
    exe = os.path.join(HOST, "_bin", "block_b200")
    if os.access(exe, os.X_OK):
        p = subprocess.run([exe, "--glob", "2", "--divi", "2,2,2"], cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
        assert p.returncode == 3 and b"MCONTACT::CONTACT_ANALYSIS (B200): ERROR" in p.stderr
        p = subprocess.run([exe, "--glob", "2", "--divi", "2,2,2"], cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300,
                           env=dict(os.environ, DDPCA_CONTINUE_ON_ERROR="1"))
        assert p.returncode == 1 and b'"error":true' in p.stdout      # the reference's convention: printed, -1 returned, the driver reports it


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "examples")) or CXX is None, reason="reference sources or g++ not available")
def test_overlay_establish_fills_the_members_exactly_as_the_reference(tmp_path):
    """SURVEY.md §8 row a2 on the CPU: the reference's MGPIS::ESTABLISH (MGPIS.h:40-53) and the overlay's, compiled into
    one program under different class names (tests/cpp/establish_parity.cpp), give bit-identical consLowe / consDiag /
    consUppe on random hierarchies with ragged rows, stored zeros and rows without a stored diagonal."""
    exe = str(tmp_path / "establish_parity")
    lib = os.path.join(ROOT, "ddpca-admm_b200", "lib")
    subprocess.check_call([CXX, "-O1", "-std=c++17", "-fopenmp", "-I" + REF, "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "establish_parity.cpp"), "-o", exe,
                           "-L" + lib, "-lddpca_b200", "-Wl,-rpath," + lib], cwd=os.path.join(ROOT, "tests", "cpp"))
    out = subprocess.check_output([exe], timeout=120).decode()
    assert out.startswith("OK levels compared: 10")
