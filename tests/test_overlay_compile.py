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
