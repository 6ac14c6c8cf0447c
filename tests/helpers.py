"""Shared helpers for the parity tests."""
import json
import os
import subprocess
import tempfile

import numpy as np

from ddpca_b200 import ddpk

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_BIN = os.path.join(ROOT, "oracle", "_ref")


def load_golden(name):
    d = ddpk.load(os.path.join(GOLDEN, name + ".ddpk.gz"))
    meta = json.load(open(os.path.join(GOLDEN, name + ".json")))
    A, P = ddpk.get_hierarchy(d)
    return d, meta, A, P


def rel(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0)


def have_ref_binary(name="beam_nodd"):
    return os.access(os.path.join(REF_BIN, name), os.X_OK)


_cache = {}


def _run_ref(cmd, cwd, attempts=4):
    """Run a reference driver.  The ADMM drivers stop the reference's loop from a watcher thread with
    _exit(); on rare occasions that races with the OpenMP runtime's teardown (SIGABRT after the dump is
    complete or before it is written) -- retry instead of failing the parity test on it."""
    last = None
    for _ in range(attempts):
        try:
            return subprocess.check_output(cmd, cwd=cwd).decode()
        except subprocess.CalledProcessError as e:
            last = e
    raise last


def run_ref_beam(glob, divi=None, solve=1):
    """Run the prebuilt reference driver (oracle/_ref/beam_nodd) and load what it dumps.
    The binary was compiled from the untouched reference; it does not need /root/reference."""
    key = (glob, tuple(divi) if divi else None, solve)
    if key in _cache:
        return _cache[key]
    tmp = tempfile.mkdtemp(prefix="ddpca_ref_")
    out = os.path.join(tmp, "beam.ddpk")
    cmd = [os.path.join(REF_BIN, "beam_nodd"), "--glob", str(glob), "--out", out, "--solve", str(solve)]
    if divi:
        cmd += ["--divi", ",".join(str(v) for v in divi)]
    txt = _run_ref(cmd, tmp)
    meta = json.loads(txt.strip().splitlines()[-1])
    d = ddpk.load(out)
    os.remove(out)
    A, P = ddpk.get_hierarchy(d)
    _cache[key] = (d, meta, A, P)
    return _cache[key]


def permute_hierarchy(A, P, perms):
    """Symmetric permutation of every level: A~ = A[p][:,p], P~ = P[p_fine][:,p_coarse]
    with p[new] = old (the device numbering of ddpca_plan_get)."""
    As, Ps = [], []
    for l, a in enumerate(A):
        m = a.to_scipy()
        p = perms[l]
        As.append(ddpk.Csr.from_scipy(m[p][:, p]))
    for l, pm in enumerate(P):
        m = pm.to_scipy()
        Ps.append(ddpk.Csr.from_scipy(m[perms[l + 1]][:, perms[l]]))
    return As, Ps


def dense_ldlt_factor(m):
    """(perm, L, D) of a small SPD Csr through a dense Cholesky: identity ordering,
    unit-lower L (strict part, CSR) and D, in the layout ddpca_ldlt_create expects.
    Test-side stand-in for Eigen::SimplicialLDLT when a fixture carries no factors."""
    a = m.to_scipy().toarray()
    c = np.linalg.cholesky(a)
    dg = np.diag(c)
    Lu = c / dg[None, :]
    D = dg * dg
    import scipy.sparse as sp

    L = sp.csr_matrix(np.tril(Lu, -1))
    return np.arange(a.shape[0], dtype=np.int32), ddpk.Csr.from_scipy(L), D


def run_ref_block(glob, divi=None, musc=1, ref_iters=0):
    """Run the prebuilt reference BLOCK driver (oracle/_ref/block_admm) and load its dump."""
    key = ("block", glob, tuple(divi) if divi else None, musc, ref_iters)
    if key in _cache:
        return _cache[key]
    tmp = tempfile.mkdtemp(prefix="ddpca_ref_")
    out = os.path.join(tmp, "block.ddpk")
    cmd = [os.path.join(REF_BIN, "block_admm"), "--glob", str(glob), "--musc", str(musc), "--out", out, "--ref-iters", str(ref_iters)]
    if divi:
        cmd += ["--divi", ",".join(str(v) for v in divi)]
    txt = _run_ref(cmd, tmp)
    meta = json.loads(txt.strip().splitlines()[-1])
    d = ddpk.load(out)
    os.remove(out)
    _cache[key] = (d, meta)
    return _cache[key]


def run_ref_beam_dd(glob, doma=(8, 1, 1), musc=1, ref_iters=0, keep_file=False):
    """Run the prebuilt reference BEAM-with-domain-decomposition driver (oracle/_ref/beam_admm)."""
    key = ("beam_dd", glob, tuple(doma), musc, ref_iters, keep_file)
    if key in _cache:
        return _cache[key]
    tmp = tempfile.mkdtemp(prefix="ddpca_ref_")
    out = os.path.join(tmp, "beam_dd.ddpk")
    cmd = [os.path.join(REF_BIN, "beam_admm"), "--glob", str(glob), "--musc", str(musc), "--out", out, "--ref-iters", str(ref_iters),
           "--doma", ",".join(str(v) for v in doma)]
    txt = _run_ref(cmd, tmp)
    meta = json.loads(txt.strip().splitlines()[-1])
    d = ddpk.load(out)
    if keep_file:
        meta["path"] = out
    else:
        os.remove(out)
    _cache[key] = (d, meta)
    return _cache[key]
