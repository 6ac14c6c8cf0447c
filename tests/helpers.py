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


def _run_ref(cmd, cwd):
    """Run a reference driver once.  A non-zero exit is a failure of the test that asked for it: the
    drivers end through main() or, for `--ref-iters K`, through a deterministic stop on the main thread
    (oracle/ref_drivers/admm_hook.h), so there is nothing to retry."""
    try:
        return subprocess.check_output(cmd, cwd=cwd, stderr=subprocess.PIPE).decode()
    except subprocess.CalledProcessError as e:
        raise AssertionError(f"reference driver failed (exit {e.returncode}): {' '.join(cmd)}\n{(e.stderr or b'').decode()[-2000:]}") from e


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


def run_ref_driver(name, args, tag=None):
    """Run the prebuilt reference driver oracle/_ref/<name> with `args` (+ --out) and load its dump.
    Returns (arrays, json line).  Cached per process."""
    key = (name, tuple(args))
    if key in _cache:
        return _cache[key]
    tmp = tempfile.mkdtemp(prefix="ddpca_ref_")
    out = os.path.join(tmp, (tag or name) + ".ddpk")
    txt = _run_ref([os.path.join(REF_BIN, name)] + [str(a) for a in args] + ["--out", out], tmp)
    meta = json.loads(txt.strip().splitlines()[-1])
    d = ddpk.load(out, copy=False)   # memory map (dumps of several GB); the mapping outlives the unlinked file
    os.remove(out)
    _cache[key] = (d, meta)
    return _cache[key]


def run_ref_torsion(homo=2, musc=2, doma=(1, 8, 4), ref_iters=0):
    """examples/TORSION.h with domain decomposition (tied interfaces, interface-eliminated coarse problem by default)."""
    return run_ref_driver("torsion_admm", ["--homo", homo, "--musc", musc, "--doma", ",".join(str(v) for v in doma), "--ref-iters", ref_iters])


def run_ref_cylinder(copy=1, loca=4, musc=1, ref_iters=0):
    """examples/CYLINDER.h: Hertzian contact of stacked cylinders, local refinement (hanging nodes), frictionless contact."""
    return run_ref_driver("cylinder_admm", ["--copy", copy, "--loca", loca, "--musc", musc, "--ref-iters", ref_iters])


def run_ref_dehw(ref_iters, dd=0, homo=1, loca=1, selo=0, musc=1, nomat=False):
    """examples/DEHW.h: worm drive, FRICTIONAL contact (mu = 0.08 driving worm / 0.2 self-locking), nodal rotations."""
    return run_ref_driver("dehw_admm", ["--dd", dd, "--homo", homo, "--loca", loca, "--selo", selo, "--musc", musc, "--ref-iters", ref_iters] + (["--nomat"] if nomat else []))


def torsion_tangential_displacement(d, disp):
    """Tangential displacement of the nodes on the outer radius of the loaded end of the TORSION shaft
    (examples/TORSION.h:39-49: analytic value T l / (G I_p) R = 1.159111630361142e-06)."""
    vals = []
    for v in range(int(d["nbody"][0])):
        if disp[v] is None:
            continue
        xyz = d[f"body{v}.nodeCoor"].reshape(-1, 3)
        u = np.asarray(disp[v]).reshape(-1, 3)
        r = np.hypot(xyz[:, 0], xyz[:, 1])
        m = (np.abs(r - 0.025) < 1e-9) & (np.abs(xyz[:, 2] - 0.1) < 1e-9)
        if m.any():
            th = np.arctan2(xyz[m, 1], xyz[m, 0])
            vals += list(-u[m, 0] * np.sin(th) + u[m, 1] * np.cos(th))
    return np.array(vals)


def dehw_friction_cases():
    """Replays the trace of the reference's DEHW interface block from tests/golden/dehw_friction.ddpk.gz (built by
    tests/golden/make_dehw_friction_fixture.py from a 260-iteration run of the untouched reference).  Yields, per
    frictional interface, (fricCoef, t, gapTerm, resuCont rows) for the sampled integration points, where
    t = inpoLagr0 l0 - inpoLagr1 l1 + pemaInpo_r0 u0 - pemaInpo_r1 u1 (MCONTACT.h:2632-2635) is evaluated with the
    multipliers of BEFORE the last update, l_{K-1} = l_K - M^-1 (S_p^T u_K - M_p aux_K) (MCONTACT.h:2691-2697)."""
    import scipy.sparse.linalg as spla

    d = ddpk.load(os.path.join(GOLDEN, "dehw_friction.ddpk.gz"))
    for k in range(int(d["niface"][0])):
        q = f"f{k}."
        t = 0.0
        for tv in range(2):
            s = q + f"s{tv}."
            M = ddpk.get_csr(d, s + "inteMass").to_scipy().tocsc()
            Mp = ddpk.get_csr(d, s + "inteMass_pena").to_scipy()
            Sp = ddpk.get_csr(d, s + "systTran_pena").to_scipy()
            Pr = ddpk.get_csr(d, s + "pemaInpo_r").to_scipy()
            IL = ddpk.get_csr(d, s + "inpoLagr").to_scipy()
            u, lam, aux = d[s + "resuDisp_used"], d[s + "inteLagr"], d[s + "inteAuxi"]
            lam_prev = lam - spla.spsolve(M, Sp.T @ u - Mp @ aux)
            t = t + (1.0 if tv == 0 else -1.0) * (IL @ lam_prev + Pr @ u)
        yield float(d[q + "fricCoef"][0]), t, d[q + "gapTerm"], d[q + "resuCont"].reshape(-1, 5)


def check_projection_against_resucont(g, st, cont, mu):
    """gamma / fricStat of the sampled points against the reference's resuCont rows (MCONTACT.h:106-118):
    normal pressure to 1e-8, Coulomb status bit-exact, tangential traction magnitude to 1e-3 -- OUTPUT_PRTR writes
    the traction as gamma_1 b_1 + gamma_2 b_2 in the integration point's own tangent basis (MCONTACT.h:110-113),
    which on the curved tooth flanks is orthonormal to ~1e-4 only, so |traction| and hypot(gamma_1, gamma_2) differ
    by that much in the reference's own data.  Points that sit on a branch boundary to within the accuracy of the
    replayed trace (|gamma_n| or the distance to the cone below 1e-9 of the largest pressure) are left out of the
    exact comparison and counted."""
    gn, g1, g2 = g[0::3], g[1::3], g[2::3]
    scale = max(np.abs(cont[:, 0]).max(), 1e-300)
    tang = np.hypot(g1, g2)
    tang_ref = np.linalg.norm(cont[:, 1:4], axis=1)
    assert np.linalg.norm(gn - cont[:, 0]) <= 1e-8 * np.linalg.norm(cont[:, 0])
    assert np.linalg.norm(tang - tang_ref) <= 1e-3 * max(np.linalg.norm(tang_ref), 1e-300)
    slide = cont[:, 4] == 1
    assert np.allclose(tang[slide], mu * gn[slide], rtol=1e-12, atol=0)      # sliding points sit ON the cone (MCONTACT.h:2654)
    edge = (np.abs(cont[:, 0]) < 1e-9 * scale) | (np.abs(tang_ref - mu * cont[:, 0]) < 1e-9 * scale) & (cont[:, 0] > 0) & (cont[:, 4] == 2)
    ok = st[1::3] == cont[:, 4].astype(np.int32)
    assert ok[~edge].all()
    return int(edge.sum())
