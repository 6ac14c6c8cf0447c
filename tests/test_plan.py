"""Host logic: the stage plan (groups / stages / permutation) the device layout is built from."""
import numpy as np
import pytest
import scipy.sparse as sp

import ddpca_b200 as dd
from ddpca_b200 import ddpk
from tests.helpers import load_golden


def _pattern_rows(a, i):
    return a.colidx[a.rowptr[i] : a.rowptr[i + 1]]


@pytest.mark.parametrize("mode", [dd.SMOOTH_LEX, dd.SMOOTH_MC])
@pytest.mark.parametrize("name", ["beam_2lev", "beam_3lev", "block_lagrange"])
def test_plan_is_valid(name, mode):
    d, meta, A, P = load_golden(name)
    for a in A:
        n = a.shape[0]
        pl = dd.Plan(a, mode)
        assert sorted(pl.perm.tolist()) == list(range(n))
        gs = pl.group_start
        assert gs[0] == 0 and gs[-1] == n and (np.diff(gs) >= 1).all() and (np.diff(gs) <= 3).all()
        # rows of a group are consecutive reference rows with one column pattern
        for g in range(pl.ngroups):
            rows = pl.perm[gs[g] : gs[g + 1]]
            assert (np.diff(rows) == 1).all()
            for r in rows[1:]:
                assert np.array_equal(_pattern_rows(a, rows[0]), _pattern_rows(a, r))
        # stages: no coupling between different groups of one stage
        m = a.to_scipy()
        mp = m[pl.perm][:, pl.perm].tocsr()
        grp = np.repeat(np.arange(pl.ngroups), np.diff(gs))
        stage_of_row = np.repeat(np.arange(pl.nstages), np.diff(pl.stage_start))
        coo = mp.tocoo()
        same_stage = stage_of_row[coo.row] == stage_of_row[coo.col]
        assert (grp[coo.row][same_stage] == grp[coo.col][same_stage]).all()


@pytest.mark.parametrize("name", ["beam_2lev", "beam_3lev"])
def test_lex_plan_keeps_the_reference_lower_upper_split(name):
    """LEX is a topological re-ordering: every coupled pair keeps its relative order, so
    strictly-lower(P A P^T) == P strictly-lower(A) P^T and the sweeps are the reference's."""
    d, meta, A, P = load_golden(name)
    for a in A:
        pl = dd.Plan(a, dd.SMOOTH_LEX)
        iperm = np.empty_like(pl.perm)
        iperm[pl.perm] = np.arange(a.shape[0], dtype=np.int32)
        coo = a.to_scipy().tocoo()
        lower = coo.row > coo.col
        assert ((iperm[coo.row] > iperm[coo.col]) == lower).all()


def test_mc_plan_uses_few_colours():
    d, meta, A, P = load_golden("beam_3lev")
    pl = dd.Plan(A[-1], dd.SMOOTH_MC)
    assert pl.nstages <= 12  # 8 for a structured 27-point hexahedral stencil


def test_plan_rejects_missing_diagonal():
    m = sp.csr_matrix(np.array([[1.0, 2.0], [3.0, 0.0]]))
    m.eliminate_zeros()
    with pytest.raises(dd.DdpcaError, match="diagonal"):
        dd.Plan(ddpk.Csr.from_scipy(m), dd.SMOOTH_MC)


def test_plan_handles_ragged_groups():
    """Individually constrained DOFs (BLOCK side faces, MULTIGRID.h:1188-1193) break the
    3-row alignment: groups of 1, 2 and 3 rows must all be accepted."""
    rng = np.random.default_rng(0)
    d, meta, A, P = load_golden("beam_2lev")
    m = A[-1].to_scipy()
    keep = np.ones(m.shape[0], dtype=bool)
    keep[rng.choice(m.shape[0], size=m.shape[0] // 7, replace=False)] = False
    sub = m[keep][:, keep]
    a = ddpk.Csr.from_scipy(sub)
    for mode in (dd.SMOOTH_LEX, dd.SMOOTH_MC):
        pl = dd.Plan(a, mode)
        sizes = np.diff(pl.group_start)
        assert set(sizes.tolist()) <= {1, 2, 3} and 1 in sizes and 3 in sizes


@pytest.mark.parametrize("mode", [dd.SMOOTH_MC, dd.SMOOTH_LEX])
def test_block_diagonal_level_planned_by_blocks_equals_whole_level_plan(mode):
    """A batch of subdomains is one block-diagonal level; its blocks are planned in parallel and merged
    (build_level_plan_blocks).  The merged plan must be the plan of the whole level: same permutation, groups, stages."""
    import scipy.sparse as sp

    from tests.helpers import load_golden

    d, meta, A, P = load_golden("beam_3lev")
    a, b = A[1].to_scipy(), A[2].to_scipy()
    cat = ddpk.Csr.from_scipy(sp.block_diag([b, a, b], format="csr"))
    off = np.cumsum([0, b.shape[0], a.shape[0], b.shape[0]])
    whole = dd.Plan(cat, mode)
    blocks = dd.Plan(cat, mode, sub_off=off)
    assert whole.nstages == blocks.nstages and whole.ngroups == blocks.ngroups
    assert np.array_equal(whole.perm, blocks.perm)
    assert np.array_equal(whole.group_start, blocks.group_start)
    assert np.array_equal(whole.stage_start, blocks.stage_start)


def test_triangular_factor_plan_equals_lex_plan_of_both_factors():
    """ddpca_ldlt_create plans the two triangular solves with a direct wavefront computation on L (build_tri_plan);
    it must be the LEX plan of I + L and of I + L^T (same permutation, single-row groups, same stages)."""
    import scipy.sparse as sp

    n = 14
    lap = sp.kronsum(sp.diags([-1, 2.5, -1], [-1, 0, 1], shape=(n, n)), sp.diags([-1, 2.5, -1], [-1, 0, 1], shape=(n, n))).toarray()
    c = np.linalg.cholesky(lap)
    Ls = sp.csr_matrix(np.tril(np.where(np.abs(c) > 1e-12, c, 0.0), -1))
    Ls.sort_indices()
    eye = sp.identity(n * n, format="csr")
    tri = dd.Plan(ddpk.Csr.from_scipy(Ls), tri=True)
    for T in (Ls + eye, Ls.T.tocsr() + eye):
        T = T.tocsr()
        T.sort_indices()
        ref = dd.Plan(ddpk.Csr.from_scipy(T), dd.SMOOTH_LEX)
        assert ref.ngroups == n * n == tri.ngroups and ref.nstages == tri.nstages
        assert np.array_equal(ref.perm, tri.perm) and np.array_equal(ref.stage_start, tri.stage_start)


def test_host_half_of_the_hierarchy_set_up_runs_without_a_device():
    """ddpca_mg_setup_dryrun: plans, permuted operators, kernel layouts, chunk tables and transfer forms of a batch of
    hierarchies are built on the host exactly as ddpca_mg_create_batch does before its uploads.  Multicolour levels above
    level 0 qualify for the TMA-staged layout, the lexicographic wavefronts of this small mesh do not (stages too
    small); the HBM footprint of a batch is the sum over its members."""
    d, meta, A, P = load_golden("beam_3lev")
    one = dd.setup_dryrun([(A, P)])
    three = dd.setup_dryrun([(A, P)] * 3)
    # (a batch has larger colours: its level 1 reaches the grid-wide passes, one hierarchy alone stays on the v1 kernels there)
    assert 1 <= one["v2_levels"] <= three["v2_levels"] == 2
    nnz = sum(a.val.size for a in A[1:])
    assert 8 * nnz < one["device_bytes"] < 40 * nnz
    assert abs(three["device_bytes"] - 3 * one["device_bytes"]) < 0.1 * three["device_bytes"]
    lex = dd.setup_dryrun([(A, P)], dd.SMOOTH_LEX)
    assert lex["v2_levels"] == 0
    assert set(one["seconds"]) >= {"plan", "permute", "layout", "transfer_permute"}
    with pytest.raises(dd.DdpcaError):
        dd.setup_dryrun([(A, P)], smoother=7)


def test_block_wise_set_up_equals_the_concatenated_one(monkeypatch):
    """A batch is planned and permuted straight from its members' arrays (no block-diagonal copy of the operators);
    DDPCA_SETUP_CONCAT=1 selects the explicit concatenation the first implementation used.  Every array that would be
    uploaded must be identical (hash over plans, layouts, chunk tables, transfer operators)."""
    d, meta, A, P = load_golden("beam_3lev")
    d2, meta2, A2, P2 = load_golden("beam_2lev")
    for mode in (dd.SMOOTH_MC, dd.SMOOTH_LEX):
        monkeypatch.delenv("DDPCA_SETUP_CONCAT", raising=False)
        direct = dd.setup_dryrun([(A, P), (A, P), (A, P)], mode)
        monkeypatch.setenv("DDPCA_SETUP_CONCAT", "1")
        concat = dd.setup_dryrun([(A, P), (A, P), (A, P)], mode)
        assert direct["checksum"] == concat["checksum"] and direct["device_bytes"] == concat["device_bytes"]
    # members of different size
    hs = [(A[1:], P[1:]), (A2, P2)] if len(A) - 1 == len(A2) else [(A, P), (A, P)]
    monkeypatch.delenv("DDPCA_SETUP_CONCAT", raising=False)
    direct = dd.setup_dryrun(hs)
    monkeypatch.setenv("DDPCA_SETUP_CONCAT", "1")
    assert dd.setup_dryrun(hs)["checksum"] == direct["checksum"]


def test_host_set_up_arrays_are_pinned_by_checksum():
    """Regression anchor of the host half (plans, permutation, layouts, chunk tables, transfer operators and their
    node-triple forms): hashes recorded with the first, serial implementation of the CSR helpers.  A batch of 8 is
    large enough for the threaded transpose / compaction paths (plan.cpp)."""
    d, meta, A, P = load_golden("beam_3lev")
    assert dd.setup_dryrun([(A, P)] * 3, dd.SMOOTH_MC)["checksum"] == 9101311763389178906
    assert dd.setup_dryrun([(A, P)], dd.SMOOTH_MC)["checksum"] == 5640188810714335141
    assert dd.setup_dryrun([(A, P)] * 3, dd.SMOOTH_LEX)["checksum"] == 2724131200358734289
    assert dd.setup_dryrun([(A, P)] * 8, dd.SMOOTH_MC)["checksum"] == 16331060792140387340
    assert dd.setup_dryrun([(A, P)] * 8, dd.SMOOTH_LEX)["checksum"] == 8606172048292588632


@pytest.mark.parametrize("name", ["beam_3lev", "block_lagrange"])
def test_planning_code_is_clean_under_address_and_ub_sanitizers(tmp_path, name):
    """tests/cpp/plan_sanitize.cpp: every planning / permutation / transpose / compaction routine of csrc/plan.cpp over
    every level of a reference-written hierarchy, compiled with -fsanitize=address,undefined (the BEAM hierarchy and the
    condensed dual-mortar system with its wide ragged rows); the native run re-checks the plan invariants as well."""
    import gzip
    import os
    import shutil
    import subprocess

    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if cxx is None:
        pytest.skip("g++ not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "plan_sanitize")
    r = subprocess.run([cxx, "-O1", "-g", "-std=c++17", "-fopenmp", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                        os.path.join(root, "tests", "cpp", "plan_sanitize.cpp"), os.path.join(root, "ddpca-admm_b200", "csrc", "plan.cpp"), "-o", exe],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0 and b"asan" in r.stdout.lower():
        pytest.skip("sanitizer runtime not installed")
    assert r.returncode == 0, r.stdout.decode()[-2000:]
    raw = str(tmp_path / (name + ".ddpk"))
    with gzip.open(os.path.join(root, "tests", "golden", name + ".ddpk.gz"), "rb") as f, open(raw, "wb") as g:
        shutil.copyfileobj(f, g)
    p = subprocess.run([exe, raw], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600, env=dict(os.environ, UBSAN_OPTIONS="halt_on_error=1"))
    assert p.returncode == 0, p.stderr.decode()[-3000:]
    assert p.stdout.decode().startswith("OK levels 3")
