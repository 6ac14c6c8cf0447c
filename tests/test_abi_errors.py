"""Host logic (no GPU): argument validation of the C ABI.  The reference never throws -- every function returns a
status and prints; the ABI mirrors that with `int` 0 / non-zero + ddpca_last_error().  Bad input must
be refused with a message BEFORE any device is touched (so these run on a box without a GPU) and must never crash:
empty and one-row operators, ragged or corrupt CSR arrays, null pointers, out-of-range counts."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

import ddpca_b200 as dd
from ddpca_b200 import ddpk
from ddpca_b200.lib import load_library


def _err(lib):
    return lib.ddpca_last_error().decode()


def test_plans_of_empty_and_single_row_operators():
    empty = ddpk.Csr((0, 0), np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0))
    pl = dd.Plan(empty, dd.SMOOTH_MC)
    assert (pl.n, pl.ngroups, pl.nstages) == (0, 0, 0)
    one = ddpk.Csr.from_scipy(sp.csr_matrix(np.array([[2.0]])))
    for mode in (dd.SMOOTH_MC, dd.SMOOTH_LEX):
        pl = dd.Plan(one, mode)
        assert (pl.n, pl.ngroups, pl.nstages) == (1, 1, 1) and pl.perm.tolist() == [0]
    # the host half of a hierarchy set-up accepts a one-level 1 x 1 "hierarchy"
    assert dd.setup_dryrun([([one], [])])["device_bytes"] > 0


@pytest.mark.parametrize("rowptr,colidx,what", [
    ([0, 2, 3], [0, 5, 1], "out of range"),            # column beyond n
    ([0, 2, 4], [1, 0, 1, 0], "not strictly increasing"),   # unsorted row (Eigen's compressed rows are sorted)
    ([0, 2, 4], [0, 0, 0, 1], "not strictly increasing"),   # duplicate entry
    ([0, 3, 2], [0, 1, 1], "not strictly increasing|row pointers"),   # decreasing row pointers
])
def test_plan_refuses_corrupt_csr_arrays(rowptr, colidx, what):
    a = ddpk.Csr((2, 2), np.array(rowptr, np.int32), np.array(colidx, np.int32), np.ones(len(colidx)))
    with pytest.raises(dd.DdpcaError, match=what):
        dd.Plan(a, dd.SMOOTH_MC)


def test_create_calls_refuse_bad_arguments_before_touching_a_device():
    lib = load_library()
    h = C.c_void_p()
    for nlev in (0, 17):   # MGPIS::maxiLeve + 1 levels; the engine supports 1..16
        assert lib.ddpca_mg_create(0, nlev, None, None, None, None, None, None, None, dd.SMOOTH_MC, C.byref(h)) != 0
        assert "ddpca_mg_create: bad argument" in _err(lib)
    assert lib.ddpca_mg_create_batch(0, 0, 2, None, None, None, None, None, None, None, dd.SMOOTH_MC, C.byref(h)) != 0
    assert lib.ddpca_admm_create(0, 0, 0, 0, C.byref(h)) != 0 and "ddpca_admm_create" in _err(lib)
    assert lib.ddpca_ldlt_create(0, 0, None, None, None, None, None, C.byref(h)) != 0 and "ddpca_ldlt_create" in _err(lib)
    assert lib.ddpca_ldlt_create_dense(0, 3, None, None, None, C.byref(h)) != 0 and "ddpca_ldlt_create_dense" in _err(lib)
    assert lib.ddpca_ldlt_create_dense(0, 40000, None, None, None, C.byref(h)) != 0 and "32768" in _err(lib)
    # the Coulomb projection works on 1 (frictionless) or 3 (frictional) components per integration point (MCONTACT.h:2636-2668)
    assert lib.ddpca_gamma_project(0, 4, 2, C.c_double(0.1), None, None, None, None) != 0 and "ddpca_gamma_project" in _err(lib)


def test_unknown_smoother_and_missing_prolongations_are_named():
    from tests.helpers import load_golden

    d, meta, A, P = load_golden("beam_2lev")
    from ddpca_b200.mgpis import hierarchy_pointers

    lib = load_library()
    ptrs = hierarchy_pointers([(A, P)])
    h = C.c_void_p()
    assert lib.ddpca_mg_create(0, len(A), *ptrs, 7, C.byref(h)) != 0 and "smoother" in _err(lib)
    noP = ptrs[:4] + (None, None, None)
    assert lib.ddpca_mg_create(0, len(A), *noP, dd.SMOOTH_MC, C.byref(h)) != 0 and "prolongation" in _err(lib)
    sec = (C.c_double * 9)()
    assert lib.ddpca_mg_setup_dryrun(1, len(A), *noP, dd.SMOOTH_MC, sec, None, None, None) != 0 and "prolongation" in _err(lib)


def test_null_handles_are_refused_and_destroy_is_idempotent_on_null():
    lib = load_library()
    assert lib.ddpca_mg_pcg(None, 1, None, None, C.c_double(1e-14), C.c_long(1), None, None, None) != 0
    assert "ddpca_mg_pcg" in _err(lib)
    assert lib.ddpca_admm_step(None, 0, None, None, None) != 0
    assert lib.ddpca_admm_row_length(None) == -1
    assert lib.ddpca_mg_destroy(None) == 0 and lib.ddpca_admm_destroy(None) == 0
    assert lib.ddpca_ldlt_destroy(None) == 0 and lib.ddpca_plan_destroy(None) == 0


def test_partition_edge_cases():
    """ddpca_partition_bodies (the bin-packing of bodies onto devices): more devices than bodies leaves
    devices empty but places every body; zero bodies or zero devices is an error."""
    lib = load_library()
    w = (C.c_double * 3)(1.0, 2.0, 3.0)
    rank = (C.c_int * 3)()
    assert lib.ddpca_partition_bodies(3, w, 0, None, 0, rank) != 0
    assert lib.ddpca_partition_bodies(0, w, 0, None, 2, rank) != 0
    assert lib.ddpca_partition_bodies(3, w, 0, None, 8, rank) == 0
    assert sorted(rank) == [0, 1, 2]            # one body each, heaviest first
    assert rank[2] == 0
    assert lib.ddpca_partition_bodies(3, w, 0, None, 1, rank) == 0 and list(rank) == [0, 0, 0]
