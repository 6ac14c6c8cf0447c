"""Host logic (no GPU): the DDPK named-array container that carries Eigen's compressed arrays between a
reference-built set-up driver (oracle/ref_drivers/ddpk_io.h writes it) and the Python mirror."""
import gzip
import os

import numpy as np
import scipy.sparse as sp

from ddpca_b200 import ddpk

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_round_trip_keeps_types_shapes_and_patterns(tmp_path):
    rng = np.random.default_rng(3)
    m = sp.random(40, 25, density=0.2, format="csr", random_state=7)
    m.sort_indices()
    d = {"scal": np.array([3], dtype=np.int64), "vec": rng.standard_normal(17), "idx": np.arange(9, dtype=np.int32), "empty": np.zeros(0)}
    ddpk.put_csr(d, "oper", ddpk.Csr.from_scipy(m))
    p = str(tmp_path / "x.ddpk")
    ddpk.save(p, d)
    e = ddpk.load(p)
    for k in ("scal", "vec", "idx", "empty"):
        assert e[k].dtype == d[k].dtype and np.array_equal(e[k], d[k])
    a = ddpk.get_csr(e, "oper")
    assert tuple(a.shape) == (40, 25) and a.nnz == m.nnz
    assert np.array_equal(a.rowptr, m.indptr) and np.array_equal(a.colidx, m.indices) and np.array_equal(a.val, m.data)
    assert a.rowptr.dtype == np.int32 and a.colidx.dtype == np.int32 and a.val.dtype == np.float64


def test_reads_gzip_fixture_written_by_the_reference_driver():
    """The fixtures under tests/golden were written by the C++ writer inside the reference-built drivers."""
    d = ddpk.load(os.path.join(GOLDEN, "beam_2lev.ddpk.gz"))
    A, P = ddpk.get_hierarchy(d)
    assert len(A) == int(d["maxiLeve"][0]) + 1 and len(P) == len(A) - 1
    for l, p in enumerate(P):
        assert p.shape[0] == A[l + 1].shape[0] and p.shape[1] == A[l].shape[0]
    for a in A:
        m = a.to_scipy()
        assert abs(m - m.T).max() < 1e-9 * abs(m).max()   # stiffness levels are symmetric
