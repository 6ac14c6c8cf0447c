"""Reader/writer for the DDPK named-array container.

DDPK is the on-disk exchange format between the reference-built setup drivers
(`oracle/ref_drivers/*.cpp`, which run the reference's own host C++ for mesh,
search and operator assembly) and this package.  Sparse operators are stored as
Eigen's compressed RowMajor arrays (int32 row pointers / column indices, FP64
values), i.e. exactly what `Eigen::SparseMatrix<double,RowMajor>::outerIndexPtr /
innerIndexPtr / valuePtr` expose, so sparsity patterns are the
reference's bit for bit.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np

_MAGIC = b"DDPK0001"
_DTYPES = {0: np.dtype("<f8"), 1: np.dtype("<i4"), 2: np.dtype("<i8")}
_CODES = {np.dtype("float64"): 0, np.dtype("int32"): 1, np.dtype("int64"): 2}


@dataclass
class Csr:
    """Compressed-sparse-row operator, Eigen RowMajor layout."""

    shape: tuple
    rowptr: np.ndarray  # int32[rows+1]
    colidx: np.ndarray  # int32[nnz], sorted within a row
    val: np.ndarray  # float64[nnz]

    @property
    def nnz(self) -> int:
        return int(self.colidx.shape[0])

    def to_scipy(self):
        import scipy.sparse as sp

        return sp.csr_matrix((self.val, self.colidx, self.rowptr), shape=self.shape)

    @staticmethod
    def from_scipy(m) -> "Csr":
        m = m.tocsr()
        m.sort_indices()
        return Csr(
            tuple(int(s) for s in m.shape),
            np.ascontiguousarray(m.indptr, dtype=np.int32),
            np.ascontiguousarray(m.indices, dtype=np.int32),
            np.ascontiguousarray(m.data, dtype=np.float64),
        )


def load(path: str, copy: bool = True) -> dict:
    """Return {name: ndarray}; arrays are copies (aligned, writable) or, with copy=False, read-only views of a
    memory map of the file (large dumps: no second copy in memory, pages shared between processes)."""
    if not copy and not str(path).endswith(".gz"):
        blob = np.memmap(path, dtype=np.uint8, mode="r")
        if bytes(blob[:8]) != _MAGIC:
            raise ValueError(f"{path}: not a DDPK file")
        out = {}
        off = 8
        n = blob.shape[0]
        while off < n:
            (nl,) = struct.unpack_from("<I", blob, off)
            off += 4
            name = bytes(blob[off : off + nl]).decode()
            off += nl
            dt, cnt = struct.unpack_from("<IQ", blob, off)
            off += 12
            off += (8 - off % 8) % 8
            dtype = _DTYPES[dt]
            out[name] = np.frombuffer(blob, dtype=dtype, count=cnt, offset=off)
            off += cnt * dtype.itemsize
        return out
    if str(path).endswith(".gz"):
        import gzip

        with gzip.open(path, "rb") as f:
            blob = f.read()
    else:
        with open(path, "rb") as f:
            blob = f.read()
    if blob[:8] != _MAGIC:
        raise ValueError(f"{path}: not a DDPK file")
    out = {}
    off = 8
    n = len(blob)
    while off < n:
        (nl,) = struct.unpack_from("<I", blob, off)
        off += 4
        name = blob[off : off + nl].decode()
        off += nl
        dt, cnt = struct.unpack_from("<IQ", blob, off)
        off += 12
        off += (8 - off % 8) % 8
        dtype = _DTYPES[dt]
        nbytes = cnt * dtype.itemsize
        out[name] = np.frombuffer(blob, dtype=dtype, count=cnt, offset=off).copy()
        off += nbytes
    return out


def save(path: str, arrays: dict) -> None:
    with open(path, "wb") as f:
        f.write(_MAGIC)
        off = 8
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            code = _CODES[a.dtype]
            nb = name.encode()
            hdr = struct.pack("<I", len(nb)) + nb + struct.pack("<IQ", code, a.size)
            f.write(hdr)
            off += len(hdr)
            pad = (8 - off % 8) % 8
            f.write(b"\0" * pad)
            off += pad
            f.write(a.tobytes())
            off += a.nbytes


def get_csr(d: dict, name: str) -> Csr:
    shp = d[name + ".shape"]
    return Csr((int(shp[0]), int(shp[1])), d[name + ".rowptr"], d[name + ".colidx"], d[name + ".val"])


def put_csr(d: dict, name: str, m: Csr) -> None:
    d[name + ".shape"] = np.array(m.shape, dtype=np.int64)
    d[name + ".rowptr"] = np.ascontiguousarray(m.rowptr, dtype=np.int32)
    d[name + ".colidx"] = np.ascontiguousarray(m.colidx, dtype=np.int32)
    d[name + ".val"] = np.ascontiguousarray(m.val, dtype=np.float64)


def get_hierarchy(d: dict):
    """(consStif[0..L], realProl[0..L-1]) as lists of Csr (MGPIS.h:12-15)."""
    L = int(d["maxiLeve"][0])
    A = [get_csr(d, f"consStif{l}") for l in range(L + 1)]
    P = [get_csr(d, f"realProl{l}") for l in range(L)]
    return A, P
