"""ctypes front-end of the CPU oracle (oracle/cg_oracle.c) and the reader restatement.

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by the product package.

Reference citations are in cg_oracle.h; `read_mtx_dense` restates the reader
(/root/reference/code/MPI/matrix_coo.cc:7-60, matrix.cc:6-22, mmio.c:96-217).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)


class _Info(C.Structure):
    _fields_ = [("k", C.c_int64), ("converged", C.c_int), ("rsold", C.c_double),
                ("rsnew", C.c_double), ("norm_x", C.c_double), ("rel_resid", C.c_double)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libcg_oracle.so")
    src = os.path.join(_HERE, "cg_oracle.c")
    hdr = os.path.join(_HERE, "cg_oracle.h")
    stale = (not os.path.exists(so)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(so) for s in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.cgo_generate_lap2d.argtypes = [C.c_int64, _dp]
        L.cgo_generate_lap2d_rows.argtypes = [C.c_int64, C.c_int64, C.c_int64, _dp, C.c_int64]
        L.cgo_init_source_term.argtypes = [C.c_int64, C.c_double, _dp]
        L.cgo_partition.argtypes = [C.c_int64, C.c_int, _i64p, _i64p]
        L.cgo_block_range.argtypes = [C.c_int64, C.c_int, C.c_int, _i64p, _i64p]
        L.cgo_row_dot.argtypes = [_dp, _dp, C.c_int64]
        L.cgo_row_dot.restype = C.c_double
        L.cgo_gemv.argtypes = [C.c_int64, C.c_int64, _dp, C.c_int64, _dp, _dp]
        L.cgo_set_gemv_chunk.argtypes = [C.c_int]
        L.cgo_det_sum.argtypes = [_dp, C.c_int64]
        L.cgo_det_sum.restype = C.c_double
        L.cgo_dot.argtypes = [_dp, _dp, C.c_int64]
        L.cgo_dot.restype = C.c_double
        L.cgo_solve.argtypes = [C.c_int64, _dp, C.c_int64, _dp, _dp, C.c_int64, C.c_double,
                                C.c_int, C.c_int, _dp, C.POINTER(_Info)]
        L.cgo_residual_check.argtypes = [C.c_int64, _dp, C.c_int64, _dp, _dp, _dp, _dp]
        _LIB = L
    return _LIB


def _p(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def generate_lap2d(n: int) -> np.ndarray:
    A = np.empty((n, n), dtype=np.float64)
    lib().cgo_generate_lap2d(n, _p(A))
    return A


def generate_lap2d_rows(n: int, row0: int, nrows: int) -> np.ndarray:
    A = np.empty((nrows, n), dtype=np.float64)
    lib().cgo_generate_lap2d_rows(n, row0, nrows, _p(A), n)
    return A


def init_source_term(n: int, h: float | None = None) -> np.ndarray:
    b = np.empty(n, dtype=np.float64)
    lib().cgo_init_source_term(n, (1.0 / n) if h is None else h, _p(b))
    return b


def partition(n: int, psize: int):
    s = (C.c_int64 * psize)()
    c = (C.c_int64 * psize)()
    lib().cgo_partition(n, psize, s, c)
    return list(s), list(c)


def chunk_partials(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Level 1 of the two-level dot: the chunk256 partials of a_i * b_i over 256-element chunks."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    n = a.size
    out = np.empty((n + 255) // 256, dtype=np.float64)
    for c in range(out.size):       # a chunk alone IS a one-chunk dot: det_sum of one value = the value
        out[c] = dot(a[256 * c:256 * (c + 1)], b[256 * c:256 * (c + 1)])
    return out


def block_range(rows: int, nblk: int, c: int):
    r0, r1 = C.c_int64(), C.c_int64()
    lib().cgo_block_range(rows, nblk, c, C.byref(r0), C.byref(r1))
    return r0.value, r1.value


def gemv(A: np.ndarray, p: np.ndarray) -> np.ndarray:
    rows, n = A.shape
    y = np.empty(rows, dtype=np.float64)
    lib().cgo_gemv(rows, n, _p(A), A.strides[0] // 8, _p(p), _p(y))
    return y


class gemv_chunk:
    """with gemv_chunk(bw): ... -- the reference-topology mat-vec order (csrc/compat.cu)."""

    def __init__(self, block_width: int):
        self.bw = block_width

    def __enter__(self):
        lib().cgo_set_gemv_chunk(self.bw)

    def __exit__(self, *exc):
        lib().cgo_set_gemv_chunk(0)


def det_sum(v: np.ndarray) -> float:
    v = np.ascontiguousarray(v, dtype=np.float64)
    return lib().cgo_det_sum(_p(v), v.size)


def dot(a: np.ndarray, b: np.ndarray) -> float:
    return lib().cgo_dot(_p(a), _p(b), a.size)


@dataclass
class SolveResult:
    x: np.ndarray
    k: int
    converged: bool
    rsold: float
    rsnew: float
    norm_x: float
    rel_resid: float
    hist: np.ndarray  # rsnew of every executed loop index


def solve(A: np.ndarray, b: np.ndarray, x0: np.ndarray | None = None, max_iter: int | None = None,
          tol: float = 1e-10, nranks: int = 1, nblk: int = 148) -> SolveResult:
    n = A.shape[0]
    max_iter = n if max_iter is None else max_iter
    x = np.zeros(n, dtype=np.float64) if x0 is None else np.array(x0, dtype=np.float64)
    hist = np.zeros(max(max_iter, 1), dtype=np.float64)
    info = _Info()
    lib().cgo_solve(n, _p(A), A.strides[0] // 8, _p(b), _p(x), max_iter, tol, nranks, nblk,
                    _p(hist), C.byref(info))
    executed = min(info.k + (1 if info.converged else 0), max_iter)
    return SolveResult(x, info.k, bool(info.converged), info.rsold, info.rsnew, info.norm_x,
                       info.rel_resid, hist[:executed].copy())


def residual_check(A: np.ndarray, b: np.ndarray, x: np.ndarray):
    nx, rr = C.c_double(), C.c_double()
    lib().cgo_residual_check(A.shape[0], _p(A), A.strides[0] // 8, _p(b), _p(x),
                             C.byref(nx), C.byref(rr))
    return nx.value, rr.value


def debug_line(k: int, rsold: float, norm_x: float, rel_resid: float) -> str:
    """The stdout line of cg.cc:152-153 / cg.cu:293-295 (std::scientific, precision 6)."""
    return "\t[STEP %d] residual = %e, ||x|| = %e, ||Ax - b||/||b|| = %e" % (
        k, np.sqrt(rsold), norm_x, rel_resid)


# ----------------------------------------------------------------------------- reader
class MtxError(Exception):
    """The reference prints a message and exit(1)s (matrix_coo.cc:12-33)."""


def read_mtx_dense(path: str) -> np.ndarray:
    """Restates Matrix::read (matrix.cc:6-22) over MatrixCOO::read (matrix_coo.cc:7-60):
    banner check (mmio.c:96-179), `matrix coordinate` required, size line after `%` comments
    (mmio.c:189-217), nz triples `%d %d %lg`, 1-based -> 0-based, later entries overwrite,
    symmetric banner mirrors every entry."""
    try:
        f = open(path, "r")
    except OSError:
        raise MtxError("Could not open matrix")
    with f:
        first = f.readline()
        tok = first.split()
        if len(tok) < 5:
            raise MtxError("Could not process Matrix Market banner.")
        banner, mtx, crd, field, sym = tok[0], tok[1].lower(), tok[2].lower(), tok[3].lower(), tok[4].lower()
        if not banner.startswith("%%MatrixMarket") or mtx != "matrix":
            raise MtxError("Could not process Matrix Market banner.")
        if crd not in ("coordinate", "array") or field not in ("real", "complex", "pattern", "integer") \
                or sym not in ("general", "symmetric", "hermitian", "skew-symmetric"):
            raise MtxError("Could not process Matrix Market banner.")
        if crd != "coordinate":
            raise MtxError("Sorry, this application does not support Market Market type: "
                           "[matrix %s %s %s]" % (crd, field, sym))
        line = f.readline()
        while line and line.startswith("%"):
            line = f.readline()
        rest = f.read().split()
        head = line.split()
        if len(head) >= 3:
            m, n, nz = int(head[0]), int(head[1]), int(head[2])
        else:  # blank line after the comments: the next three integers (mmio.c:207-214)
            m, n, nz = int(rest[0]), int(rest[1]), int(rest[2])
            rest = rest[3:]
    A = np.zeros((m, n), dtype=np.float64)
    is_sym = sym == "symmetric"
    for z in range(nz):
        i, j, a = int(rest[3 * z]) - 1, int(rest[3 * z + 1]) - 1, float(rest[3 * z + 2])
        A[i, j] = a
        if is_sym:
            A[j, i] = a
    return A


def write_lap2d_5pt_mtx(path: str, g: int = 100) -> None:
    """Regenerates the reference's only input fixture, lap2D_5pt_n100.mtx (a true 5-point
    Laplacian on a g x g grid, lower triangle, `coordinate real symmetric`), entry for entry:
    column-major over the lower triangle, rows ascending inside a column."""
    n = g * g
    entries = []
    for j in range(n):
        entries.append((j, j, 4))
        if (j + 1) % g != 0:
            entries.append((j + 1, j, -1))
        if j + g < n:
            entries.append((j + g, j, -1))
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real symmetric\n")
        f.write("% Generated 20-Nov-2014\n")
        f.write("%d %d %d\n" % (n, n, len(entries)))
        for i, j, v in entries:
            f.write("%d %d %2d\n" % (i + 1, j + 1, v))
