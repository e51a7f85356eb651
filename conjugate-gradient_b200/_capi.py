"""ctypes binding of libcgb200.so (the C ABI declared in include/cgb200.h).

Plumbing only: every call goes straight to the sm_100a library.  There is no CPU fallback --
if the library is missing or no B200 is present the calls raise CgbError.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcgb200.so")

UNIQUE_ID_BYTES = 128
EXCHANGE_BLOB_BYTES = 128
_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)


class CgbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"cgb200 error {code}: {msg}")
        self.code = code


class Layout(C.Structure):
    _fields_ = [("n", C.c_int64), ("ld", C.c_int64), ("rows", C.c_int64), ("row0", C.c_int64),
                ("rank", C.c_int), ("world", C.c_int), ("device", C.c_int), ("nblk", C.c_int),
                ("sm_count", C.c_int), ("nchunks", C.c_int64)]


class SolveInfo(C.Structure):
    _fields_ = [("k", C.c_int64), ("converged", C.c_int), ("rsold", C.c_double),
                ("rsnew", C.c_double), ("seconds", C.c_double), ("iterations", C.c_int64)]


# every symbol include/cgb200.h declares: (name, restype, argtypes)
_CTX = C.c_void_p
SIGNATURES = [
    ("cgb_abi_version", C.c_int, []),
    ("cgb_last_error", C.c_char_p, []),
    ("cgb_device_count", C.c_int, [C.POINTER(C.c_int)]),
    ("cgb_partition", C.c_int, [C.c_int64, C.c_int, _i64p, _i64p]),
    ("cgb_create", C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(_CTX)]),
    ("cgb_destroy", C.c_int, [_CTX]),
    ("cgb_comm_unique_id", C.c_int, [C.c_void_p]),
    ("cgb_comm_init", C.c_int, [_CTX, C.c_void_p]),
    ("cgb_exchange_export", C.c_int, [_CTX, C.c_void_p]),
    ("cgb_exchange_import", C.c_int, [_CTX, C.c_void_p]),
    ("cgb_generate_lap2d", C.c_int, [_CTX]),
    ("cgb_set_matrix_rows", C.c_int, [_CTX, _dp, C.c_int64, C.c_int64, C.c_int64]),
    ("cgb_set_matrix_coo", C.c_int, [_CTX, C.c_int64, _i32p, _i32p, _dp, C.c_int]),
    ("cgb_get_matrix_rows", C.c_int, [_CTX, _dp, C.c_int64, C.c_int64, C.c_int64]),
    ("cgb_init_source_term", C.c_int, [C.c_int64, C.c_double, _dp]),
    ("cgb_set_rhs", C.c_int, [_CTX, _dp]),
    ("cgb_set_option", C.c_int, [_CTX, C.c_char_p, C.c_int64]),
    ("cgb_get_option", C.c_int, [_CTX, C.c_char_p, _i64p]),
    ("cgb_gemv_variant_count", C.c_int, []),
    ("cgb_gemv_variant_name", C.c_char_p, [C.c_int]),
    ("cgb_autotune", C.c_int, [_CTX, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float)]),
    ("cgb_get_layout", C.c_int, [_CTX, C.POINTER(Layout)]),
    ("cgb_solve", C.c_int, [_CTX, _dp, C.c_int64, C.c_double, _dp, C.POINTER(SolveInfo)]),
    ("cgb_solve_begin", C.c_int, [_CTX, _dp, C.c_int64, C.c_double, C.c_int]),
    ("cgb_iterate", C.c_int, [_CTX, C.c_int64, C.POINTER(C.c_float)]),
    ("cgb_solve_end", C.c_int, [_CTX, _dp, _dp, C.POINTER(SolveInfo)]),
    ("cgb_residual_check", C.c_int, [_CTX, _dp, _dp]),
    ("cgb_gemv", C.c_int, [_CTX, _dp, _dp, _dp, _dp]),
    ("cgb_dot", C.c_int, [_CTX, _dp, _dp, _dp]),
    ("cgb_bench_gemv", C.c_int, [_CTX, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    ("cgb_bench_read", C.c_int, [_CTX, C.c_int, C.POINTER(C.c_float)]),
    ("cgb_last_gemv_timing", C.c_int, [_CTX, C.POINTER(C.c_float), _i64p]),
    ("cgb_launch_count", C.c_int, [_CTX, _i64p]),
    ("cgb_trace_read", C.c_int, [_CTX, C.c_int, C.POINTER(C.c_uint64), C.c_int64, _i64p, _i64p]),
]

_lib = None


def load():
    """Load libcgb200.so (built in-tree by __graft_entry__.build()).  Fails loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CgbError(-1, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; "
                               "g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, res, args in SIGNATURES:
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _check(rc: int):
    if rc != 0:
        raise CgbError(rc, load().cgb_last_error().decode("utf-8", "replace"))


def _p(a):
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], \
        "expected a C-contiguous float64 array"
    return a.ctypes.data_as(_dp)


def device_count() -> int:
    n = C.c_int(0)
    _check(load().cgb_device_count(C.byref(n)))
    return n.value


def partition(n: int, psize: int):
    s = (C.c_int64 * psize)()
    c = (C.c_int64 * psize)()
    _check(load().cgb_partition(n, psize, s, c))
    return list(s), list(c)


def init_source_term(n: int, h: float | None = None, out: np.ndarray | None = None) -> np.ndarray:
    """b of CGSolver::init_source_term (host libm loop inside the library, as in the reference)."""
    b = np.empty(n, dtype=np.float64) if out is None else out
    _check(load().cgb_init_source_term(n, (1.0 / n) if h is None else h, _p(b)))
    return b


def unique_id() -> bytes:
    buf = C.create_string_buffer(UNIQUE_ID_BYTES)
    _check(load().cgb_comm_unique_id(buf))
    return buf.raw


def gemv_variants():
    lib = load()
    return [lib.cgb_gemv_variant_name(i).decode() for i in range(lib.cgb_gemv_variant_count())]


class Context:
    """One rank (= one GPU, one row shard).  Thin object wrapper over cgb_ctx*."""

    def __init__(self, n: int, rank: int = 0, world: int = 1, device: int = 0):
        self._lib = load()
        self._h = _CTX()
        _check(self._lib.cgb_create(n, rank, world, device, C.byref(self._h)))
        self.n = n

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cgb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- wiring
    def comm_init(self, uid: bytes):
        buf = C.create_string_buffer(uid, UNIQUE_ID_BYTES)
        _check(self._lib.cgb_comm_init(self._h, buf))

    def exchange_export(self) -> bytes:
        buf = C.create_string_buffer(EXCHANGE_BLOB_BYTES)
        _check(self._lib.cgb_exchange_export(self._h, buf))
        return buf.raw

    def exchange_import(self, blobs):
        """blobs: one exchange_export() result per rank, in rank order."""
        raw = b"".join(blobs)
        assert len(raw) == EXCHANGE_BLOB_BYTES * len(blobs)
        buf = C.create_string_buffer(raw, len(raw))
        _check(self._lib.cgb_exchange_import(self._h, buf))

    def layout(self) -> Layout:
        lay = Layout()
        _check(self._lib.cgb_get_layout(self._h, C.byref(lay)))
        return lay

    def set_option(self, key: str, value: int):
        _check(self._lib.cgb_set_option(self._h, key.encode(), int(value)))

    def get_option(self, key: str) -> int:
        v = C.c_int64()
        _check(self._lib.cgb_get_option(self._h, key.encode(), C.byref(v)))
        return v.value

    def autotune(self, iters: int = 0) -> dict:
        """One-off tile-shape selection (cgb_autotune); returns the choice and every candidate's time."""
        names = gemv_variants()
        chosen = C.c_int(-1)
        us = (C.c_float * len(names))()
        _check(self._lib.cgb_autotune(self._h, iters, C.byref(chosen), us))
        return {"chosen": names[chosen.value],
                "us_per_iteration": {nm: round(float(t), 2) for nm, t in zip(names, us) if t >= 0}}

    # -- inputs
    def generate_lap2d(self):
        _check(self._lib.cgb_generate_lap2d(self._h))

    def set_matrix_rows(self, rows: np.ndarray, first_row: int = 0):
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        _check(self._lib.cgb_set_matrix_rows(self._h, _p(rows), first_row, rows.shape[0], rows.shape[1]))

    def set_matrix_coo(self, irn, jcn, val, symmetric: bool):
        irn = np.ascontiguousarray(irn, dtype=np.int32)
        jcn = np.ascontiguousarray(jcn, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        _check(self._lib.cgb_set_matrix_coo(self._h, irn.size, irn.ctypes.data_as(_i32p),
                                            jcn.ctypes.data_as(_i32p), _p(val), int(bool(symmetric))))

    def get_matrix_rows(self, first_row: int, nrows: int) -> np.ndarray:
        out = np.empty((nrows, self.n), dtype=np.float64)
        _check(self._lib.cgb_get_matrix_rows(self._h, _p(out), first_row, nrows, self.n))
        return out

    def set_rhs(self, b: np.ndarray):
        b = np.ascontiguousarray(b, dtype=np.float64)
        assert b.size == self.n
        _check(self._lib.cgb_set_rhs(self._h, _p(b)))

    # -- solve
    def solve(self, x: np.ndarray, max_iter: int, tol: float = 1e-10, history: bool = False):
        assert x.dtype == np.float64 and x.size == self.n and x.flags["C_CONTIGUOUS"]
        hist = np.zeros(max(max_iter, 1), dtype=np.float64) if history else None
        info = SolveInfo()
        _check(self._lib.cgb_solve(self._h, _p(x), max_iter, tol, _p(hist), C.byref(info)))
        return info, (hist[:info.iterations].copy() if history else None)

    def solve_begin(self, x0, max_iter: int, tol: float = 1e-10, history: bool = False):
        _check(self._lib.cgb_solve_begin(self._h, _p(x0), max_iter, tol, int(history)))

    def iterate(self, iters: int) -> float:
        ms = C.c_float()
        _check(self._lib.cgb_iterate(self._h, iters, C.byref(ms)))
        return ms.value

    def solve_end(self, x=None, hist=None) -> SolveInfo:
        info = SolveInfo()
        _check(self._lib.cgb_solve_end(self._h, _p(x), _p(hist), C.byref(info)))
        return info

    def residual_check(self):
        nx, rr = C.c_double(), C.c_double()
        _check(self._lib.cgb_residual_check(self._h, C.byref(nx), C.byref(rr)))
        return nx.value, rr.value

    # -- kernel-level hooks
    def gemv(self, v: np.ndarray, want_partials: bool = False):
        lay = self.layout()
        y = np.empty(lay.rows, dtype=np.float64)
        bp = np.empty(lay.nchunks, dtype=np.float64) if want_partials else None   # chunk partials of v.(A v)
        pap = C.c_double()
        v = np.ascontiguousarray(v, dtype=np.float64)
        _check(self._lib.cgb_gemv(self._h, _p(v), _p(y), _p(bp), C.byref(pap)))
        return (y, bp, pap.value) if want_partials else (y, pap.value)

    def dot(self, a: np.ndarray, b: np.ndarray) -> float:
        out = C.c_double()
        _check(self._lib.cgb_dot(self._h, _p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)),
                                 C.byref(out)))
        return out.value

    def bench_gemv(self, variant: int = -1, reps: int = 10) -> float:
        ms = C.c_float()
        _check(self._lib.cgb_bench_gemv(self._h, variant, reps, C.byref(ms)))
        return ms.value

    def bench_read(self, reps: int = 10) -> float:
        ms = C.c_float()
        _check(self._lib.cgb_bench_read(self._h, reps, C.byref(ms)))
        return ms.value

    def last_gemv_timing(self):
        ms, n = C.c_float(), C.c_int64()
        _check(self._lib.cgb_last_gemv_timing(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def trace_read(self, which: int):
        """(records[L, blocks, 8] uint64 in launch order, launches seen) of the diagnostic
        timeline of kernel `which` (0 mat-vec, 1 update_xr, 2 update_p); option "trace" = L."""
        cap = self.get_option("trace")
        lay = self.layout()
        nb = lay.nblk if which == 0 else lay.nchunks
        buf = np.zeros(cap * nb * 8, dtype=np.uint64)
        seen, blocks = C.c_int64(), C.c_int64()
        _check(self._lib.cgb_trace_read(self._h, which, buf.ctypes.data_as(C.POINTER(C.c_uint64)),
                                        buf.size, C.byref(seen), C.byref(blocks)))
        rec = buf.reshape(cap, blocks.value, 8)
        n = seen.value
        if n >= cap:   # ring: oldest kept launch first
            rec = np.roll(rec, -(n % cap), axis=0)
        else:
            rec = rec[:n]
        return rec, n

    def launch_count(self) -> int:
        n = C.c_int64()
        _check(self._lib.cgb_launch_count(self._h, C.byref(n)))
        return n.value
