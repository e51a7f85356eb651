#!/usr/bin/env python
"""One small pass over every kernel of libcgb200.so -- written to run under compute-sanitizer

    compute-sanitizer --tool memcheck --error-exitcode 1 python profiles/sanitize_small.py

(closed on this GPU pool in round 2, so it ran plain; it still found the looped-back wait on a ragged
shard that tests/test_gpu_parity.py::test_autotune_alone_on_a_ragged_shard now pins).

Sizes are tiny and ragged on purpose (n not a multiple of any tile edge): every mat-vec variant incl. the
tensor-map ones, both schedules of the CG loop, the 8-way shard in loopback (LL exchange on one GPU), the
reference-topology kernels ("compat"), the device-side COO densification, the dot and the DEBUG block.
Each result is still compared with what the un-instrumented run must give (schedules agree bit for bit).
"""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cgb = importlib.import_module("conjugate-gradient_b200")


def solve(ctx, n, iters, schedule):
    ctx.set_option("schedule", schedule)
    x = np.zeros(n)
    info, hist = ctx.solve(x, iters, 1e-10, history=True)
    nx, rr = ctx.residual_check()
    return x, hist, info.k, nx, rr


def main():
    names = cgb.gemv_variants()
    n = 1237
    with cgb.Context(n) as ctx:
        ctx.generate_lap2d()
        b = cgb.init_source_term(n)
        ctx.set_rhs(b)
        v = np.linspace(-1.0, 1.0, n)
        y0 = None
        for i, name in enumerate(names):
            ctx.set_option("gemv_variant", i)
            y, pap = ctx.gemv(v)
            if y0 is None:
                y0 = y
            assert np.array_equal(y, y0), name
        ctx.set_option("gemv_variant", 0)
        ref = solve(ctx, n, 40, 0)
        for v_i in range(len(names)):
            if not names[v_i].startswith("tma_"):
                continue
            ctx.set_option("gemv_variant", v_i)
            got = solve(ctx, n, 40, 1)
            assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]) and got[2:] == ref[2:], names[v_i]
        ctx.set_option("gemv_variant", 0)
        ctx.set_option("schedule", 1)
        ctx.autotune(4)
        assert abs(ctx.dot(v, b) - ctx.dot(b, v)) == 0.0
        for bw, transposed in ((32, 0), (64, 1)):
            ctx.set_option("num_threads", 256)
            ctx.set_option("block_width", bw)
            ctx.set_option("transposed", transposed)
            ctx.set_option("compat", 1)
            solve(ctx, n, 10, 0)
    # 8-way shard, loopback: the LL exchange, the peer-store path and the persistent kernel's waits on one GPU
    n = 2311
    with cgb.Context(n, 0, 8, 0) as ctx:
        ctx.set_option("loopback", 1)
        ctx.generate_lap2d()
        ctx.set_rhs(cgb.init_source_term(n))
        out = []
        for schedule in (0, 1):
            ctx.set_option("schedule", schedule)
            ctx.solve_begin(None, 30, 0.0, True)
            ctx.iterate(30)
            hist = np.zeros(30)
            ctx.solve_end(None, hist)
            out.append(hist)
        assert np.array_equal(out[0], out[1], equal_nan=True)
    # device-side COO densification, duplicates and symmetric expansion
    n = 517
    rng = np.random.default_rng(7)
    nz = 20000
    irn = rng.integers(0, n, nz).astype(np.int32)   # 0-based, as MatrixCOO::read hands them over
    jcn = rng.integers(0, n, nz).astype(np.int32)
    val = rng.standard_normal(nz)
    with cgb.Context(n) as ctx:
        for sym in (False, True):
            ctx.set_matrix_coo(irn, jcn, val, sym)
            A = ctx.get_matrix_rows(0, n)
            assert np.isfinite(A).all()
    print("sanitize_small: all passes done")


if __name__ == "__main__":
    main()
