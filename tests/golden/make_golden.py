#!/usr/bin/env python
"""Generates tests/golden/*.npz by RUNNING THE REFERENCE ITSELF in this container.

The reference ships no tests or golden vectors, so the pins are
outputs of its own unmodified sources (/root/reference/code/MPI/*.cc, compiled by
oracle/Makefile into oracle/_ref/ against the stand-in mpi.h/cblas.h) with a real
OpenBLAS 0.3.15 behind cblas_* ("openblas") and with plain left-to-right loops ("naive").
Run from the repo root:   python tests/golden/make_golden.py [--full | --ranks | --weak]
(--full adds BASELINE.json's full-size configs: N=20000 to convergence and N=40000 x 200
iterations, OpenBLAS provider only; ~2 minutes and 13 GB of host memory.)
(--weak adds BASELINE.json configs[3], the weak-scaling ladder N = 20000 sqrt(G) x 200 iterations,
for the sizes the reference can run: N = 20000 (G=1) and N = 28284 (G=2, 6.4 GB); G=4 is the
N=40000 fixture of --full; G=8, N = 56568, overflows the reference's `int` index i*m_n+j
(matrix.hh:17) -- no reference run exists for it.)
(--ranks adds multi-rank runs of the reference: P = 2, 4, 8 ranks forked on this host by
oracle/ref_shim/mpi_fork.cc -- the reference's own iteration count depends on P.)
Needs /root/reference (to build oracle/_ref); the produced fixtures are committed so the
tests never need it.

Each fixture holds: n, max_iter, k (the integer printed in "[STEP k]"), the r'r history
(one value per executed loop index), final x, and the numbers of the DEBUG line
(cg.cc:152-153).
"""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.dirname(os.path.abspath(__file__))
LINE = re.compile(r"\[STEP (\d+)\] residual = (\S+), \|\|x\|\| = (\S+), \|\|Ax - b\|\|/\|\|b\|\| = (\S+)")


def run_ref(args, blas, tail=(), threads=8):
    """args + [results file] + tail: `cgsolver N outfile [max_iter]` puts the outfile second."""
    with tempfile.TemporaryDirectory() as td:
        env = dict(os.environ, CGREF_BLAS=blas, CGREF_HIST=os.path.join(td, "hist"),
                   CGREF_XOUT=os.path.join(td, "x"), OPENBLAS_NUM_THREADS=str(threads),
                   OMP_NUM_THREADS=str(threads))
        res = subprocess.run(list(args) + [os.path.join(td, "results.txt")] + list(tail), env=env,
                             check=True, capture_output=True, text=True)
        m = LINE.search(res.stdout)
        assert m, res.stdout
        hist = np.fromfile(os.path.join(td, "hist"), dtype=np.float64)
        x = np.fromfile(os.path.join(td, "x"), dtype=np.float64)
        row = open(os.path.join(td, "results.txt")).read().strip()
    # ddot(v, v) log = [rsnew of every executed iteration ..., DEBUG r.r, b.b, x.x]
    # (rsold at cg.cc:91 is ddot(r_sub, p_sub) on two different buffers: not logged)
    return dict(k=int(m.group(1)), resid_print=float(m.group(2)), norm_x=float(m.group(3)),
                rel_resid=float(m.group(4)), hist=hist[:-3], x=x, stdout_line=m.group(0),
                results_row=row)


def run_ref_ranks(n, ranks, max_iter=None, threads=8):
    """The reference at `ranks` MPI ranks (fork shim); history = the all-reduced r'r values."""
    exe = os.path.join(REF, "cgsolver_ref_mp")
    with tempfile.TemporaryDirectory() as td:
        per = max(1, threads // ranks)
        env = dict(os.environ, CGREF_BLAS="openblas", CGREF_NP=str(ranks),
                   CGREF_ALLREDUCE=os.path.join(td, "ar"), CGREF_XOUT=os.path.join(td, "x"),
                   OPENBLAS_NUM_THREADS=str(per), OMP_NUM_THREADS=str(per))
        args = [exe, str(n), os.path.join(td, "results.txt")] + ([] if max_iter is None else [str(max_iter)])
        res = subprocess.run(args, env=env, check=True, capture_output=True, text=True)
        m = LINE.search(res.stdout)
        assert m, res.stdout
        ar = np.fromfile(os.path.join(td, "ar"), dtype=np.float64)
        x = np.fromfile(os.path.join(td, "x"), dtype=np.float64)
        row = open(os.path.join(td, "results.txt")).read().strip()
    # ar = [r.p, (p'Ap, r'r) per executed iteration]
    return dict(k=int(m.group(1)), resid_print=float(m.group(2)), norm_x=float(m.group(3)),
                rel_resid=float(m.group(4)), hist=ar[2::2].copy(), x=x, stdout_line=m.group(0),
                results_row=row)


def save(name, meta, runs):
    d = dict(meta)
    for blas, r in runs.items():
        for key, v in r.items():
            d[f"{blas}_{key}"] = v
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    for blas, r in runs.items():
        print(f"{name:28s} {blas:9s} k={r['k']:5d} iters_logged={len(r['hist']):5d} "
              f"||x||={r['norm_x']:.6e} relres={r['rel_resid']:.6e}")


def main():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    gen = os.path.join(REF, "cgsolver_ref")
    mtx = os.path.join(REF, "cgsolver_ref_mtx")
    if "--ranks" in sys.argv:
        for n, ranks in [(4096, 2), (4096, 4), (4096, 8), (2048, 2), (1000, 3)]:
            runs = {"openblas": run_ref_ranks(n, ranks)}
            save(f"ranks_n{n}_p{ranks}", dict(n=n, max_iter=n, kind="generate_lap2d", ranks=ranks), runs)
        return
    if "--weak" in sys.argv:
        for n, max_iter in [(20000, 200), (28284, 200)]:
            runs = {"openblas": run_ref([gen, str(n)], "openblas", [str(max_iter)])}
            save(f"full_n{n}_it{max_iter}", dict(n=n, max_iter=max_iter, kind="generate_lap2d"), runs)
        return
    if "--full" in sys.argv:
        for n, max_iter in [(20000, None), (40000, 200)]:
            tail = [] if max_iter is None else [str(max_iter)]
            runs = {"openblas": run_ref([gen, str(n)], "openblas", tail)}
            name = f"full_n{n}" + ("" if max_iter is None else f"_it{max_iter}")
            save(name, dict(n=n, max_iter=n if max_iter is None else max_iter, kind="generate_lap2d"), runs)
        return
    for n, max_iter in [(1024, None), (2048, None), (4096, None), (1448, 200), (1000, 50)]:
        tail = [] if max_iter is None else [str(max_iter)]
        runs = {blas: run_ref([gen, str(n)], blas, tail) for blas in ("openblas", "naive")}
        name = f"gen_n{n}" + ("" if max_iter is None else f"_it{max_iter}")
        save(name, dict(n=n, max_iter=n if max_iter is None else max_iter, kind="generate_lap2d"), runs)
    with tempfile.TemporaryDirectory() as td:
        for g in (30, 100):
            path = os.path.join(td, f"lap2D_5pt_n{g}.mtx")
            O.write_lap2d_5pt_mtx(path, g)
            runs = {blas: run_ref([mtx, path], blas) for blas in ("openblas", "naive")}
            save(f"mtx_lap2d_5pt_n{g}", dict(n=g * g, max_iter=g * g, kind="mtx", grid=g), runs)


if __name__ == "__main__":
    main()
