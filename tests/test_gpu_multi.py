"""Multi-GPU parity (needs >= 2 B200s on the box: `gpurun --gpus 2 -- python -m pytest tests -m gpu`).

Row-sharded A, one rank per GPU, the two exchanges -- ncclAllGather and the exchange fused into
the mat-vec kernel (peer stores + flags) -- against the CPU oracle's emulated-rank solve, bit
for bit (partial sums combined in rank order make the P-rank run reproducible).
Ranks are threads of this process here (peer access); the one-process-per-GPU path (CUDA IPC)
is covered by test_torchrun_two_ranks and by bench.py --gpus N.
"""
import json
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(240)]   # a stuck collective must not eat GPU time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu(cgb):
    try:
        return cgb.device_count()
    except cgb.CgbError:
        return 0


def _run_ranks(G, fn):
    """fn(rank) on G threads (ctypes releases the GIL, the collective steps overlap)."""
    out, err = [None] * G, []

    def body(r):
        try:
            out[r] = fn(r)
        except BaseException as e:  # noqa: BLE001 - reported below
            err.append((r, e))

    th = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(G)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    if any(t.is_alive() for t in th):
        # a rank blocked inside a CUDA / NCCL call cannot be interrupted, and closing its context would
        # block too: leave the process at once instead of holding the GPUs until an outer timeout
        stuck = [r for r, t in enumerate(th) if t.is_alive()]
        sys.stderr.write("test_gpu_multi: ranks %s are stuck (errors so far: %r)\n" % (stuck, err))
        sys.stderr.flush()
        os._exit(97)
    if err:
        raise err[0][1]
    return out


def _solve_sharded(cgb, O, n, G, max_iter, exchange, setup=None, variant=None, schedule=None):
    b = O.init_source_term(n)
    ctxs = [cgb.Context(n, r, G, r) for r in range(G)]
    try:
        uid = cgb.unique_id()
        _run_ranks(G, lambda r: ctxs[r].comm_init(uid))
        blobs = [c.exchange_export() for c in ctxs]
        for c in ctxs:
            c.exchange_import(blobs)
            c.set_option("exchange", exchange)
            if variant is not None:
                c.set_option("gemv_variant", variant)
            if schedule is not None:
                c.set_option("schedule", schedule)
        in_use = ctxs[0].get_option("schedule_in_use")
        if schedule is not None and exchange == 1:
            assert in_use == schedule, "the requested schedule is not the one in use"

        def rank_body(r):
            c = ctxs[r]
            if setup is None:
                c.generate_lap2d()
            else:
                setup(c)
            c.set_rhs(b)
            x = np.zeros(n)
            info, hist = c.solve(x, max_iter=max_iter, tol=1e-10, history=True)
            nx, rr = c.residual_check()
            return x, info, hist, nx, rr

        res = _run_ranks(G, rank_body)
        nblk = ctxs[0].layout().nblk
    finally:
        for c in ctxs:
            c.close()
    return res, nblk, b


# (exchange, schedule): ncclAllGather between the kernels of a graph; the exchange fused into the
# mat-vec kernel of the graph; the persistent kernel (whole loop in one launch, csrc/persist.cu)
MODES = {"nccl": (0, 0), "fused-graph": (1, 0), "fused-persistent": (1, 1)}


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("n,max_iter", [(1024, 1024), (1001, 120), (2050, 200)])
def test_two_gpu_solve_bitwise_vs_emulated_ranks(cgb, O, n, max_iter, mode):
    if _ngpu(cgb) < 2:
        pytest.skip("needs 2 GPUs")
    G = 2
    exchange, schedule = MODES[mode]
    res, nblk, b = _solve_sharded(cgb, O, n, G, max_iter, exchange, schedule=schedule)
    ref = O.solve(O.generate_lap2d(n), b, max_iter=max_iter, nranks=G, nblk=nblk)
    for r, (x, info, hist, nx, rr) in enumerate(res):
        assert info.k == ref.k and bool(info.converged) == ref.converged, r
        assert np.array_equal(hist, ref.hist), r
        assert np.array_equal(x, ref.x), r            # every rank ends with the full x
        assert nx == ref.norm_x and rr == ref.rel_resid, r


@pytest.mark.parametrize("mode", list(MODES))
def test_all_gpus_match_the_reference(cgb, O, golden_dir, mode):
    """G = every GPU on the box against the UNMODIFIED reference: its 1-rank run and, when a
    fixture exists (P = 2, 4, 8), its own P = G rank run (forked MPI ranks, ranks_n4096_pG.npz).
    Residual norms 1e-10 before the rounding floor, x 1e-9; the iteration count within +-1 of a
    count the reference itself produces for this system ({358, 359, 385} over its BLAS
    providers and rank counts: the tail is order-dependent)."""
    from parity_util import check_against_reference, reference_k_set
    G = _ngpu(cgb)
    if G < 2:
        pytest.skip("needs >= 2 GPUs")
    g1 = np.load(os.path.join(golden_dir, "gen_n4096.npz"))
    n = int(g1["n"])
    ks = reference_k_set(golden_dir, g1)
    exchange, schedule = MODES[mode]
    res, nblk, b = _solve_sharded(cgb, O, n, G, n, exchange, schedule=schedule)
    fixtures = [g1]
    pg = os.path.join(golden_dir, "ranks_n4096_p%d.npz" % G)
    if os.path.exists(pg):
        fixtures.append(np.load(pg))
    for x, info, hist, nx, rr in res:
        for g in fixtures:
            check_against_reference(info.k, hist, x, g, "openblas", "gen_n4096 G=%d" % G, k_refs=ks)
    ref = O.solve(O.generate_lap2d(n), b, max_iter=n, nranks=G, nblk=nblk)
    assert res[0][1].k == ref.k and np.array_equal(res[0][2], ref.hist)


def test_two_gpu_gemv_hook_and_remainder_shard(cgb, O):
    """cgb_gemv on sharded ranks: each rank returns its own rows; the chunk partials of p'Ap (over
    the GLOBAL vector, after the exchange) and their total are the same on every rank -- and the
    same as on one GPU.  n = 1003 gives the last rank the remainder."""
    if _ngpu(cgb) < 2:
        pytest.skip("needs 2 GPUs")
    n, G = 1003, 2
    rng = np.random.default_rng(5)
    A = rng.standard_normal((n, n))
    p = rng.standard_normal(n)
    y_ref = O.gemv(A, p)
    for exchange in (0, 1):
        ctxs = [cgb.Context(n, r, G, r) for r in range(G)]
        try:
            uid = cgb.unique_id()
            _run_ranks(G, lambda r: ctxs[r].comm_init(uid))
            blobs = [c.exchange_export() for c in ctxs]
            for c in ctxs:
                c.exchange_import(blobs)
                c.set_option("exchange", exchange)

            def body(r):
                ctxs[r].set_matrix_rows(A)
                return ctxs[r].gemv(p, want_partials=True)

            res = _run_ranks(G, body)
            starts, counts = cgb.partition(n, G)
            cp_ref = O.chunk_partials(p, y_ref)
            for r in range(G):
                y, cp, pap = res[r]
                assert np.array_equal(y, y_ref[starts[r]:starts[r] + counts[r]])
                assert np.array_equal(cp, cp_ref)
                assert pap == O.det_sum(cp_ref) == O.dot(p, y_ref)
        finally:
            for c in ctxs:
                c.close()


def test_torchrun_two_ranks(cgb, O, tmp_path):
    """One process per GPU (torchrun, NCCL + CUDA IPC wiring), both exchanges, vs the oracle."""
    if _ngpu(cgb) < 2:
        pytest.skip("needs 2 GPUs")
    out = tmp_path / "mp.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(ROOT, "tests", "mp_solve_worker.py"), "--size", "1536", "--max-iter", "150",
           "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = json.loads(out.read_text())
    n = 1536
    b = O.init_source_term(n)
    ref = O.solve(O.generate_lap2d(n), b, max_iter=150, nranks=2, nblk=got["nblk"])
    for mode in ("nccl", "fused", "persistent"):
        assert got[mode]["k"] == ref.k
        assert np.array_equal(np.array(got[mode]["hist"]), ref.hist), mode
        assert np.array_equal(np.array(got[mode]["x"]), ref.x), mode
        assert got[mode]["ranks_agree"], mode
