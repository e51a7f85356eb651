#!/usr/bin/env python
"""Timeline of the CG iteration under its PRODUCTION schedule (CUDA graph + programmatic dependent
launch), from the %globaltimer stamps every CTA writes with the option "trace" (cgb_trace_read).

    python profiles/trace_iter.py --case 40000:1 --case 40000:8 [--set gemv_variant=6] [--npz DIR]

`N:W` = rank 0's shard of a W-way row split of an N x N system on ONE GPU; W > 1 runs in
"loopback" (the rank fills every slot of the fused exchange itself: its own timing and traffic
are real, the numbers it iterates on are not a CG solve).  With --real the script is launched by
torchrun on W GPUs and traces every rank of a real W-rank solve.

Per case one JSON line (all times in microseconds, medians over the traced launches):
  iter            last-CTA exit of mat-vec k+1  -  last-CTA exit of mat-vec k
  matvec_span     last-CTA exit  -  first "dependency met" stamp (p final)    = the mat-vec's share
  gap             first "dependency met" of k+1  -  last-CTA exit of k        = vector kernels + sync
  first_tile      first tile consumed  -  dependency met (per CTA, median / max)
  exit_spread     last exit - {min, median} exit over the CTAs                = the tail
  xr / p          entry -> dependency met -> scalar known -> exit of block 0 of the vector kernels
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cgb = importlib.import_module("conjugate-gradient_b200")


def us(a):
    return float(a) / 1e3


def summarize(g, xr, p, rows, n, skip=4):
    """g: [L, nblk, 8] mat-vec records, xr / p: [L, nchunks, 8]."""
    g = g.astype(np.int64)
    L = g.shape[0]
    entry, prefill, dep, first, lastissue, rowsdone, ext = (g[:, :, i] for i in range(7))
    ex_max = ext.max(axis=1)
    ex_min = ext.min(axis=1)
    ex_med = np.median(ext, axis=1)
    dep_min = dep.min(axis=1)
    dep_max = dep.max(axis=1)
    sl = slice(skip, L - 1)
    it = np.diff(ex_max)[skip:]
    span = (ex_max - dep_min)[sl]
    gap = (dep_min[1:] - ex_max[:-1])[skip:]
    ft = (first - dep)[sl]
    out = {
        "launches": int(L), "ctas": int(g.shape[1]),
        "iter_us": us(np.median(it)), "iter_us_min": us(it.min()), "iter_us_max": us(it.max()),
        "matvec_span_us": us(np.median(span)),
        "gap_us": us(np.median(gap)),
        "dep_skew_us": us(np.median((dep_max - dep_min)[sl])),
        "first_tile_us_med": us(np.median(ft)), "first_tile_us_max": us(np.median(ft.max(axis=1))),
        "exit_spread_max_min_us": us(np.median((ex_max - ex_min)[sl])),
        "exit_spread_max_med_us": us(np.median((ex_max - ex_med)[sl])),
        "entry_to_dep_us_med": us(np.median((dep - entry)[sl])),
        "prefill_us_med": us(np.median((prefill - entry)[sl])),
        "lastissue_to_exit_us_med": us(np.median((ext - lastissue)[sl])),
        "rowsdone_to_exit_us_med": us(np.median((ext - rowsdone)[sl])),
        "stream_gbs_span": 8.0 * rows * n / (np.median(span) * 1e-9) / 1e9,
        "stream_gbs_iter": 8.0 * rows * n / (np.median(it) * 1e-9) / 1e9,
    }
    # per-CTA busy time (dependency met -> exit) vs its rows: who is slow?
    busy = (ext - dep)[sl]
    nrows = (g[0, :, 7] >> 32).astype(np.int64)
    per_row = np.median(busy, axis=0) / np.maximum(nrows, 1)
    out["cta_busy_us_p5_p50_p95_max"] = [us(np.percentile(np.median(busy, axis=0), q)) for q in (5, 50, 95, 100)]
    out["cta_us_per_row_p5_p50_p95"] = [us(np.percentile(per_row, q)) for q in (5, 50, 95)]
    out["rows_min_max"] = [int(nrows.min()), int(nrows.max())]
    for name, v in (("xr", xr), ("p", p)):
        if v is None or v.shape[0] < skip + 2:
            continue
        v = v.astype(np.int64)
        b0 = v[skip:, 0, :]
        out[name + "_block0_wait_us"] = us(np.median(b0[:, 1] - b0[:, 0]))
        out[name + "_block0_scalar_us"] = us(np.median(b0[:, 2] - b0[:, 1])) if name == "xr" else None
        out[name + "_dep_to_lastexit_us"] = us(np.median(v[skip:, :, 3].max(axis=1) - v[skip:, :, 1].min(axis=1)))
    if xr is not None and p is not None and xr.shape[0] == L and p.shape[0] == L:
        xr = xr.astype(np.int64)
        p = p.astype(np.int64)
        # chain: mat-vec last exit -> xr dependency met -> xr last exit -> p dep -> p last exit -> next mat-vec dep
        a = (xr[:, :, 1].min(axis=1) - ex_max)[sl]
        b = (xr[:, :, 3].max(axis=1) - xr[:, :, 1].min(axis=1))[sl]
        c_ = (p[:, :, 1].min(axis=1) - xr[:, :, 3].max(axis=1))[sl]
        d = (p[:, :, 3].max(axis=1) - p[:, :, 1].min(axis=1))[sl]
        e = (dep_min[1:] - p[:-1, :, 3].max(axis=1))[skip:]
        out["chain_us"] = {"matvec_exit->xr_dep": us(np.median(a)), "xr": us(np.median(b)),
                           "xr_exit->p_dep": us(np.median(c_)), "p": us(np.median(d)),
                           "p_exit->matvec_dep": us(np.median(e))}
    return out


def summarize_persistent(g, rows, n, skip=4):
    """g: [L, nblk, 8] records of the persistent kernel (one per iteration and CTA): 0 iteration
    start, 3 first tile consumed, 5 rows stored, 1 alpha known, 2 own r'r partials
    published, 4 beta known, 6 own p chunks published."""
    g = g.astype(np.int64)
    L = g.shape[0]
    t0, alpha, rrpub, first, beta, rowsd, ppub = (g[:, :, i] for i in range(7))
    sl = slice(skip, L - 1)
    it = np.diff(rowsd.max(axis=1))[skip:]
    med = lambda a: us(np.median(a))
    last_rows = rowsd.max(axis=1)
    out = {
        "launches": int(L), "ctas": int(g.shape[1]),
        "iter_us": med(it), "iter_us_min": us(it.min()), "iter_us_max": us(it.max()),
        "stream_gbs_iter": 8.0 * rows * n / (np.median(it) * 1e-9) / 1e9,
        # per-CTA phase durations (medians over CTAs and iterations)
        "M_first_tile_to_rows_done_us": med((rowsd - first)[sl]),
        "start_to_first_tile_us": med((first - t0)[sl]),
        "rows_done_spread_max_min_us": med((rowsd.max(axis=1) - rowsd.min(axis=1))[sl]),
        "rows_done_spread_max_med_us": med((rowsd.max(axis=1) - np.median(rowsd, axis=1))[sl]),
        # the chain after the LAST CTA of the GPU has stored its rows
        "chain_us": {
            "last_rows_done->alpha(med cta)": med((np.median(alpha, axis=1) - last_rows)[sl]),
            "alpha->rr_published(med)": med((rrpub - alpha)[sl]),
            "alpha->rr_published(max cta)": med((rrpub - alpha).max(axis=1)[sl]),
            "last_rr_published->beta(med cta)": med((np.median(beta, axis=1) - rrpub.max(axis=1))[sl]),
            "beta->p_published(med)": med((ppub - beta)[sl]),
            "last_p_published->first_tile_next(med cta)": med((np.median(first[1:], axis=1) - ppub[:-1].max(axis=1))[skip:]),
            "last_rows_done->first_tile_next(med cta)": med((np.median(first[1:], axis=1) - last_rows[:-1])[skip:]),
        },
    }
    busy = (rowsd - first)[sl]
    out["cta_M_us_p5_p50_p95_max"] = [us(np.percentile(np.median(busy, axis=0), q)) for q in (5, 50, 95, 100)]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", action="append", default=[])
    ap.add_argument("--iters", type=int, default=64)
    ap.add_argument("--set", action="append", default=[], help="k=v,k=v option set (repeatable)")
    ap.add_argument("--npz", default=None, help="directory for the raw records")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "trace_iter.jsonl"))
    ap.add_argument("--real", action="store_true", help="under torchrun: trace a real W-rank solve")
    a = ap.parse_args()
    cases = a.case or ["40000:1", "40000:8"]
    sets = [dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in s.split(",") if kv) for s in (a.set or [""])]
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    rank = int(os.environ.get("RANK", "0"))
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if a.real:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    outf = open(a.out, "a") if rank == 0 else None
    for case in cases:
        n, world = (int(t) for t in case.split(":"))
        if a.real:
            assert world == world_env, "--real: the case's W must equal WORLD_SIZE"
            ctx = cgb.Context(n, rank, world, int(os.environ.get("LOCAL_RANK", "0")))
            wiring = importlib.import_module("conjugate-gradient_b200.wiring")
            wiring.wire(ctx, rank, world, dist, cgb.unique_id, nccl=False)
        else:
            ctx = cgb.Context(n, 0, world, 0)
            if world > 1:
                ctx.set_option("loopback", 1)
        ctx.generate_lap2d()
        ctx.set_rhs(cgb.init_source_term(n))
        lay = ctx.layout()
        for opts in sets:
            for k, v in opts.items():
                ctx.set_option(k, v)
            ctx.set_option("trace", 0)
            # untraced timing first (the stamps cost a few hundred ns per CTA)
            best = 1e30
            for _ in range(3):
                ctx.solve_begin(None, 200, 0.0, False)
                ms = ctx.iterate(200)
                info = ctx.solve_end(None, None)
                best = min(best, ms / max(1, info.iterations))
            ctx.set_option("trace", a.iters)
            ctx.solve_begin(None, a.iters, 0.0, False)
            ms = ctx.iterate(a.iters)
            info = ctx.solve_end(None, None)
            g, ng = ctx.trace_read(0)
            xr, _ = ctx.trace_read(1)
            p, _ = ctx.trace_read(2)
            lay = ctx.layout()
            rec = {"case": case, "mode": "real" if a.real else ("loopback" if world > 1 else "single"),
                   "rank": rank, "opts": opts, "variant": cgb.gemv_variants()[ctx.get_option("gemv_variant")],
                   "rows": int(lay.rows), "n": n, "untraced_us_per_iter": best * 1e3,
                   "traced_us_per_iter": ms * 1e3 / max(1, info.iterations)}
            persistent = ctx.get_option("schedule_in_use") == 1
            rec["schedule"] = "persistent" if persistent else "graph"
            rec.update(summarize_persistent(g, int(lay.rows), n) if persistent
                       else summarize(g, xr, p, int(lay.rows), n))
            if a.npz:
                os.makedirs(a.npz, exist_ok=True)
                tag = "%s_%s_r%d_%s" % (case.replace(":", "w"), rec["mode"], rank,
                                        "_".join("%s%d" % kv for kv in sorted(opts.items())) or "default")
                np.savez_compressed(os.path.join(a.npz, "trace_" + tag + ".npz"), gemv=g, xr=xr, p=p)
            if a.real:
                allrec = [None] * world
                dist.all_gather_object(allrec, rec)
                if rank == 0:
                    for r in allrec:
                        print(json.dumps(r)); outf.write(json.dumps(r) + "\n")
            else:
                print(json.dumps(rec)); outf.write(json.dumps(rec) + "\n")
            if outf:
                outf.flush()
            ctx.set_option("trace", 0)
        if a.real:
            dist.barrier()
        ctx.close()
    if a.real:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
