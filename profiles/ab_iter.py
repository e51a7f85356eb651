#!/usr/bin/env python
"""A/B timing of the CG iteration loop on one B200: options (schedule, pdl, graph, gemv_variant, ...) x N.
    python profiles/ab_iter.py --sizes 40000,14142 --set pdl=0,graph_unroll=4 --set pdl=1,graph_unroll=8
    python profiles/ab_iter.py --sizes 40000:8,40000:4 --set schedule=0 --set schedule=1
`N:W` = rank 0's shard of a W-way row split in "loopback" (one GPU plays one rank of W; timing and
traffic of that rank are real, the numbers are not a CG solve -- but they are deterministic, so
the bitwise cross-check between option sets still holds).
Prints one JSON line per (N, option set): ms per iteration (device time of cgb_iterate, best of
--reps) and the implied GB/s of A traffic."""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cgb = importlib.import_module("conjugate-gradient_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="40000,14142")
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--set", action="append", default=[], help="k=v,k=v option set (repeatable)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ab_iter.jsonl"))
    a = ap.parse_args()
    sets = [dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in s.split(",") if kv) for s in (a.set or [""])]
    with open(a.out, "w") as out:
        for case in a.sizes.split(","):
            n, world = (int(t) for t in (case.split(":") + ["1"])[:2])
            with cgb.Context(n, 0, world, 0) as ctx:
                if world > 1:
                    ctx.set_option("loopback", 1)
                rows = ctx.layout().rows
                ctx.generate_lap2d()
                ctx.set_rhs(cgb.init_source_term(n))
                ref_hist = None
                for opts in sets:
                    for k, v in opts.items():
                        ctx.set_option(k, v)
                    best = 1e30
                    for rep in range(a.reps + 1):
                        ctx.solve_begin(None, a.iters, 0.0 if world > 1 else 1e-10, rep == 0)
                        ms = ctx.iterate(a.iters)
                        hist = np.zeros(a.iters) if rep == 0 else None
                        info = ctx.solve_end(None, hist)
                        if rep == 0:      # warm-up run doubles as the bitwise cross-check
                            if ref_hist is None:
                                ref_hist = hist
                            same = bool(np.array_equal(hist, ref_hist, equal_nan=True))
                        else:
                            best = min(best, ms)
                    per = best / info.iterations
                    rec = dict(n=n, world=world, rows=int(rows), opts=opts, ms_per_iter=per, it_per_s=1e3 / per,
                               gbs=8.0 * rows * n / per / 1e6, iterations=int(info.iterations),
                               variant=cgb.gemv_variants()[ctx.get_option("gemv_variant")],
                               schedule_in_use=ctx.get_option("schedule_in_use"),
                               hist_equal_to_first_set=same)
                    print(json.dumps(rec)); out.write(json.dumps(rec) + "\n"); out.flush()


if __name__ == "__main__":
    main()
