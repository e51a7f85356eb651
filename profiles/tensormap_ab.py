#!/usr/bin/env python
"""A/B of the mat-vec producer: 1-D bulk copies (tma_*, SASS UBLKCP) against tiled copies through a
2-D tensor map (tm2d_*, SASS UTMALDG.2D), same tile shapes, same consumers, same bits.

    python profiles/tensormap_ab.py [--rounds 5] [--reps 20] [--once]   -> JSON lines on stdout

Shapes: the BASELINE configuration on one GPU with rows dividing evenly into row blocks (69856), the
configuration itself (70000: ragged row blocks -- a tensor-map box always carries TR rows), and the
8-GPU shard of it.  Variants are interleaved round by round so that clock / thermal drift hits both
alike; the figure is the best round of each.  --once: one launch of each variant (for ncu).
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cgb = importlib.import_module("conjugate-gradient_b200")

PAIRS = [("tma_w8r1c1024s3", "tm2d_w8r1c1024s3"), ("tma_w8r2c512s3", "tm2d_w8r2c512s3"),
         ("tma_w4r4c512s3", "tm2d_w4r4c512s3")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=5)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--shapes", default="69856:1,70000:1,70000:8")
    args = ap.parse_args()
    names = cgb.gemv_variants()
    for shape in args.shapes.split(","):
        n, world = (int(t) for t in shape.split(":"))
        with cgb.Context(n, 0, world, 0) as ctx:
            ctx.generate_lap2d()
            ctx.set_rhs(cgb.init_source_term(n))
            lay = ctx.layout()
            gbytes = lay.rows * n * 8 / 1e9
            if args.once:
                for pair in PAIRS[:1]:
                    for name in pair:
                        ctx.bench_gemv(names.index(name), 1)
                continue
            best = {}
            for rnd in range(args.rounds + 1):  # round 0 warms up
                for pair in PAIRS:
                    for name in pair:
                        ms = ctx.bench_gemv(names.index(name), args.reps)
                        if rnd:
                            best[name] = min(best.get(name, 1e9), ms)
            for a, b in PAIRS:
                print(json.dumps({"n": n, "world": world, "rows": int(lay.rows), "bulk_1d": a, "tensor_2d": b,
                                  "bulk_ms": round(best[a], 4), "tensor_ms": round(best[b], 4),
                                  "bulk_GBps": round(gbytes / best[a] * 1e3, 1),
                                  "tensor_GBps": round(gbytes / best[b] * 1e3, 1),
                                  "tensor_over_bulk_time": round(best[b] / best[a], 4)}), flush=True)


if __name__ == "__main__":
    main()
