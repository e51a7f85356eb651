#!/usr/bin/env python
"""Sweeps the mat-vec variants of libcgb200.so on the resident generator matrix and prints
achieved GB/s (algorithmic bytes 8*rows*N per launch / CUDA-event time), next to a plain
read-only streaming kernel over the same shard.  Run on a B200:
    python profiles/tune_gemv.py [--out gpurun_out/tune_gemv.jsonl]
"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cgb = importlib.import_module("conjugate-gradient_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "tune_gemv.jsonl"))
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--cases", default="40000:1,40000:8,20000:1,10000:1")
    ap.add_argument("--only", default=None, help="comma-separated variant names (default: all)")
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    names = cgb.gemv_variants()
    with open(args.out, "w") as out:
        for case in args.cases.split(","):
            n, world = (int(t) for t in case.split(":"))
            with cgb.Context(n, 0, world, 0) as ctx:   # rank 0's shard of a `world`-way split
                ctx.generate_lap2d()
                lay = ctx.layout()
                gb = 8.0 * lay.rows * n / 1e9
                ms = ctx.bench_read(args.reps)
                rec = dict(case=case, kernel="read_stream", ms=ms, gbs=gb / ms * 1e3,
                           frac_of_measured_copy=gb / ms * 1e3 / peak)
                print(json.dumps(rec)); out.write(json.dumps(rec) + "\n")
                for v, name in enumerate(names):
                    if args.only and name not in args.only.split(","):
                        continue
                    try:
                        ms = ctx.bench_gemv(v, args.reps)
                    except cgb.CgbError as e:
                        print(f"# {case} {name}: {e}")
                        continue
                    rec = dict(case=case, kernel=name, rows=lay.rows, n=n, ms=ms, gbs=gb / ms * 1e3,
                               frac_of_measured_copy=gb / ms * 1e3 / peak,
                               frac_of_8tbs=gb / ms * 1e3 / 8000.0)
                    print(json.dumps(rec)); out.write(json.dumps(rec) + "\n")
                    out.flush()


if __name__ == "__main__":
    main()
