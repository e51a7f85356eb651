#!/usr/bin/env python
"""Stress of the persistent kernel on tiny / ragged systems (few rows per CTA, CTAs without rows,
re-balancing with 0..3 rows): every configuration in its own process (a device-side timeout
faults the CUDA context), repeated; prints the configurations that failed.
    python profiles/stress_persist.py [--reps 10]"""
import argparse
import importlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(n, variant, balance, reps, iters):
    sys.path.insert(0, ROOT)
    cgb = importlib.import_module("conjugate-gradient_b200")
    import numpy as np
    with cgb.Context(n, 0, 1, 0) as ctx:
        ctx.set_option("gemv_variant", variant)
        ctx.set_option("schedule", 1)
        ctx.set_option("balance", balance)
        assert ctx.get_option("schedule_in_use") == 1
        ctx.generate_lap2d()
        ctx.set_rhs(cgb.init_source_term(n))
        ref = None
        for rep in range(reps):
            x = np.zeros(n)
            info, hist = ctx.solve(x, max_iter=iters, tol=1e-10, history=True)
            if ref is None:
                ref = (x.copy(), hist.copy(), info.k)
            else:
                assert info.k == ref[2] and np.array_equal(hist, ref[1], equal_nan=True) and \
                    np.array_equal(x, ref[0], equal_nan=True), "run-to-run difference"
    print("ok")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--sizes", default="255,1,17,148,149,257,300,600,1001,2369")
    ap.add_argument("--child", nargs=5, type=int, default=None)
    a = ap.parse_args()
    if a.child:
        child(*a.child)
        return 0
    sys.path.insert(0, ROOT)
    cgb = importlib.import_module("conjugate-gradient_b200")
    names = cgb.gemv_variants()
    persist = [i for i, nm in enumerate(names) if nm in
               ("tma_w8r2c512s3", "tma_w4r4c512s3", "tma_w8r1c1024s3", "tma_w4r2c1024s3", "tma_w4r1c2048s2")]
    SIZES = [int(t) for t in a.sizes.split(",")]
    bad = []
    env = dict(os.environ, CGB_SPIN_TIMEOUT_MS="2000")
    for n in SIZES:
        for v in persist:
            for bal in (1, 0):
                r = subprocess.run([sys.executable, __file__, "--child", str(n), str(v), str(bal), str(a.reps),
                                    str(min(n, 90))], capture_output=True, text=True, env=env, timeout=300)
                ok = r.returncode == 0 and "ok" in r.stdout
                print(json.dumps({"n": n, "variant": names[v], "balance": bal, "ok": ok,
                                  "err": "" if ok else (r.stderr.strip().splitlines() or ["?"])[-1][60:100] + " ... " + (r.stderr.strip().splitlines() or ["?"])[-1][-60:]}), flush=True)
                if not ok:
                    bad.append((n, names[v], bal))
    print("FAILED:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
