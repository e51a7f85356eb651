#!/usr/bin/env python
"""Times the full CG iteration (device-resident, CUDA events) for a few variants.
    python profiles/quick_iter.py N [variants...]"""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
cgb = importlib.import_module("conjugate-gradient_b200")

def source_term(n):
    i = np.arange(n, dtype=np.float64)
    return -2.0 * i * np.pi * np.pi * np.sin(10.0 * np.pi * i / n) ** 2

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
variants = [int(v) for v in sys.argv[2:]] or [0]
names = cgb.gemv_variants()
with cgb.Context(n) as ctx:
    ctx.generate_lap2d()
    ctx.set_rhs(source_term(n))
    for v in variants:
        for graph in (1, 0):
            ctx.set_option("gemv_variant", v); ctx.set_option("graph", graph)
            ctx.solve_begin(np.zeros(n), 1000, 1e-10)
            ctx.iterate(20)
            ms = ctx.iterate(100)
            info = ctx.solve_end()
            it_ms = ms / 100
            print(json.dumps(dict(n=n, variant=names[v], graph=graph, ms_per_iter=it_ms,
                                  it_per_s=1e3 / it_ms, gbs=8.0 * n * n / it_ms / 1e6, k=info.k)))
