#!/usr/bin/env python
"""Prints compact tables from the JSON-lines files the profiling scripts write
(ab_iter.py, trace_iter.py, tune_gemv.py):  python profiles/summ.py FILE..."""
import json
import sys

for f in sys.argv[1:]:
    print("==", f)
    for l in open(f):
        try:
            r = json.loads(l)
        except Exception:
            if l.strip():
                print("   ", l.rstrip()[:200])
            continue
        if "ms_per_iter" in r:
            print("  n=%d w=%d rows=%d %-44s %-18s sched=%d  %8.2f us/iter  %5.0f GB/s  same=%s" % (
                r["n"], r.get("world", 1), r.get("rows", r["n"]), str(r["opts"]), r.get("variant", "?"),
                r.get("schedule_in_use", -1), r["ms_per_iter"] * 1e3, r["gbs"], r["hist_equal_to_first_set"]))
        elif "iter_us" in r:
            print("  %s %s %s %s rank=%s untraced %.1f traced %.1f iter %.1f" % (
                r["case"], r["opts"], r["variant"], r.get("schedule"), r.get("rank"), r["untraced_us_per_iter"],
                r["traced_us_per_iter"], r["iter_us"]))
            for k in ("M_first_tile_to_rows_done_us", "rows_done_spread_max_med_us", "cta_M_us_p5_p50_p95_max",
                      "matvec_span_us", "gap_us", "dep_skew_us", "cta_busy_us_p5_p50_p95_max", "chain_us"):
                if k in r:
                    print("      ", k, r[k])
        elif "kernel" in r:
            print("  %-10s %-22s %8.1f us %6.0f GB/s" % (r["case"], r["kernel"], r["ms"] * 1e3, r["gbs"]))
