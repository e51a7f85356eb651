#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json: CG iterations/s and mat-vec HBM GB/s
(fraction of roofline) on generate_lap2d_matrix systems, 1/2/4/8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # N = 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W          # N > 1, one rank per GPU
    python bench.py --impl reference ...                                # the reference's CPU path

Workload (config.workload): BASELINE.json configs[2] -- generate_lap2d_matrix N = 40000, dense
fp64, fixed 200 CG iterations, A row-sharded over the N GPUs by the reference's
partition_matrix rule (strong scaling: the total work does not grow with N).  `--workload weak`
runs configs[3] instead (N = 20000 * sqrt(G), 3.2 GB of A per GPU).

A "step" is one pass of the hot path over that system: 200 iterations of CGSolver::solve's loop
(code/MPI/cg.cc:96-137) = 200 x (mat-vec + gather + fused x/r update + fused p update).
  value : iterations/s over K steps with everything resident in HBM (cgb_solve_begin outside,
          cgb_iterate inside the timed region; device time from CUDA events on the solver's
          stream, max over ranks).
  e2e   : iterations/s through the call a user of the reference makes -- CGSolver::solve, i.e.
          cgb_set_rhs + cgb_solve + cgb_residual_check (the DEBUG block the reference keeps in
          its timed region, cg_main.cc:53-55) -- with x0 / b / x in pinned HOST buffers, the
          host<->device copies inside the timed region (wall clock, max over ranks).
  roofline : the mat-vec kernel; algorithmic bytes per launch = 8 * rows_g * N (A read once),
          duration = average of per-launch CUDA-event pairs inside a real iteration loop.
  cpu_baseline : the UNMODIFIED reference MPI solver (oracle/_ref/cgsolver_ref, built from
          /root/reference by oracle/Makefile) on this box's host cores, OpenBLAS threads = all
          cores, on a bounded sample (fewer iterations of the same N).

The product path is libcgb200.so through its C ABI only.  oracle/ is executed solely for the
cpu_baseline leg and for --impl reference.
"""
import argparse
import importlib
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "cg_iterations_per_second"
UNIT = "iterations/s"
ITERS_PER_STEP = 200          # configs[2] / configs[3]: fixed 200 iterations
NOMINAL_HBM_GBS = 8000.0      # north_star's denominator ("B200 peak ~8 TB/s")
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md, if MEASURED_PEAKS.json is absent


def workload_n(kind: str, gpus: int) -> int:
    if kind == "weak":
        return {1: 20000, 2: 28284, 4: 40000, 8: 56568}.get(gpus, int(20000 * math.sqrt(gpus)))
    return 40000


def base_config(kind: str, n: int, gpus: int, iters: int = ITERS_PER_STEP) -> dict:
    return {
        "workload": ("generate_lap2d_matrix N=%d dense fp64, fixed %d CG iterations per step, "
                     "b=init_source_term(1/N), x0=0, tol=1e-10" % (n, iters)),
        "baseline_config": "configs[2]" if kind == "strong" else "configs[3]",
        "n": n,
        "iterations_per_step": iters,
        "matrix_bytes": 8 * n * n,
        "sharding": "rows/%d (partition_matrix, cg.cc:236-268)" % gpus,
        "l2": "inputs larger than L2 (A shard %.1f GB per GPU >> 126 MB)" % (8.0 * n * n / gpus / 1e9),
    }


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, device copy)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------ reference arm
def ref_binary(ranks: int = 1):
    return os.path.join(ROOT, "oracle", "_ref", "cgsolver_ref" if ranks == 1 else "cgsolver_ref_mp")


def run_reference_cpu(n: int, iters: int, threads: int, skip: int = 0, ranks: int = 1):
    """Runs the unmodified reference MPI solver for `iters` iterations and returns steady-state
    numbers from rank 0's dgemv timestamps.  ranks == 1: one rank, OpenBLAS with `threads`
    threads.  ranks > 1: `ranks` MPI ranks forked on this host (oracle/ref_shim/mpi_fork.cc),
    threads // ranks OpenBLAS threads each; every rank holds the full matrix, as in the reference."""
    exe = ref_binary(ranks)
    if not os.path.exists(exe):
        raise FileNotFoundError(exe + " (built by oracle/Makefile from /root/reference)")
    with tempfile.TemporaryDirectory() as td:
        per = max(1, threads // ranks)
        env = dict(os.environ, CGREF_BLAS="auto", CGREF_GEMV_TIMES=os.path.join(td, "t"),
                   CGREF_NP=str(ranks), OPENBLAS_NUM_THREADS=str(per), OMP_NUM_THREADS=str(per))
        t0 = time.time()
        res = subprocess.run([exe, str(n), os.path.join(td, "results.txt"), str(iters)], env=env,
                             capture_output=True, text=True)
        wall = time.time() - t0
        if res.returncode != 0:
            raise RuntimeError("reference solver failed: " + res.stderr[-400:])
        stamps = np.fromfile(os.path.join(td, "t"), dtype=np.float64)
        row = open(os.path.join(td, "results.txt")).read().strip()
    # stamps: [init dgemv, loop dgemv x executed, DEBUG dgemv, exit]
    executed = len(stamps) - 3
    assert executed >= 1, (len(stamps), res.stdout)
    lo = 1 + min(skip, executed - 1)
    timed = executed - (lo - 1)
    loop_s = float(stamps[1 + executed] - stamps[lo])
    blas = [ln for ln in res.stderr.splitlines() if ln.startswith("[cgref]")]
    return dict(iters=executed, timed_iters=timed, loop_seconds=loop_s, it_per_s=timed / loop_s,
                solve_seconds=float(row.split(",")[2]), wall_seconds=wall,
                blas=blas[0] if blas else "?", debug_line=res.stdout.strip())


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = workload_n(args.workload, args.gpus) if args.n is None else args.n
    cores = os.cpu_count() or 1
    if n >= 46341:
        print(json.dumps({"impl": "reference", "unavailable": "the reference indexes with int (matrix.hh:17, "
                          "cg.cc:80): N = %d >= 46341 overflows i*m_n+j, it cannot run this size" % n}))
        return 0
    # One run of the unmodified reference covers all W + K steps.  A step is the workload's own
    # 200 iterations whenever the whole run then stays within ~2000 iterations (a few minutes on the
    # host cores); otherwise fewer iterations per step -- config.iterations_per_step says which.
    nsteps = max(1, args.steps + args.warmup)
    per_step = ITERS_PER_STEP if args.iters is None else args.iters
    if per_step * nsteps > args.ref_max_iters:
        per_step = max(1, args.ref_max_iters // nsteps)
    total = per_step * nsteps
    ranks = args.cpu_ranks if args.cpu_ranks else 1
    try:
        r = run_reference_cpu(n, total, cores, skip=per_step * args.warmup, ranks=ranks)
    except Exception as e:  # the oracle always exists in a built tree; say why if it does not
        print(json.dumps({"impl": "reference", "unavailable": str(e)[:200]}))
        return 0
    sample = ("unmodified reference code/MPI solver (oracle/_ref), %d rank(s) x %d thread(s), %s; N=%d; "
              "%d iterations per step, %d warm-up + %d timed steps as ONE solve of %d iterations "
              "(cost per iteration does not depend on the iteration index); steady-state loop time "
              "from dgemv timestamps" % (ranks, max(1, cores // ranks), r["blas"], n, per_step,
                                         args.warmup, args.steps, total))
    value = r["it_per_s"]
    cfg = base_config(args.workload, n, args.gpus, per_step)
    cfg["note"] = ("the reference replicates the FULL matrix on every rank (cg.cc:169): %d rank(s) "
                   "fit this host's memory budget; the row sharding of config.sharding is the GPU arm's" % ranks)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * r["loop_seconds"] / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.workload == "strong" else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic (generate_lap2d_matrix, init_source_term)",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gemv_gbs": 8.0 * n * n * value / 1e9,
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.path = tempfile.NamedTemporaryFile(prefix="clocks_", suffix=".csv", delete=False).name
        self.proc = None
        try:
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(device), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.out,
                stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.close()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
                power.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)),
                "power_w_max": float(max(power)), "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------ our arm
def pinned(n):
    import torch
    return torch.zeros(n, dtype=torch.float64).pin_memory().numpy()


def product_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("bench.py --gpus %d must be launched with torch.distributed.run "
                     "--nproc-per-node %d (one rank per GPU)" % (args.gpus, args.gpus))
        sys.exit("WORLD_SIZE (%d) != --gpus (%d)" % (world, args.gpus))
    import torch
    cgb = importlib.import_module("conjugate-gradient_b200")   # fails loudly without the .so
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; the product has no CPU fallback "
                 "(use --impl reference for the CPU path)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v: float) -> float:
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v: float) -> float:
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    n = workload_n(args.workload, world) if args.n is None else args.n
    ctx = cgb.Context(n, rank, world, local)
    if world > 1:
        # wiring (replaces MPI_Init): NCCL id for the ncclAllGather baseline, exchange blobs
        # (CUDA IPC) for the fused peer-store exchange; --exchange 0/1 picks one (default fused)
        wiring = importlib.import_module("conjugate-gradient_b200.wiring")
        wiring.wire(ctx, rank, world, dist, cgb.unique_id)
    if args.variant is not None:
        ctx.set_option("gemv_variant", args.variant)
    for key in ("graph", "graph_unroll", "poll_every", "exchange", "pdl", "l2_prefetch", "schedule"):
        v = getattr(args, key)
        if v is not None:
            ctx.set_option(key, v)
    exchange_name = ("none (1 GPU)" if world == 1 else
                     ["ncclAllGather", "fused peer stores + flags in the mat-vec kernel"][ctx.get_option("exchange")])
    ctx.generate_lap2d()                          # device-side generate_lap2d_matrix
    # init_source_term (cg.cc:218-234) on the host, bit-identical to the reference's libm loop
    b_host = pinned(n)
    cgb.init_source_term(n, 1.0 / n, out=b_host)
    x_host = pinned(n)
    ctx.set_rhs(b_host)
    iters = ITERS_PER_STEP if args.iters is None else args.iters
    tuned = None
    if args.variant is None and not args.no_autotune:
        tuned = ctx.autotune()                    # one-off shape selection, outside every timed region
        barrier()
    lay = ctx.layout()
    persistent = ctx.get_option("schedule_in_use") == 1
    variant_name = cgb.gemv_variants()[ctx.get_option("gemv_variant")]

    # ---- value: device-resident steps
    def step():
        ctx.solve_begin(None, iters, 1e-10, False)   # x0 = 0 set on the device; not timed
        ms = ctx.iterate(iters)
        info = ctx.solve_end(None, None)
        return ms, info

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    l0 = ctx.launch_count()
    t_wall0 = time.perf_counter()
    dev_ms = 0.0
    info = None
    for _ in range(args.steps):
        ms, info = step()
        dev_ms += ms
    barrier()
    wall_s = time.perf_counter() - t_wall0
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    dev_ms = allmax(dev_ms)
    launches_all = int(allsum(float(launches)))
    done_iters = int(info.iterations)
    value = args.steps * done_iters / (dev_ms * 1e-3)
    ms_per_step = dev_ms / args.steps
    sqrt_rsold = math.sqrt(info.rsold)

    # ---- roofline of the dominant kernel.  Persistent schedule: ONE kernel runs the whole loop, its
    # launch duration is the CUDA-event time of the timed region above (one launch per step).
    # Graph schedule: the mat-vec kernel, per-launch event pairs inside a real loop (profile mode).
    rows_max = max(cgb.partition(n, world)[1])
    shard_bytes = 8.0 * rows_max * n               # slowest rank's shard, A read exactly once per iteration
    peak, peak_src = measured_peak()

    def graph_schedule_numbers():
        """The graph schedule (4 kernels per iteration) on the same resident system (A/B beside the headline)."""
        keep = ctx.get_option("schedule")
        ctx.set_option("schedule", 0)
        step()
        ms, inf = step()
        it_s = int(inf.iterations) / (allmax(ms) * 1e-3)
        ctx.set_option("profile", 1)
        ctx.solve_begin(None, iters, 1e-10, False)
        ctx.iterate(iters)
        ctx.solve_end(None, None)
        g_ms, g_n = ctx.last_gemv_timing()
        ctx.set_option("profile", 0)
        ctx.set_option("schedule", keep)
        return it_s, allmax(g_ms), int(g_n)

    graph_it_s, gemv_ms, gemv_launches = graph_schedule_numbers()
    alone_ms = allmax(ctx.bench_gemv(-1, 20))
    read_ms = allmax(ctx.bench_read(20))
    if persistent:
        kernel = "cg_persist_kernel (%s): whole CG loop, %d iterations per launch" % (variant_name, done_iters)
        bytes_per_launch = shard_bytes * done_iters
        launch_ms = ms_per_step
        launches_timed = args.steps
    else:
        kernel = "gemv_tma_kernel (%s)" % variant_name
        bytes_per_launch = shard_bytes
        launch_ms = gemv_ms
        launches_timed = gemv_launches
    achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "gemv_traffic.json")
    if os.path.exists(tr_path):
        try:
            tr = json.load(open(tr_path))
            key = "%s_n%d_rows%d" % ("persist" if persistent else "gemv", n, rows_max)
            per_iter = tr.get(key, {}).get("dram_bytes_per_iteration")
            if per_iter is not None:
                traffic = per_iter * (done_iters if persistent else 1)
        except Exception:
            traffic = None
    timeline = None
    if persistent:
        # production-schedule phase times from the %globaltimer stamps of one traced step
        try:
            ctx.set_option("trace", min(iters, 64))
            step()
            rec, seen = ctx.trace_read(0)
            ctx.set_option("trace", 0)
            g = rec.astype(np.int64)[4:]
            first, rowsd = g[:, :, 3], g[:, :, 5]
            last = rowsd.max(axis=1)
            timeline = {
                "matvec_phase_us_median_cta": float(np.median(rowsd - first)) / 1e3,
                "matvec_phase_us_slowest_cta": float(np.median((rowsd - first).max(axis=1))) / 1e3,
                "vector_phases_us": float(np.median(np.median(first[1:], axis=1) - last[:-1])) / 1e3,
                "iteration_us": float(np.median(np.diff(last))) / 1e3,
                "source": "%globaltimer stamps of every CTA (cgb_trace_read), rank 0, one traced step",
            }
        except Exception as e:  # diagnostic only
            timeline = {"error": str(e)[:120]}
    roofline = {
        "bound": "hbm", "kernel": kernel,
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": bytes_per_launch, "launch_ms": launch_ms,
        "launches_timed": int(launches_timed),
        "frac_of_nominal_8TBs": achieved / NOMINAL_HBM_GBS,
        "matvec_kernel_alone_gbs": shard_bytes / (alone_ms * 1e-3) / 1e9,
        "matvec_kernel_in_graph_loop_gbs": shard_bytes / (gemv_ms * 1e-3) / 1e9,
        "ldg_read_stream_gbs": shard_bytes / (read_ms * 1e-3) / 1e9,
        "full_iteration_gbs_per_gpu": 8.0 * n * n / world * value / 1e9,
        "full_iteration_frac_of_nominal": 8.0 * n * n / world * value / 1e9 / NOMINAL_HBM_GBS,
        "graph_schedule_it_per_s": graph_it_s,
        "timeline": timeline,
    }

    # ---- e2e: the reference-facing solve with host buffers (also the run the parity check reads)
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_step(history=False):
        x_host[:] = 0.0
        ctx.set_rhs(b_host)                                   # H2D 8N
        inf, hist = ctx.solve(x_host, max_iter=iters, tol=1e-10, history=history)  # H2D 8N (x0), D2H 8N (x)
        nx, rr = ctx.residual_check()                         # DEBUG block, D2H 2 doubles
        return inf, nx, rr, hist

    inf, nx, rr, hist = e2e_step(history=True)                # warm-up, with the r'r history for `check`
    x_check = x_host.copy()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        inf, nx, rr, _ = e2e_step()
    barrier()
    e2e_s = allmax(time.perf_counter() - t0)
    e2e_value = e2e_steps * int(inf.iterations) / e2e_s
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 16 * n,
           "d2h_bytes_per_step": 8 * n + 16 + 48, "steps": e2e_steps,
           "ms_per_step": 1e3 * e2e_s / e2e_steps,
           "call": "cgb_set_rhs + cgb_solve + cgb_residual_check (= CGSolver::solve incl. DEBUG block), pinned host x0/b/x",
           "norm_x": nx, "rel_resid": rr}

    # ---- parity check where the driver sees it: the full r'r history, k and x of this run against
    # the golden output of the UNMODIFIED reference (same N, same iteration cap), and a digest
    # of (x, history) that must be identical on every rank (all ranks compute every scalar
    # redundantly from identical data in identical order)
    check = parity_check(n, iters, inf, hist, x_check, nx, rr, b_host, dist, world)

    if dist is not None:
        dist.barrier()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    if rank != 0:
        return 0

    # ---- CPU baseline (rank 0 alone, after the other ranks have left: all host cores are free)
    cpu_baseline = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        ranks = args.cpu_ranks if args.cpu_ranks else 1
        if n >= 46341:
            cpu_baseline = {"value": None, "unit": UNIT, "cores": cores, "kind": "reference",
                            "sample": "unavailable: the reference indexes with int (matrix.hh:17, cg.cc:80); "
                                      "N = %d >= 46341 overflows i*m_n+j, it cannot run this size" % n}
        else:
            try:
                r = run_reference_cpu(n, args.cpu_iters, cores, ranks=ranks)
                cpu_baseline = {
                    "value": r["it_per_s"], "unit": UNIT, "cores": cores, "kind": "reference",
                    "sample": ("unmodified reference code/MPI solver (oracle/_ref), %d rank(s) x %d thread(s), %s, "
                               "N=%d, %d of the %d iterations; steady-state loop time from dgemv timestamps "
                               "(%.2f s loop, %.1f s process incl. its -O0 matrix generation); the reference "
                               "replicates the full matrix on every rank, so rank count is bounded by host memory"
                               % (ranks, max(1, cores // ranks), r["blas"], n, r["iters"], iters,
                                  r["loop_seconds"], r["wall_seconds"])),
                    "gemv_gbs": 8.0 * n * n * r["it_per_s"] / 1e9,
                }
            except Exception as e:
                cpu_baseline = {"value": None, "unit": UNIT, "cores": cores, "kind": "reference",
                                "sample": "unavailable: " + str(e)[:200]}

    cfg = base_config(args.workload, n, world, iters)
    cfg.update({"gemv_variant": variant_name, "nblk": lay.nblk,
                "schedule": ("persistent cooperative kernel (1 launch per step)" if persistent
                             else "CUDA graph of 4 kernels per iteration"),
                "autotune": tuned, "options": ctx_opts(args), "parallelism": "rows%d" % world,
                "exchange": exchange_name})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if args.workload == "strong" else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic (generate_lap2d_matrix, init_source_term)",
        "config": cfg, "ms_per_iteration": ms_per_step / done_iters,
        "gemv_gbs_per_gpu": shard_bytes * value / 1e9, "wall_s_timed_region": wall_s,
        "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu_baseline,
        "gpu_launches": launches_all, "clocks": clocks, "check": check,
    }
    print(json.dumps(line))
    return 0


GOLDEN = {20000: "full_n20000_it200.npz", 28284: "full_n28284_it200.npz", 40000: "full_n40000_it200.npz"}


def parity_check(n, iters, info, hist, x, nx, rr, b, dist, world):
    """Compares this run with the reference's own output for the same system (tests/golden/, produced
    by tests/golden/make_golden.py from the unmodified reference; data only -- the oracle is not
    involved) under north_star's tolerances, and the ranks with each other bit for bit."""
    import hashlib
    out = {"k": int(info.k), "iterations": int(info.iterations), "sqrt_rsold": math.sqrt(info.rsold)}
    digest = hashlib.sha256(x.tobytes() + hist.tobytes() + np.array([nx, rr]).tobytes()).hexdigest()
    digests = [digest]
    if dist is not None:
        digests = [None] * world
        dist.all_gather_object(digests, digest)
    out["ranks_bitwise_identical"] = len(set(digests)) == 1
    out["digest"] = digest[:16]
    # size-independent property: the recursive residual of CG equals the true residual ||A x - b||
    # the DEBUG block recomputes with a fresh mat-vec
    true_resid = rr * float(np.linalg.norm(b))
    out["recursive_vs_true_residual_rel"] = abs(true_resid - math.sqrt(info.rsold)) / true_resid
    ok = out["ranks_bitwise_identical"] and out["recursive_vs_true_residual_rel"] <= 1e-9
    gname = GOLDEN.get(n)
    gpath = os.path.join(ROOT, "tests", "golden", gname) if gname else None
    if gpath and os.path.exists(gpath) and iters == ITERS_PER_STEP:
        g = np.load(gpath)
        h_ref, x_ref = g["openblas_hist"][:iters], g["openblas_x"]
        m = min(len(hist), len(h_ref))
        rel = np.abs(np.sqrt(hist[:m]) - np.sqrt(h_ref[:m])) / np.sqrt(h_ref[:m])
        out.update({
            "golden": gname, "reference_k": int(g["openblas_k"]),
            "reference_printed_sqrt_rsold": float(g["openblas_resid_print"]),
            "history_values_compared": int(m),
            "history_norm_max_rel_err": float(rel.max()), "history_tolerance": 1e-10,
            "x_rel_err": float(np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref)), "x_tolerance": 1e-9,
        })
        ok = (ok and m == iters and abs(out["k"] - out["reference_k"]) <= 1
              and out["history_norm_max_rel_err"] <= 1e-10 and out["x_rel_err"] <= 1e-9)
    else:
        out["golden"] = None
        out["note"] = (("no reference run exists for this size: N >= 46341 overflows the reference's int index "
                        "(matrix.hh:17)" if n >= 46341 else
                        "no golden reference output is stored for this (N, iteration count)") +
                       "; checked by the recursive-vs-true residual identity and rank agreement; the "
                       "same kernels are compared bitwise with the 64-bit oracle in tests/")
    out["ok"] = bool(ok)
    return out


def ctx_opts(args):
    return {k: getattr(args, k) for k in ("graph", "graph_unroll", "poll_every", "exchange", "pdl", "l2_prefetch", "schedule")
            if getattr(args, k) is not None} or "defaults"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", choices=["strong", "weak"], default="strong")
    ap.add_argument("--size", dest="n", type=int, default=None, help="override N (not a bench line)")
    ap.add_argument("--iters", type=int, default=None, help="override iterations per step")
    ap.add_argument("--variant", type=int, default=None)
    ap.add_argument("--graph", type=int, default=None)
    ap.add_argument("--graph-unroll", dest="graph_unroll", type=int, default=None)
    ap.add_argument("--poll-every", dest="poll_every", type=int, default=None)
    ap.add_argument("--exchange", type=int, default=None)
    ap.add_argument("--pdl", type=int, default=None)
    ap.add_argument("--l2-prefetch", dest="l2_prefetch", type=int, default=None)
    ap.add_argument("--schedule", type=int, default=None, help="1 persistent kernel (default), 0 CUDA graph of 4 kernels per iteration")
    ap.add_argument("--no-autotune", action="store_true", help="keep the default mat-vec tile shape")
    ap.add_argument("--ref-max-iters", dest="ref_max_iters", type=int, default=1000,
                    help="--impl reference: cap on the iterations of the whole run (bounded sample)")
    ap.add_argument("--cpu-iters", dest="cpu_iters", type=int, default=40)
    ap.add_argument("--cpu-ranks", dest="cpu_ranks", type=int, default=1,
                    help="MPI ranks of the CPU reference (forked on this host; default 1 rank x all threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        print("bench.py: note: fewer than 3 warm-up steps requested", file=sys.stderr)
    if args.impl == "reference":
        return reference_arm(args)
    return product_arm(args)


if __name__ == "__main__":
    sys.exit(main())
