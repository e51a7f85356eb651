"""CPU tests of the N > 1 host logic (world_size 2, gloo): the wiring that replaces MPI_Init,
the sharding rule, the oracle's emulated ranks, and bench.py's reference arm under torchrun."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(nproc, script_args, port, timeout=300):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port)] + script_args
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)


def test_wiring_world2_gloo(tmp_path, cgb):
    """Both ranks receive rank 0's id and every rank's exchange blob in rank order; shards
    follow partition_matrix (remainder to the last rank)."""
    out = tmp_path / "w.json"
    r = _torchrun(2, [os.path.join(ROOT, "tests", "mp_solve_worker.py"), "--backend", "gloo",
                      "--size", "1001", "--out", str(out)], 29741)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    recs = sorted(json.loads(out.read_text()), key=lambda d: d["rank"])
    assert [d["rank"] for d in recs] == [0, 1]
    for d in recs:
        assert d["uid"] == "U" * 128
        assert d["blob_ranks"] == [0, 1]
    starts, counts = cgb.partition(1001, 2)
    assert [d["first"] for d in recs] == starts and [d["rows"] for d in recs] == counts == [500, 501]


def test_shard_of_is_partition_matrix(cgb):
    import importlib
    wiring = importlib.import_module("conjugate-gradient_b200.wiring")
    for n, p in [(10, 1), (10, 3), (40000, 8), (56569, 8), (7, 7)]:
        starts, counts = cgb.partition(n, p)
        assert [wiring.shard_of(n, r, p) for r in range(p)] == list(zip(starts, counts))


def test_reference_arm_under_torchrun_world2(tmp_path):
    """`bench.py --impl reference` launched like the product arm: rank 0 alone runs and prints
    ONE JSON line, the other rank exits 0 silently."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "cgsolver_ref")):
        import pytest
        pytest.skip("oracle/_ref not built (needs /root/reference once)")
    r = _torchrun(2, [os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                      "--steps", "2", "--warmup", "1", "--size", "600"], 29743)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    assert d["metric"] == "cg_iterations_per_second" and d["unit"] == "iterations/s"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] == "reference"


def test_product_arm_refuses_to_run_without_gpu():
    """No CPU fallback: bench.py's product arm exits non-zero on a box without a GPU."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
