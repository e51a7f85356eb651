"""GPU tests of the drop-in command line (conjugate-gradient_b200/host/cgsolver): both reference
forms, the DEBUG stdout line, the results-file rows (code/MPI/cg_main.cc:57-64,
code/CUDA/cg_main.cc:54-60) -- numbers checked against the CPU oracle."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CGSOLVER = os.path.join(ROOT, "conjugate-gradient_b200", "host", "cgsolver")
LINE = re.compile(r"^\t\[STEP (\d+)\] residual = (\d\.\d{6}e[+-]\d\d), \|\|x\|\| = (\d\.\d{6}e[+-]\d\d), "
                  r"\|\|Ax - b\|\|/\|\|b\|\| = (\d\.\d{6}e[+-]\d\d)$")


def _run(args, env=None, timeout=300):
    return subprocess.run([CGSOLVER] + args, capture_output=True, text=True, timeout=timeout,
                          env=dict(os.environ, **(env or {})))


def _step_line(stdout):
    """The DEBUG line (NCCL may print its version banner on stdout when NCCL_DEBUG is set)."""
    hits = [ln for ln in stdout.splitlines() if ln.startswith("\t[STEP")]
    assert len(hits) == 1, stdout
    return hits[0]


def test_form1_generated(O, tmp_path):
    """cgsolver N outfile [max_iter]: DEBUG line identical to the oracle's, row `n,psize,seconds`
    appended (not truncated)."""
    out = tmp_path / "res.txt"
    out.write_text("1024,1,0.0610\n")                       # an earlier row must survive
    js = tmp_path / "side.jsonl"
    r = _run(["1024", str(out)], env={"CGB_JSON": str(js)})
    assert r.returncode == 0, r.stderr
    ref = O.solve(O.generate_lap2d(1024), O.init_source_term(1024), nranks=1, nblk=148)
    lines = r.stdout.splitlines()
    assert len(lines) == 1 and LINE.match(lines[0])
    assert lines[0] == O.debug_line(ref.k, ref.rsold, ref.norm_x, ref.rel_resid)
    rows = out.read_text().splitlines()
    assert rows[0] == "1024,1,0.0610" and len(rows) == 2
    n, psize, sec = rows[1].split(",")
    assert (int(n), int(psize)) == (1024, 1) and 0 < float(sec) < 60
    side = json.loads(js.read_text())
    assert side["k"] == ref.k and side["converged"] is True
    # weak-scaling form: max_iter caps the loop; k printed = max_iter
    r = _run(["1448", str(out), "200"])
    assert r.returncode == 0
    ref = O.solve(O.generate_lap2d(1448), O.init_source_term(1448), max_iter=200, nranks=1, nblk=148)
    assert r.stdout.splitlines()[0] == O.debug_line(200, ref.rsold, ref.norm_x, ref.rel_resid)
    assert len(out.read_text().splitlines()) == 3


def test_form2_matrix_market(O, tmp_path):
    """cgsolver file.mtx NUM_THREADS BLOCK_WIDTH true/false outfile, with cg.run's comma-
    terminated tokens ("64,"), the sticky-scientific time line and the plain results row."""
    mtx = str(tmp_path / "lap30.mtx")
    O.write_lap2d_5pt_mtx(mtx, 30)
    A = O.read_mtx_dense(mtx)
    n = A.shape[0]
    ref = O.solve(A, O.init_source_term(n), max_iter=n, nranks=1, nblk=148)
    out = tmp_path / "res2.txt"
    for nt, bw, t in (("64,", "16,", "true"), ("1024", "4096", "false")):
        r = _run([mtx, nt, bw, t, str(out)])
        assert r.returncode == 0, r.stderr
        lines = r.stdout.splitlines()
        assert lines[0] == O.debug_line(ref.k, ref.rsold, ref.norm_x, ref.rel_resid)
        assert re.match(r"^Time for CG \(dense solver\)  = \d\.\d{6}e[+-]\d\d \[s\]$", lines[1])
    rows = out.read_text().splitlines()
    assert [row.split(",")[:2] for row in rows] == [["64", "16"], ["1024", "4096"]]
    for row in rows:
        assert re.match(r"^\d+,\d+,\d*\.?\d+(e-\d+)?$", row)
        assert 0 < float(row.split(",")[2]) < 60


def test_form2_compat_kernel_mode(O, tmp_path):
    """CGB_KERNEL=compat: the reference's own launch topologies, NUM_THREADS / BLOCK_WIDTH
    literal; the printed numbers are the oracle's for that chunked order."""
    mtx = str(tmp_path / "lap30.mtx")
    O.write_lap2d_5pt_mtx(mtx, 30)
    A = O.read_mtx_dense(mtx)
    n = A.shape[0]
    out = tmp_path / "res.txt"
    for nt, bw, t in (("64", "16", "true"), ("8", "900", "false")):
        with O.gemv_chunk(int(bw)):
            ref = O.solve(A, O.init_source_term(n), max_iter=n, nranks=1, nblk=148)
        r = _run([mtx, nt, bw, t, str(out)], env={"CGB_KERNEL": "compat"})
        assert r.returncode == 0, r.stderr
        assert _step_line(r.stdout) == O.debug_line(ref.k, ref.rsold, ref.norm_x, ref.rel_resid)
    assert [row.split(",")[:2] for row in out.read_text().splitlines()] == [["64", "16"], ["8", "900"]]


def test_multi_gpu_cli_psize_column(O, tmp_path, cgb):
    if cgb.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = tmp_path / "res.txt"
    for mode in ("fused", "nccl"):
        r = _run(["2048", str(out), "150"], env={"CGB_GPUS": "2", "CGB_EXCHANGE": mode})
        assert r.returncode == 0, r.stderr
        ref = O.solve(O.generate_lap2d(2048), O.init_source_term(2048), max_iter=150, nranks=2, nblk=148)
        assert _step_line(r.stdout) == O.debug_line(150, ref.rsold, ref.norm_x, ref.rel_resid)
    assert [row.split(",")[:2] for row in out.read_text().splitlines()] == [["2048", "2"]] * 2
