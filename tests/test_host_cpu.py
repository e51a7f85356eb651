"""CPU tests of the C++ host side (conjugate-gradient_b200/host): the Matrix Market reader
against the oracle's restatement AND against the reference's own reader (oracle/_ref/
mtx_dump_ref, the unmodified matrix_coo.cc / matrix.cc / mmio.c), and the command line's
argument handling.  No GPU needed: nothing here computes."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "conjugate-gradient_b200", "host")
DUMP = os.path.join(HOST, "mtx_dump")
CGSOLVER = os.path.join(HOST, "cgsolver")
REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "mtx_dump_ref")

FILES = {
    "sym_small": "%%MatrixMarket matrix coordinate real symmetric\n% comment\n%another\n"
                 "4 4 5\n1 1 4\n2 1 -1\n2 2 4.5e0\n4 3 -1.25E-2\n4 4 1\n",
    "general_dups": "%%MatrixMarket matrix coordinate real general\n3 3 5\n1 1 2.5\n3 1 -1\n"
                    "2 2 1e0\n1 1 7\n1 3 0.125\n",
    "blank_after_comments": "%%MatrixMarket matrix coordinate real symmetric\n%c\n\n2 2 2\n1 1 4\n2 1 -1\n",
    "spaces_and_case": "%%MatrixMarket MATRIX Coordinate REAL Symmetric\n  3   3   3\n  1  1   4\n"
                       " 2   1  -1\n   3 3    2\n",
    "integer_field": "%%MatrixMarket matrix coordinate integer general\n2 2 2\n1 2 3\n2 1 -4\n",
    "sym_offdiag_dup": "%%MatrixMarket matrix coordinate real symmetric\n3 3 3\n2 1 5\n1 2 6\n3 3 1\n",
}
BAD = {
    "array": "%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n",
    "no_banner": "1 1 1\n1 1 1\n",
    "bad_object": "%%MatrixMarket vector coordinate real general\n1 1 1\n1 1 1\n",
    "bad_field": "%%MatrixMarket matrix coordinate quaternion general\n1 1 1\n1 1 1\n",
    "short_banner": "%%MatrixMarket matrix coordinate\n1 1 1\n1 1 1\n",
}


def _need(path):
    if not os.path.exists(path):
        pytest.skip(os.path.relpath(path, ROOT) + " not built")


def _dump(tool, src, out, dense=False):
    return subprocess.run([tool, src, out] + (["--dense"] if dense else []), capture_output=True,
                          text=True, timeout=60)


def _read_coo(path):
    raw = open(path, "rb").read()
    m, n, nz, sym = np.frombuffer(raw[:16], dtype=np.int32)
    irn = np.frombuffer(raw[16:16 + 4 * nz], dtype=np.int32)
    jcn = np.frombuffer(raw[16 + 4 * nz:16 + 8 * nz], dtype=np.int32)
    a = np.frombuffer(raw[16 + 8 * nz:16 + 16 * nz], dtype=np.float64)
    return int(m), int(n), int(nz), int(sym), irn, jcn, a


def _read_dense(path):
    raw = open(path, "rb").read()
    m, n = np.frombuffer(raw[:8], dtype=np.int32)
    return np.frombuffer(raw[8:], dtype=np.float64).reshape(int(m), int(n))


@pytest.mark.parametrize("name", sorted(FILES))
def test_reader_matches_oracle_and_reference(O, tmp_path, name):
    _need(DUMP)
    src = tmp_path / (name + ".mtx")
    src.write_text(FILES[name])
    r = _dump(DUMP, str(src), str(tmp_path / "p.bin"))
    assert r.returncode == 0, r.stdout + r.stderr
    r = _dump(DUMP, str(src), str(tmp_path / "pd.bin"), dense=True)
    assert r.returncode == 0
    dense = _read_dense(str(tmp_path / "pd.bin"))
    assert np.array_equal(dense, O.read_mtx_dense(str(src)))      # the oracle's restatement
    if os.path.exists(REF_DUMP):                                    # the reference's own reader
        assert _dump(REF_DUMP, str(src), str(tmp_path / "r.bin")).returncode == 0
        assert _dump(REF_DUMP, str(src), str(tmp_path / "rd.bin"), dense=True).returncode == 0
        mine, ref = _read_coo(str(tmp_path / "p.bin")), _read_coo(str(tmp_path / "r.bin"))
        assert mine[:4] == ref[:4]
        for x, y in zip(mine[4:], ref[4:]):
            assert np.array_equal(x, y)
        assert np.array_equal(dense, _read_dense(str(tmp_path / "rd.bin")))


def test_reader_on_the_reference_fixture_family(O, tmp_path):
    """lap2D_5pt_n100.mtx regenerated entry for entry (10000 x 10000, 29800 triples)."""
    _need(DUMP)
    src = str(tmp_path / "lap.mtx")
    O.write_lap2d_5pt_mtx(src, 100)
    assert _dump(DUMP, src, str(tmp_path / "p.bin")).returncode == 0
    m, n, nz, sym, irn, jcn, a = _read_coo(str(tmp_path / "p.bin"))
    assert (m, n, nz, sym) == (10000, 10000, 29800, 1)
    assert irn.min() == 0 and jcn.max() == 9999 and set(np.unique(a)) == {-1.0, 4.0}
    if os.path.exists(REF_DUMP):
        assert _dump(REF_DUMP, src, str(tmp_path / "r.bin")).returncode == 0
        assert open(str(tmp_path / "p.bin"), "rb").read() == open(str(tmp_path / "r.bin"), "rb").read()


@pytest.mark.parametrize("name", sorted(BAD) + ["missing"])
def test_reader_rejects_like_the_reference(O, tmp_path, name):
    """Same message on stdout and exit(1) as matrix_coo.cc:12-33."""
    _need(DUMP)
    src = tmp_path / (name + ".mtx")
    if name != "missing":
        src.write_text(BAD[name])
    mine = _dump(DUMP, str(src), str(tmp_path / "p.bin"))
    assert mine.returncode == 1
    with pytest.raises(O.MtxError):
        O.read_mtx_dense(str(src))
    if os.path.exists(REF_DUMP):
        ref = _dump(REF_DUMP, str(src), str(tmp_path / "r.bin"))
        assert ref.returncode == mine.returncode
        assert ref.stdout == mine.stdout
    if name == "missing":
        assert mine.stdout == "Could not open matrix"
    if name == "array":
        assert "does not support Market Market type: [matrix array real general]" in mine.stdout


def test_cgsolver_usage_and_no_gpu_errors(tmp_path):
    _need(CGSOLVER)
    r = subprocess.run([CGSOLVER], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "Usage:" in r.stderr and "N outfile [max_iter]" in r.stderr
    assert "file.mtx NUM_THREADS BLOCK_WIDTH true/false outfile" in r.stderr
    src = tmp_path / "a.mtx"
    src.write_text(FILES["sym_small"])
    r = subprocess.run([CGSOLVER, str(src), "64"], capture_output=True, text=True, timeout=60)
    # form 2 needs all five arguments; its usage path exits 0 like the reference CUDA program's
    # usage() (code/CUDA/cg_main.cc:11-18), where the MPI form above returns 1 (cg_main.cc:22-26)
    assert r.returncode == 0 and "Usage:" in r.stderr and r.stdout == ""
    # unreadable matrix: the reader's message + exit(1), before any GPU work
    r = subprocess.run([CGSOLVER, str(tmp_path / "nope.mtx"), "64", "16", "true", str(tmp_path / "o.txt")],
                       capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and r.stdout == "Could not open matrix"
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:  # no CPU fallback: a clear error, no results row
        out = tmp_path / "res.txt"
        r = subprocess.run([CGSOLVER, "64", str(out)], capture_output=True, text=True, timeout=60)
        assert r.returncode == 2 and "cgb_create failed" in r.stderr
        assert not out.exists()


def test_reader_randomised_against_the_reference_reader(O, tmp_path):
    """40 seeded random coordinate files (general / symmetric, duplicates, comment blocks,
    mixed number formats, ragged whitespace): product reader == the reference's own reader
    == the oracle's restatement, entry for entry and densified."""
    _need(DUMP)
    rng = np.random.default_rng(2024)
    for case in range(40):
        n = int(rng.integers(1, 12))
        sym = bool(rng.integers(0, 2))
        nz = int(rng.integers(0, 3 * n + 1))
        lines = ["%%MatrixMarket matrix coordinate real " + ("symmetric" if sym else "general")]
        for _ in range(int(rng.integers(0, 4))):
            lines.append("% " + "x" * int(rng.integers(0, 30)))
        lines.append("%d %d %d" % (n, n, nz))
        for _ in range(nz):
            i, j = int(rng.integers(1, n + 1)), int(rng.integers(1, n + 1))
            if sym and j > i:
                i, j = j, i
            v = float(rng.standard_normal()) * 10.0 ** int(rng.integers(-3, 4))
            fmt = ["%d %d %.17g", "%d  %d   %e", " %d %d %g", "%d\t%d\t%.3f"][int(rng.integers(0, 4))]
            lines.append(fmt % (i, j, v))
        src = tmp_path / ("r%02d.mtx" % case)
        src.write_text("\n".join(lines) + "\n")
        assert _dump(DUMP, str(src), str(tmp_path / "p.bin")).returncode == 0
        assert _dump(DUMP, str(src), str(tmp_path / "pd.bin"), dense=True).returncode == 0
        dense = _read_dense(str(tmp_path / "pd.bin"))
        assert np.array_equal(dense, O.read_mtx_dense(str(src))), case
        if os.path.exists(REF_DUMP):
            assert _dump(REF_DUMP, str(src), str(tmp_path / "r.bin")).returncode == 0
            assert _dump(REF_DUMP, str(src), str(tmp_path / "rd.bin"), dense=True).returncode == 0
            assert open(str(tmp_path / "p.bin"), "rb").read() == open(str(tmp_path / "r.bin"), "rb").read(), case
            assert np.array_equal(dense, _read_dense(str(tmp_path / "rd.bin"))), case
