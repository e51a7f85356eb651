"""CPU tests of the drop-in boundary: libcgb200.so loads, exports every symbol that
include/cgb200.h declares, host-only logic works, and compute entry points FAIL LOUDLY
without a GPU (there is no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "cgb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cgb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(cgb):
    lib = cgb.load()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/cgb200.h but not exported"
    bound = {s[0] for s in cgb.SIGNATURES}
    assert set(declared) == bound, set(declared) ^ bound


def test_abi_version_and_variants(cgb):
    lib = cgb.load()
    assert lib.cgb_abi_version() == 1
    names = cgb.gemv_variants()
    assert len(names) >= 4 and len(set(names)) == len(names)
    assert lib.cgb_gemv_variant_name(-1) is None and lib.cgb_gemv_variant_name(999) is None


def test_partition_is_the_reference_rule(cgb, O):
    for n, p in [(10, 1), (10, 3), (40000, 8), (56569, 8), (7, 7), (1 << 33, 5)]:
        assert cgb.partition(n, p) == O.partition(n, p)
    with pytest.raises(cgb.CgbError):
        cgb.partition(10, 0)


@pytest.mark.parametrize("n", [1, 2, 100, 4096, 20000, 56568])
def test_init_source_term_bitwise_vs_oracle(cgb, O, n):
    """cgb_init_source_term (the one entry point that needs no GPU: a host libm loop like the
    reference's, cg.cc:218-234) == the oracle's restatement, bit for bit, incl. h != 1/n."""
    assert np.array_equal(cgb.init_source_term(n), O.init_source_term(n))
    assert np.array_equal(cgb.init_source_term(n, 0.37 / n), O.init_source_term(n, 0.37 / n))
    b = cgb.init_source_term(n)
    assert b[0] == 0.0 and np.all(b <= 0.0)


def test_no_cpu_fallback(cgb):
    """Without a CUDA device every compute entry point returns an error; nothing is computed
    on the host."""
    try:
        ndev = cgb.device_count()
    except cgb.CgbError as e:
        ndev = 0
        assert e.code == 3  # CGB_ERR_NO_DEVICE
    if ndev > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(cgb.CgbError) as ei:
        cgb.Context(128)
    assert ei.value.code == 3
    lib = cgb.load()
    assert lib.cgb_solve(None, None, 1, 1e-10, None, None) != 0
    assert b"null" in lib.cgb_last_error()


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "conjugate-gradient_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".cc")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                for needle in ("cg_oracle", "import oracle", "oracle/", "/root/reference"):
                    hits = [ln for ln in text.splitlines() if needle in ln
                            and not ln.strip().startswith(("//", "#", "*", "/*"))]
                    assert not hits, (f, needle, hits[:2])
    for f in ("include/cgb200.h",):
        assert "cg_oracle" not in open(os.path.join(ROOT, f)).read()


def test_missing_library_fails_loudly(cgb, monkeypatch):
    """Without libcgb200.so the binding raises (pointing at the build step) instead of
    computing anything another way."""
    import importlib
    capi = importlib.import_module("conjugate-gradient_b200._capi")
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", "/nonexistent/libcgb200.so")
    with pytest.raises(cgb.CgbError) as ei:
        capi.load()
    assert "no CPU fallback" in str(ei.value)
    with pytest.raises(cgb.CgbError):
        capi.partition(10, 2)


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/cgb200.h compiles as strict C99 and the library links from a plain C program --
    the boundary a cgo / JNI / ctypes binding would use (no C++ or torch types)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    lib_dir = os.path.join(ROOT, "conjugate-gradient_b200")
    if not gcc or not os.path.exists(os.path.join(lib_dir, "libcgb200.so")):
        pytest.skip("gcc or libcgb200.so missing")
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include "cgb200.h"
#include <stdio.h>
int main(void) {
    int64_t s[3], c[3];
    double b[4];
    cgb_ctx *ctx = NULL;
    int ndev = -1, rc;
    if (cgb_partition(10, 3, s, c) != CGB_OK) return 1;
    if (cgb_init_source_term(4, 0.25, b) != CGB_OK || b[0] != 0.0) return 2;
    rc = cgb_device_count(&ndev);
    if (rc != CGB_OK && cgb_create(16, 0, 1, 0, &ctx) != CGB_ERR_NO_DEVICE) return 3;
    printf("%d %lld %lld %lld\n", cgb_abi_version(), (long long)c[0], (long long)c[1], (long long)c[2]);
    return 0;
}
''')
    exe = tmp_path / "abi"
    subprocess.check_call([gcc, "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-L", lib_dir, "-lcgb200", "-Wl,-rpath," + lib_dir, "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.split() == ["1", "3", "3", "4"]
