"""Binary-level drop-in: the reference's OWN test program (test/test_bcsr.cpp, unmodified, compiled against the
reference's own headers) linked against libtsgemm_b200.so instead of sparse/bcsr.c (oracle/Makefile `drivers`).
It builds a BCSR matrix with bcsr_from_dense, multiplies with bcsr_sgemm_basic(X, *W_bcsr, ...) -- struct by value --
compares against its own CPU gemm_basic with its own compare() (abs 1e-4) and free()s the arrays itself."""
import os
import subprocess

import pytest

import __graft_entry__ as ge

pytestmark = pytest.mark.gpu


def test_reference_test_bcsr_program_passes_on_the_product():
    exe = os.path.join(ge.ROOT, "oracle", "_ref", "ref_test_bcsr_on_b200")
    if not os.path.exists(exe):
        pytest.skip("driver not built (needs /root/reference at build time)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "Test passed! Results match." in out.stdout, out.stdout  # test/test_bcsr.cpp:37
