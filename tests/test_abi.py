"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol the headers
declare (plus the C++-mangled names objects compiled against the REFERENCE's headers look for), fails loudly instead
of falling back to the CPU, and the public headers compile as C11 and C++17."""
import ctypes
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import __graft_entry__ as ge

ROOT = ge.ROOT
INC = os.path.join(ROOT, "include")


@pytest.fixture(scope="module")
def t():
    if not os.path.exists(os.path.join(ge.PKG_DIR, "libtsgemm_b200.so")):
        ge.build()
    mod = ge.load()
    mod.lib()
    return mod


def _declared_functions(path):
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    names = set()
    for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", src):
        name = m.group(1)
        if name in ("defined", "sizeof", "static_assert") or name.startswith("__"):
            continue
        names.add(name)
    return names


@pytest.mark.parametrize("header", ["sparse/tcsc.h", "sparse/bcsr.h", "tsgemm_b200.h"])
def test_every_declared_symbol_is_exported(t, header):
    L = t.lib()
    names = _declared_functions(os.path.join(INC, header))
    assert len(names) >= 7
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, f"{header}: not exported: {missing}"


def test_reference_names_present(t):
    """The exact entry points of the reference: sparse/tcsc.h:19-48, sparse/bcsr.h:14-39."""
    L = t.lib()
    for n in ("tcsc_from_dense tcsc_sgemm_basic tcsc_sgemm_optimized tcsc_sgemm_prelu_basic tcsc_sgemm_prelu_optimized_separate "
              "tcsc_sgemm_prelu_optimized_onthego tcsc_free bcsr_from_dense bcsr_sgemm_basic bcsr_sgemm_prelu_basic bcsr_sgemm_avx "
              "bcsr_sgemm_prelu_avx bcsr_sgemm_avx2 tsg_sparse_gemm_f32 tsg_sparse_format_build_i32").split():
        assert hasattr(L, n), n


def test_cxx_mangled_names_exported(t):
    """Objects compiled against the reference's own headers (no extern "C") look for these."""
    out = subprocess.run(["nm", "-D", "--defined-only", t.LIB_PATH], capture_output=True, text=True, check=True).stdout
    for sym in ("_Z15tcsc_from_densePfii", "_Z22tcsc_sgemm_prelu_basicPfPK6tcsc_tS_fS_iii", "_Z16tcsc_sgemm_basicPfPK6tcsc_tS_S_iii",
                "_Z20tcsc_sgemm_optimizedPfPK6tcsc_tS_S_iii", "_Z35tcsc_sgemm_prelu_optimized_separatePfPK6tcsc_tS_fS_iii",
                "_Z34tcsc_sgemm_prelu_optimized_onthegoPfPK6tcsc_tS_fS_iii", "_Z9tcsc_freeP6tcsc_t", "_Z15bcsr_from_densePfiiii",
                "_Z16bcsr_sgemm_basicPf6bcsr_tS_S_iii", "_Z22bcsr_sgemm_prelu_basicPf6bcsr_tS_fS_iii", "_Z14bcsr_sgemm_avxPf6bcsr_tS_S_iii",
                "_Z20bcsr_sgemm_prelu_avxPf6bcsr_tS_fS_iii", "_Z15bcsr_sgemm_avx2Pf6bcsr_tS_S_iii"):
        assert re.search(rf"\b{re.escape(sym)}\b", out), sym


def test_headers_compile_as_c11_and_cxx17():
    c_src = '#include "sparse/tcsc.h"\n#include "sparse/bcsr.h"\n#include "common.h"\n#include "tsgemm_b200.h"\n' \
            "int main(void){ gemm_func f = (gemm_func)tcsc_sgemm_basic; prelu_func g = (prelu_func)tcsc_sgemm_prelu_basic; return f==0 && g==0; }\n"
    cxx_src = c_src.replace('#include "tsgemm_b200.h"\n', '#include "tsgemm_b200.h"\n#include "SparseGEMM.h"\n')
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "a.c"), "w").write(c_src)
        open(os.path.join(d, "a.cpp"), "w").write(cxx_src)
        subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", INC, "-c", os.path.join(d, "a.c"), "-o", os.path.join(d, "a.o")], check=True)
        subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", INC, "-c", os.path.join(d, "a.cpp"), "-o", os.path.join(d, "b.o")], check=True)


def test_no_cpu_fallback(t):
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(t.TsgError, match="no CUDA device|CUDA"):
        t.tcsc_from_dense(np.eye(8, dtype=np.float32))
    assert t.lib().tsg_device_check() != 0
    assert "fallback" in t.last_error()


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the package or include/ may import, link or name it."""
    bad = []
    for base in (ge.PKG_DIR, INC):
        for dirpath, _, files in os.walk(base):
            for f in files:
                if f.endswith((".so", ".o", ".pyc")):
                    continue
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"pyoracle|liboracle|libref_oracle|orc_[a-z]+_|from oracle|import oracle", txt):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
    out = subprocess.run(["ldd", os.path.join(ge.PKG_DIR, "libtsgemm_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_partition_is_a_partition(t):
    for N in (1, 31, 32, 100, 512, 4096, 14336, 16384, 1000):
        for world in (1, 2, 3, 4, 8):
            nxt = 0
            for r in range(world):
                c0, nc = t.partition(N, r, world)
                assert c0 == nxt and nc >= 0
                if r < world - 1:
                    assert nc % 32 == 0
                nxt = c0 + nc
            assert nxt == N


def test_reference_programs_link_against_the_library():
    """oracle/Makefile `drivers`: the reference's unmodified main.cpp / test_bcsr.cpp objects (compiled against the
    reference's own headers) and its SparseGEMM.cpp (compiled against include/SparseGEMM.h) link against the library.
    Built only where /root/reference exists; running them needs a B200 (tests/test_gpu_dropin.py)."""
    if not os.path.isdir("/root/reference/sparse"):
        pytest.skip("reference tree not present")
    ge.build()
    for exe in ("ref_main_on_b200", "ref_test_bcsr_on_b200", "ref_sparsegemm_on_b200"):
        path = os.path.join(ROOT, "oracle", "_ref", exe)
        assert os.path.exists(path), exe
        needed = subprocess.run(["ldd", path], capture_output=True, text=True).stdout
        assert "libtsgemm_b200.so" in needed, exe
        undefined = subprocess.run(["nm", "-u", path], capture_output=True, text=True).stdout
        if exe == "ref_main_on_b200":
            assert "_Z22tcsc_sgemm_prelu_basicPfPK6tcsc_tS_fS_iii" in undefined  # resolved by the library at load time
        if exe == "ref_sparsegemm_on_b200":
            assert "tsg_sparse_gemm_f32" in undefined
