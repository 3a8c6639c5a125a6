"""generateSparseMatrix patterns (reference SparseGEMM.h:53-102) as counter-based generators.

CPU part: (1) the oracle's restatement has the structure of the reference's two patterns; (2) the product's per-element
arithmetic (csrc/gen_pattern.h, the header its CUDA kernels are built from) is compiled with g++ and must reproduce the
oracle bit for bit -- the two were written independently (thresholds by bisection vs. a sort of the row's keys).
GPU part: the device generators against the oracle."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "sparse-matrix-multiplication-benchmark_b200", "csrc")

SHAPES = [(8, 64, 4), (16, 200, 10), (5, 37, 3), (64, 512, 2), (3, 1000, 100), (7, 24, 20), (4, 96, 1)]

HARNESS = r"""
#include "gen_pattern.h"
extern "C" void run_window(int *out, int H, int W, int nz, unsigned long long seed) {
    for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w) out[(long long)h * W + w] = gp_window_value(seed, h, w, W, nz);
}
extern "C" void run_skewed(int *out, int H, int W, int nz, unsigned long long seed) {
    for (int h = 0; h < H; ++h) {
        int np, nm;
        gp_row_limits(seed, h, W, nz, &np, &nm);
        const uint64_t tp = np > 0 ? gp_kth_key_serial(seed, h, W, np) : 0, ta = np + nm > 0 ? gp_kth_key_serial(seed, h, W, np + nm) : 0;
        for (int w = 0; w < W; ++w) out[(long long)h * W + w] = gp_skew_value(gp_key(seed, h, w, W), np, nm, tp, ta);
    }
}
"""


@pytest.fixture(scope="module")
def header_lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("gp")
    src = d / "harness.cpp"
    src.write_text(HARNESS)
    so = d / "libgp.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", CSRC, str(src), "-o", str(so)], check=True)
    L = C.CDLL(str(so))
    ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    for f in (L.run_window, L.run_skewed):
        f.argtypes = [ip, C.c_int, C.c_int, C.c_int, C.c_ulonglong]
    return L


@pytest.mark.parametrize("H,W,nz", [s for s in SHAPES if s[2] >= 2])
def test_window_pattern_structure(port, H, W, nz):
    m = port.gen_sparse_pattern(H, W, nz, True, 7)
    assert set(np.unique(m)) <= {-1, 0, 1}
    win = 2 * nz
    for h in range(H):
        for w0 in range(0, W, win):
            seg = m[h, w0:w0 + win]
            assert not seg[1::2].any(), "only even offsets are written (SparseGEMM.h:60-61)"
            if w0 + win <= W:  # a complete window holds exactly one +1 and one -1 (SparseGEMM.h:62-66)
                assert (seg == 1).sum() == 1 and (seg == -1).sum() == 1
            else:
                assert (seg == 1).sum() <= 1 and (seg == -1).sum() <= 1
    assert np.array_equal(m, port.gen_sparse_pattern(H, W, nz, True, 7)) and not np.array_equal(m, port.gen_sparse_pattern(H, W, nz, True, 8))


@pytest.mark.parametrize("H,W,nz", SHAPES)
def test_skewed_pattern_structure(port, H, W, nz):
    m = port.gen_sparse_pattern(H, W, nz, False, 11)
    per_row, half, dmax = W // nz, (W // nz) // 2, (W // nz) // 20 + 1
    for h in range(H):
        p, n = int((m[h] == 1).sum()), int((m[h] == -1).sum())
        d = p - half
        assert 0 <= d <= dmax, "limitPos = half + d, d in [0, W/nonZero/20+1] (SparseGEMM.h:73-76)"
        assert n == max(half - d, 0), "limitNeg = half - d (SparseGEMM.h:77)"
    if H >= 16 and dmax >= 2:
        skews = (m == 1).sum(1) - (m == -1).sum(1)
        assert len(set(skews.tolist())) > 1, "the +/- skew varies from row to row"
    assert np.array_equal(m, port.gen_sparse_pattern(H, W, nz, False, 11))


def test_pattern_argument_checks(port):
    with pytest.raises(ValueError):
        port.gen_sparse_pattern(4, 16, 1, True, 1)  # the reference never terminates for nonZero = 1 (SparseGEMM.h:62-64)
    with pytest.raises(ValueError):
        port.gen_sparse_pattern(4, 16, 0, False, 1)
    assert port.gen_sparse_pattern(0, 16, 4, False, 1).shape == (0, 16)


@pytest.mark.parametrize("H,W,nz", SHAPES)
def test_product_header_arithmetic_matches_oracle(port, header_lib, H, W, nz):
    out = np.empty((H, W), np.int32)
    if nz >= 2:
        header_lib.run_window(out, H, W, nz, 7)
        assert np.array_equal(out, port.gen_sparse_pattern(H, W, nz, True, 7))
    header_lib.run_skewed(out, H, W, nz, 11)
    assert np.array_equal(out, port.gen_sparse_pattern(H, W, nz, False, 11))


def test_patterns_feed_the_tcsc_builder(port):
    """the patterns are inputs of SparseFormat (SparseGEMM.cpp:96-101): the oracle builder accepts them and sees the
    same counts"""
    m = port.gen_sparse_pattern(128, 256, 8, False, 3)
    w = port.tcsc_from_dense(m.astype(np.float32))
    assert w.n_elem_pos == int((m == 1).sum()) and w.n_elem_neg == int((m == -1).sum())


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,nz", SHAPES + [(512, 4096, 10)])
def test_device_generators_match_oracle(port, H, W, nz):
    import torch

    import __graft_entry__ as ge
    t = ge.load()
    if nz >= 2:
        got = t.gen_sparse_pattern(H, W, nz, True, 7).cpu().numpy()
        assert np.array_equal(got, port.gen_sparse_pattern(H, W, nz, True, 7))
    got = t.gen_sparse_pattern(H, W, nz, False, 11).cpu().numpy()
    assert np.array_equal(got, port.gen_sparse_pattern(H, W, nz, False, 11))
    gf = t.gen_sparse_pattern(H, W, nz, False, 11, dtype=torch.float32).cpu().numpy()
    assert np.array_equal(gf, got.astype(np.float32))
