"""GPU parity tests for the BCSR variant, through the reference-named C entry points (sparse/bcsr.h:14-39).

Bars: b_row_start / b_col_idx bit-exact and b_values bit-exact against the reference's golden vectors (every golden
case has no empty block-row, where the reference's row pointers are well defined); Y of bcsr_sgemm_basic / _avx /
_avx2 bit-exact against the oracle (ascending-k accumulation from the bias, one rounding per term, ternary blocks =>
x*val exact); bcsr_sgemm_prelu_* equal PReLU of that, i.e. the north-star math, not the reference's literal loop."""
import numpy as np
import pytest

import __graft_entry__ as ge
from tests.golden.make_golden import BCSR_CASES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def t():
    import torch
    assert torch.cuda.is_available()
    mod = ge.load()
    mod.lib()
    return mod


@pytest.mark.parametrize("case", BCSR_CASES, ids=[c[0] for c in BCSR_CASES])
def test_bcsr_golden(t, port, golden, case):
    name, M, K, N, r, c, num, den, seed = case
    Wd = port.gen_ternary(K, N, seed, num, den)
    w = t.bcsr_from_dense(Wd, r, c)
    try:
        assert (w.r, w.c, w.br, w.bc, w.k) == (r, c, K // r, N // c, int(golden[f"bcsr.{name}.k"]))
        assert np.array_equal(w.b_row_start, golden[f"bcsr.{name}.row_start"])
        assert np.array_equal(w.b_col_idx, golden[f"bcsr.{name}.col_idx"])
        assert np.array_equal(w.b_values, golden[f"bcsr.{name}.values"])
        Xi, B2 = port.gen_intvalued((M, K), seed + 1, 512), np.full(N, 2.0, np.float32)
        assert np.array_equal(t.bcsr_sgemm_basic(Xi, w, B2, N), golden[f"bcsr.{name}.int.Y_bias"])
        Xu, Bu = port.gen_uniform((M, K), seed + 1), port.gen_uniform((N,), seed + 2)
        y = t.bcsr_sgemm_basic(Xu, w, Bu, N)
        assert np.array_equal(y, golden[f"bcsr.{name}.real.basic"])
        assert np.array_equal(t.bcsr_sgemm_avx(Xu, w, Bu, N), y) and np.array_equal(t.bcsr_sgemm_avx2(Xu, w, Bu, N), y)
        yp = t.bcsr_sgemm_prelu_basic(Xu, w, Bu, 0.2, N)
        assert np.array_equal(yp, np.where(y < 0, np.float32(0.2) * y, y))
        assert np.array_equal(t.bcsr_sgemm_prelu_avx(Xu, w, Bu, 0.2, N), yp)
        # the reference's literal prelu loop returns something else: reported, not matched
        lit = golden[f"bcsr.{name}.real.prelu_basic_literal"]
        print(f"{name}: max |PReLU(X*W+b) - reference literal loop| = {np.abs(yp - lit).max():.3e}")
    finally:
        w.free()


@pytest.mark.parametrize("shape", [(64, 512, 2048, 1, 8, 1, 2, 11), (130, 96, 100, 2, 4, 1, 4, 12), (33, 64, 64, 8, 8, 1, 10, 13),
                                   (40, 60, 66, 3, 5, 1, 2, 14), (5, 128, 256, 1, 16, 1, 3, 15), (256, 1024, 512, 1, 8, 1, 10, 16)])
def test_bcsr_vs_oracle(t, port, shape):
    """includes non-divisible shapes (bcsr.c:24-25 drops the remainder), generic block widths and empty block-rows"""
    M, K, N, r, c, num, den, seed = shape
    Wd = port.gen_ternary(K, N, seed, num, den)
    w = t.bcsr_from_dense(Wd, r, c)
    wo = port.bcsr_from_dense(Wd, r, c)  # standard CSR pointers
    try:
        assert w.k == wo.k and np.array_equal(w.b_row_start, wo.b_row_start) and np.array_equal(w.b_col_idx, wo.b_col_idx)
        assert np.array_equal(w.b_values, wo.b_values)
        X, B = port.gen_uniform((M, K), seed + 1), port.gen_uniform((N,), seed + 2)
        y = t.bcsr_sgemm_basic(X, w, B, N)
        assert np.array_equal(y, port.bcsr_sgemm_basic(X, wo, B, N))
        assert np.array_equal(t.bcsr_sgemm_prelu_basic(X, w, B, 0.2, N), port.bcsr_sgemm_prelu_math(X, wo, B, 0.2, N))
    finally:
        w.free()


def test_bcsr_known_answer_and_quirk(t, golden):
    w = t.bcsr_from_dense(np.ascontiguousarray(golden["kat_test_c.dense"]), 2, 2)  # test/test.c:13
    assert w.b_row_start.tolist() == [0, 2, 3] and w.b_col_idx.tolist() == [0, 1, 1] and w.k == 3
    w.free()
    q = t.bcsr_from_dense(np.ascontiguousarray(golden["kat_quirk.dense"]), 1, 2)
    assert q.b_row_start.tolist() == [0, 1, 1, 1, 2]  # standard CSR; the reference leaves [0, 1, 2, ?, ?] (bcsr.c:114-117)
    assert np.array_equal(q.b_col_idx, golden["kat_quirk.col_idx"]) and np.array_equal(q.b_values, golden["kat_quirk.values"])
    q.free()


def test_bcsr_agrees_with_tcsc(t, port):
    M, K, N = 128, 512, 512
    Wd = port.gen_ternary(K, N, 21, 1, 2)
    X, B = port.gen_intvalued((M, K), 22, 512), np.full(N, 2.0, np.float32)
    wb, wt = t.bcsr_from_dense(Wd, 1, 8), t.tcsc_from_dense(Wd)
    assert np.array_equal(t.bcsr_sgemm_basic(X, wb, B, N), t.tcsc_sgemm_basic(X, wt, B))
    wb.free()
    wt.free()

