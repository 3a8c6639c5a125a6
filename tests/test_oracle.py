"""The oracle (oracle/tsg_oracle.c) pinned against outputs of the unmodified reference.

Every array in tests/golden/reference_vectors.npz was produced by the reference itself
(tests/golden/make_golden.py); where oracle/_ref is present the same checks also run live against it on fresh
seeds.  Integer/index results must be bit-exact.  fp32 results must be bit-exact too, because the oracle keeps the
reference's summation order and the reference is compiled without -ffast-math.
"""
import numpy as np
import pytest

from tests.golden.make_golden import BCSR_CASES, TCSC_CASES

ARR = ("csp", "csn", "rip", "rin")


def _same_tcsc(w, golden, prefix):
    for nm, arr in zip(ARR, w.arrays()):
        exp = golden[f"{prefix}.{nm}"]
        assert arr.dtype == np.int32 and np.array_equal(arr, exp), f"{prefix}.{nm}"


# ---- known-answer vectors --------------------------------------------------------------------
def test_kat_test_c_matrix(port, golden):
    """test/test.c:6-11 -- and the values quoted in SURVEY.md 8c, typed here independently of the npz."""
    w = port.tcsc_from_dense(golden["kat_test_c.dense"])
    _same_tcsc(w, golden, "kat_test_c.tcsc")
    assert w.n_elem_pos == 0 and w.n_elem_neg == 7
    assert w.col_start_pos.tolist() == [0, 0, 0, 0, 0]
    assert w.col_start_neg.tolist() == [0, 1, 3, 5, 7]
    assert w.row_index_neg.tolist() == [0, 0, 1, 2, 3, 0, 2]
    b = port.bcsr_from_dense(golden["kat_test_c.dense"], 2, 2)
    assert b.b_row_start.tolist() == [0, 2, 3] and b.b_col_idx.tolist() == [0, 1, 1] and b.k == 3
    assert np.array_equal(b.b_row_start, golden["kat_test_c.bcsr22.row_start"])
    assert np.array_equal(b.b_col_idx, golden["kat_test_c.bcsr22.col_idx"])
    assert np.array_equal(b.b_values, golden["kat_test_c.bcsr22.values"])


@pytest.mark.parametrize("name", ["odd", "special", "zeros", "allpos", "allneg"])
def test_kat_float_predicates(port, golden, name):
    """tcsc.c:14-17: only ==1.0f / ==-1.0f count; 0.5, 2, NaN, inf, -0.0, 1-ulp are dropped."""
    _same_tcsc(port.tcsc_from_dense(golden[f"kat_{name}.dense"]), golden, f"kat_{name}.tcsc")


def test_kat_odd_values_spelled_out(port, golden):
    w = port.tcsc_from_dense(golden["kat_odd.dense"])
    assert (w.n_elem_pos, w.n_elem_neg) == (5, 3)
    assert w.col_start_pos.tolist() == [0, 2, 4, 5] and w.col_start_neg.tolist() == [0, 1, 2, 3]
    assert w.row_index_pos.tolist() == [0, 3, 1, 3, 2] and w.row_index_neg.tolist() == [1, 2, 0]


def test_kat_int_predicates(port, golden):
    """SparseGEMM.h:26-33: >=1 / <=-1, so +-2, 5, -7 are non-zeros."""
    _same_tcsc(port.sparse_format(golden["kat_oddi.dense"]), golden, "kat_oddi.tcsc")


def test_kat_bcsr_quirk(port, golden):
    """bcsr.c:114-117: the reference only appends a row pointer for non-empty block-rows."""
    q = port.bcsr_from_dense(golden["kat_quirk.dense"], 1, 2, quirk=True, tail_fill=-7)
    assert np.array_equal(q.b_row_start[:3], golden["kat_quirk.row_start_defined_prefix"])
    assert q.b_row_start[3:].tolist() == [-7, -7]
    assert np.array_equal(q.b_col_idx, golden["kat_quirk.col_idx"]) and q.k == int(golden["kat_quirk.k"])
    assert np.array_equal(q.b_values, golden["kat_quirk.values"])
    std = port.bcsr_from_dense(golden["kat_quirk.dense"], 1, 2)
    assert std.b_row_start.tolist() == [0, 1, 1, 1, 2]


def test_generators_frozen(port, golden):
    assert np.array_equal(port.gen_ternary(16, 16, 42, 1, 10), golden["gen.ternary_1_10.seed42"])
    assert np.array_equal(port.gen_ternary(8, 8, 7, 1, 2, np.int32), golden["gen.ternary_i32_1_2.seed7"])
    assert np.array_equal(port.gen_uniform((4, 16), 43), golden["gen.uniform.seed43"])
    assert np.array_equal(port.gen_intvalued((4, 16), 43, 512), golden["gen.intvalued.seed43"])


def test_generator_statistics(port):
    W = port.gen_ternary(512, 512, 42, 1, 10)
    assert set(np.unique(W)) == {-1.0, 0.0, 1.0}
    assert abs((W == 1).mean() - 0.05) < 0.003 and abs((W == -1).mean() - 0.05) < 0.003
    X = port.gen_uniform((256, 256), 43)
    assert X.min() >= -1.0 and X.max() < 1.0 and abs(X.mean()) < 0.01
    Xi = port.gen_intvalued((128, 128), 5, 512)
    assert Xi.min() >= -512 and Xi.max() <= 512 and np.array_equal(Xi, np.round(Xi))


# ---- seeded cases: index arrays and all kernel variants ------------------------------------------------------------
@pytest.mark.parametrize("case", TCSC_CASES, ids=[c[0] for c in TCSC_CASES])
def test_tcsc_vs_reference_golden(port, golden, case):
    name, M, K, N, num, den, seed = case
    Wd = port.gen_ternary(K, N, seed, num, den)
    w = port.tcsc_from_dense(Wd)
    _same_tcsc(w, golden, f"tcsc.{name}")
    _same_tcsc(port.sparse_format(port.gen_ternary(K, N, seed, num, den, np.int32)), golden, f"tcsc.{name}")
    Xi, B2 = port.gen_intvalued((M, K), seed + 1, 512), np.full(N, 2.0, np.float32)
    for y in (port.tcsc_sgemm_basic(Xi, w, B2), port.tcsc_sgemm_optimized(Xi, w, B2), port.sparse_gemm(Xi, w, B2)):
        assert np.array_equal(y, golden[f"tcsc.{name}.int.Y_bias"])
    for y in (port.tcsc_sgemm_prelu_basic(Xi, w, B2, 0.25), port.tcsc_sgemm_prelu_optimized_separate(Xi, w, B2, 0.25),
              port.tcsc_sgemm_prelu_optimized_onthego(Xi, w, B2, 0.25), port.sparse_gemm_prelu(Xi, w, B2, 0.25)):
        assert np.array_equal(y, golden[f"tcsc.{name}.int.Y_prelu"])
    Xu, Bu = port.gen_uniform((M, K), seed + 1), port.gen_uniform((N,), seed + 2)
    pairs = {
        "basic": port.tcsc_sgemm_basic(Xu, w, Bu),
        "optimized": port.tcsc_sgemm_optimized(Xu, w, Bu),
        "prelu_basic": port.tcsc_sgemm_prelu_basic(Xu, w, Bu, 0.2),
        "prelu_separate": port.tcsc_sgemm_prelu_optimized_separate(Xu, w, Bu, 0.2),
        "prelu_onthego": port.tcsc_sgemm_prelu_optimized_onthego(Xu, w, Bu, 0.2),
        "sparse_gemm": port.sparse_gemm(Xu, w, Bu),
        "sparse_gemm_prelu": port.sparse_gemm_prelu(Xu, w, Bu, 0.2),
        "gemm_basic": port.gemm_basic(Xu, Wd, Bu),
    }
    for key, y in pairs.items():
        assert np.array_equal(y, golden[f"tcsc.{name}.real.{key}"]), key  # bit-exact: same summation order
    # the fp64 anchor bounds both; the -ffast-math build of the reference differs in the last bits only
    y64 = port.tcsc_sgemm_f64(Xu, w, Bu, 0.2)
    scale = np.maximum(np.abs(y64), 1.0)
    assert np.max(np.abs(pairs["prelu_basic"] - y64) / scale) <= 1e-5
    assert np.max(np.abs(golden[f"tcsc.{name}.real.prelu_basic_fastmath"] - y64) / scale) <= 1e-5


@pytest.mark.parametrize("case", BCSR_CASES, ids=[c[0] for c in BCSR_CASES])
def test_bcsr_vs_reference_golden(port, golden, case):
    name, M, K, N, r, c, num, den, seed = case
    Wd = port.gen_ternary(K, N, seed, num, den)
    b = port.bcsr_from_dense(Wd, r, c)
    assert b.k == int(golden[f"bcsr.{name}.k"])
    assert np.array_equal(b.b_row_start, golden[f"bcsr.{name}.row_start"])
    assert np.array_equal(b.b_col_idx, golden[f"bcsr.{name}.col_idx"])
    assert np.array_equal(b.b_values, golden[f"bcsr.{name}.values"])
    bq = port.bcsr_from_dense(Wd, r, c, quirk=True)
    assert np.array_equal(bq.b_row_start, b.b_row_start)  # no empty block-row => quirk is invisible
    Xi, B2 = port.gen_intvalued((M, K), seed + 1, 512), np.full(N, 2.0, np.float32)
    assert np.array_equal(port.bcsr_sgemm_basic(Xi, b, B2, N), golden[f"bcsr.{name}.int.Y_bias"])
    Xu, Bu = port.gen_uniform((M, K), seed + 1), port.gen_uniform((N,), seed + 2)
    yb = port.bcsr_sgemm_basic(Xu, b, Bu, N)
    assert np.array_equal(yb, golden[f"bcsr.{name}.real.basic"])
    assert np.array_equal(port.bcsr_sgemm_prelu_literal(Xu, b, Bu, 0.2, N), golden[f"bcsr.{name}.real.prelu_basic_literal"])
    if c == 8:
        assert np.array_equal(yb, golden[f"bcsr.{name}.real.avx"])
        assert np.array_equal(port.bcsr_sgemm_prelu_literal(Xu, b, Bu, 0.2, N), golden[f"bcsr.{name}.real.prelu_avx_literal"])
    if c == 8 and r == 8:
        assert np.array_equal(yb, golden[f"bcsr.{name}.real.avx2"])
    # BCSR and dense agree (test_bcsr.cpp:36, abs tol 1e-4)
    assert port.compare(yb, port.gemm_basic(Xu, Wd, Bu), 1e-4)


def test_bcsr_prelu_literal_is_not_prelu(port):
    """the reference's bcsr_sgemm_prelu_* activates every partial sum; document the gap."""
    Wd = port.gen_ternary(64, 64, 9, 1, 2)
    b = port.bcsr_from_dense(Wd, 1, 8)
    X, B = port.gen_uniform((4, 64), 10), port.gen_uniform((64,), 11)
    lit = port.bcsr_sgemm_prelu_literal(X, b, B, 0.2, 64)
    math = port.bcsr_sgemm_prelu_math(X, b, B, 0.2, 64)
    assert np.max(np.abs(lit - math)) > 1e-2
    w = port.tcsc_from_dense(Wd)
    assert port.compare(math, port.tcsc_sgemm_prelu_basic(X, w, B, 0.2), 1e-4)


# ---- live cross-check against oracle/_ref on fresh seeds -----------------------------------------------------------
@pytest.mark.parametrize("shape", [(16, 200, 150, 1, 4, 91), (64, 512, 512, 1, 2, 92), (7, 1000, 33, 1, 20, 93)])
def test_live_reference_crosscheck(port, ref, shape):
    M, K, N, num, den, seed = shape
    Wd = port.gen_ternary(K, N, seed, num, den)
    w, wr = port.tcsc_from_dense(Wd), ref.tcsc_from_dense(Wd)
    for a, b in zip(w.arrays(), wr.arrays()):
        assert np.array_equal(a, b)
    X, B = port.gen_uniform((M, K), seed + 1), port.gen_uniform((N,), seed + 2)
    for fn in ("tcsc_sgemm_basic", "tcsc_sgemm_optimized"):
        assert np.array_equal(getattr(port, fn)(X, w, B), getattr(ref, fn)(X, wr, B)), fn
    for fn in ("tcsc_sgemm_prelu_basic", "tcsc_sgemm_prelu_optimized_separate", "tcsc_sgemm_prelu_optimized_onthego"):
        assert np.array_equal(getattr(port, fn)(X, w, B, 0.2), getattr(ref, fn)(X, wr, B, 0.2)), fn
    assert ref.compare(port.tcsc_sgemm_basic(X, w, B), ref.gemm_basic(X, Wd, B))  # main.cpp:317


def test_openmp_build_of_the_reference_gives_the_same_bits(port, ref):
    """bench.py --impl reference times sparseGEMM_PReLU<float> from the -fopenmp build (SparseGEMM.h:151-168 parallelises
    over m, so each output keeps its summation order): same bits as the single-threaded build and as the oracle port"""
    from oracle.pyoracle import Ref, ref_available
    if not ref_available("omp"):
        pytest.skip("oracle/_ref/libref_oracle_omp.so not built (no OpenMP-capable g++)")
    omp = Ref("omp")
    assert omp.omp_max_threads() >= 1 and "-fopenmp" in omp.build_flags
    Wd = port.gen_ternary(256, 192, 5, 1, 4)
    X, b = port.gen_uniform((67, 256), 6), port.gen_uniform((192,), 7)
    W = port.tcsc_from_dense(Wd)
    want = ref.sparse_gemm_prelu(X, W, b, 0.2)
    assert np.array_equal(omp.sparse_gemm_prelu(X, W, b, 0.2), want)
    assert np.array_equal(port.sparse_gemm_prelu(X, W, b, 0.2), want)
    assert omp.time_sparse_gemm_prelu(X, W, b, 0.2, reps=1) > 0
