"""GPU parity tests for the TCSC path, through the C-ABI (reference-named entry points and the device-level API).

Bars (BASELINE.json north_star):
  * TCSC index arrays: bit-exact against the reference's golden vectors and against the oracle;
  * Y, tiled kernel (M >= 32): BIT-EXACT for every entry point, for real-valued X too, because each entry point
    reproduces its reference function's summation order;
  * Y, skinny kernel (M < 32): bit-exact on integer-valued X; max |y - y64| / max(|y64|, 1) <= 1e-5 otherwise, with the
    reference's own error printed alongside.
"""
import numpy as np
import pytest

import __graft_entry__ as ge
from tests.golden.make_golden import TCSC_CASES

pytestmark = pytest.mark.gpu

TOL = 1e-5  # north_star: max relative error <= 1e-5 versus a double-precision accumulation (denominator max(|y64|,1))


@pytest.fixture(scope="module")
def t():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    mod = ge.load()
    mod.lib()
    assert mod.lib().tsg_device_check() == 0, mod.last_error()
    return mod


def rel_err(y, y64):
    return float(np.max(np.abs(y.astype(np.float64) - y64) / np.maximum(np.abs(y64), 1.0))) if y.size else 0.0


def same_arrays(w, exp_arrays):
    for a, b in zip(w.arrays(), exp_arrays):
        assert a.dtype == np.int32 and np.array_equal(a, b)


# ---- conversion ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["test_c", "odd", "special", "zeros", "allpos", "allneg"])
def test_convert_known_answers(t, golden, name):
    d = golden[f"kat_{name}.dense"]
    w = t.tcsc_from_dense(np.ascontiguousarray(d))
    same_arrays(w, [golden[f"kat_{name}.tcsc.{k}"] for k in ("csp", "csn", "rip", "rin")])
    w.free()


def test_convert_int_predicates(t, golden):
    sf = t.SparseFormat(np.ascontiguousarray(golden["kat_oddi.dense"]))
    same_arrays(sf, [golden[f"kat_oddi.tcsc.{k}"] for k in ("csp", "csn", "rip", "rin")])


@pytest.mark.parametrize("case", TCSC_CASES, ids=[c[0] for c in TCSC_CASES])
def test_convert_golden(t, port, golden, case):
    name, M, K, N, num, den, seed = case
    w = t.tcsc_from_dense(port.gen_ternary(K, N, seed, num, den))
    exp = [golden[f"tcsc.{name}.{k}"] for k in ("csp", "csn", "rip", "rin")]
    same_arrays(w, exp)
    sf = t.SparseFormat(port.gen_ternary(K, N, seed, num, den, np.int32))
    same_arrays(sf, exp)
    w.free()


@pytest.mark.parametrize("shape", [(4096, 4096, 1, 10, 42), (1000, 3000, 1, 2, 7), (33, 65, 1, 3, 8), (5000, 17, 1, 100, 9), (1, 1, 1, 1, 3)])
def test_convert_vs_oracle_and_device_generator(t, port, shape):
    import torch
    K, N, num, den, seed = shape
    Wd_dev = t.gen_ternary(K, N, seed, num, den)
    Wd = port.gen_ternary(K, N, seed, num, den)
    assert np.array_equal(Wd_dev.cpu().numpy(), Wd), "device generator differs from the oracle generator"
    exp = port.tcsc_from_dense(Wd)
    w = t.DeviceTcsc.from_dense(Wd_dev)
    for a, b in zip(w.download(), exp.arrays()):
        assert np.array_equal(a, b)
    wi = t.DeviceTcsc.from_dense(t.gen_ternary(K, N, seed, num, den, dtype=torch.int32))
    for a, b in zip(wi.download(), exp.arrays()):
        assert np.array_equal(a, b)


def test_generators_match_oracle(t, port):
    assert np.array_equal(t.gen_uniform((37, 129), 43).cpu().numpy(), port.gen_uniform((37, 129), 43))
    assert np.array_equal(t.gen_intvalued((37, 129), 5, 512).cpu().numpy(), port.gen_intvalued((37, 129), 5, 512))
    full = port.gen_ternary(64, 96, 11, 1, 3)
    assert np.array_equal(t.gen_ternary_slice(64, 96, 32, 40, 11, 1, 3).cpu().numpy(), full[:, 32:72])


# ---- GEMM: golden vectors (M small => skinny kernel by default; the tiled kernel is forced as well) -----------------------
VARIANTS = [("tcsc_sgemm_basic", None, "basic"), ("tcsc_sgemm_optimized", None, "optimized"), ("tcsc_sgemm_prelu_basic", 0.2, "prelu_basic"),
            ("tcsc_sgemm_prelu_optimized_separate", 0.2, "prelu_separate"), ("tcsc_sgemm_prelu_optimized_onthego", 0.2, "prelu_onthego")]


def call(t, fn, X, w, B, a):
    return getattr(t, fn)(X, w, B) if a is None else getattr(t, fn)(X, w, B, a)


@pytest.mark.parametrize("force", [0, 1], ids=["auto(skinny)", "tiled"])
@pytest.mark.parametrize("case", TCSC_CASES, ids=[c[0] for c in TCSC_CASES])
def test_gemm_golden(t, port, golden, case, force):
    name, M, K, N, num, den, seed = case
    Wd = port.gen_ternary(K, N, seed, num, den)
    w = t.tcsc_from_dense(Wd)
    wo = port.tcsc_from_dense(Wd)
    t.lib().tsg_tcsc_set_kernel(force)
    try:
        Xi, B2 = port.gen_intvalued((M, K), seed + 1, 512), np.full(N, 2.0, np.float32)
        for fn, a, _ in VARIANTS:
            a_i = None if a is None else 0.25
            key = "Y_bias" if a is None else "Y_prelu"
            assert np.array_equal(call(t, fn, Xi, w, B2, a_i), golden[f"tcsc.{name}.int.{key}"]), fn  # exact, any order
        Xu, Bu = port.gen_uniform((M, K), seed + 1), port.gen_uniform((N,), seed + 2)
        y64 = {None: port.tcsc_sgemm_f64(Xu, wo, Bu), 0.2: port.tcsc_sgemm_f64(Xu, wo, Bu, 0.2)}
        for fn, a, gkey in VARIANTS:
            y = call(t, fn, Xu, w, Bu, a)
            ref = golden[f"tcsc.{name}.real.{gkey}"]
            if force == 1:
                assert np.array_equal(y, ref), f"{fn}: tiled kernel must reproduce the reference bit for bit"
            e, e_ref = rel_err(y, y64[a]), rel_err(ref, y64[a])
            assert e <= TOL, f"{fn}: rel err {e:.3e} (reference's own: {e_ref:.3e})"
    finally:
        t.lib().tsg_tcsc_set_kernel(0)
        w.free()


# ---- GEMM: live oracle comparison, tiled kernel, all entry points, ragged shapes -----------------------------------------
SHAPES = [  # M, K, N, num, den, seed
    (64, 512, 512, 1, 2, 42),      # BASELINE.json configs[0]
    (128, 1024, 768, 1, 10, 50),
    (200, 300, 130, 1, 100, 51),   # ragged, many empty columns, single chunk
    (33, 37, 29, 1, 3, 52),        # nothing divisible by anything
    (129, 700, 257, 1, 4, 53),     # one row past a tile, one column past a tile
    (32, 1, 40, 1, 2, 54),         # K = 1
    (256, 2048, 64, 9, 10, 55),    # 10 % sparsity: long lists
    (96, 4096, 96, 1, 20, 56),
]


@pytest.mark.parametrize("shape", SHAPES, ids=[f"M{s[0]}K{s[1]}N{s[2]}d{s[3]}/{s[4]}" for s in SHAPES])
def test_gemm_bit_exact_vs_oracle(t, port, shape):
    M, K, N, num, den, seed = shape
    Wd = port.gen_ternary(K, N, seed, num, den)
    w, wo = t.tcsc_from_dense(Wd), port.tcsc_from_dense(Wd)
    X, B = port.gen_uniform((M, K), seed + 1), port.gen_uniform((N,), seed + 2)
    try:
        for fn, a, _ in VARIANTS:
            y = call(t, fn, X, w, B, a)
            yo = call(port, fn, X, wo, B, a)
            assert np.array_equal(y, yo), f"{fn}: max |d| = {np.abs(y - yo).max():.3e}"
        # SparseGEMM.h entry points on the raw arrays
        y = np.empty((M, N), np.float32)
        t.sparseGEMM(X, *w.arrays(), B, y, M, N, K)
        assert np.array_equal(y, port.sparse_gemm(X, wo, B))
        t.sparseGEMM_PReLU(X, *w.arrays(), B, y, M, N, K, 0.25)
        assert np.array_equal(y, port.sparse_gemm_prelu(X, wo, B, 0.25))
    finally:
        w.free()


@pytest.mark.parametrize("M", [1, 2, 3, 5, 8, 9, 17, 31])
def test_skinny_rows(t, port, M):
    K, N, seed = 1024, 640, 60 + M
    Wd = port.gen_ternary(K, N, seed, 1, 10)
    w, wo = t.tcsc_from_dense(Wd), port.tcsc_from_dense(Wd)
    Xi, B2 = port.gen_intvalued((M, K), seed + 1, 512), np.full(N, 2.0, np.float32)
    assert np.array_equal(t.tcsc_sgemm_prelu_basic(Xi, w, B2, 0.25), port.tcsc_sgemm_prelu_basic(Xi, wo, B2, 0.25))
    X, B = port.gen_uniform((M, K), seed + 1), port.gen_uniform((N,), seed + 2)
    y64 = port.tcsc_sgemm_f64(X, wo, B, 0.2)
    e = rel_err(t.tcsc_sgemm_prelu_basic(X, w, B, 0.2), y64)
    e_ref = rel_err(port.tcsc_sgemm_prelu_basic(X, wo, B, 0.2), y64)
    assert e <= TOL, f"rel err {e:.3e} (reference's own: {e_ref:.3e})"
    w.free()


def test_device_pointers_and_streams(t, port):
    """X, B, Y as CUDA pointers (used in place), on a non-default stream."""
    import torch
    M, K, N = 256, 1024, 512
    Wd = port.gen_ternary(K, N, 70, 1, 10)
    w, wo = t.tcsc_from_dense(Wd), port.tcsc_from_dense(Wd)
    X, B = port.gen_uniform((M, K), 71), port.gen_uniform((N,), 72)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        t.use_torch_stream()
        Xd, Bd = torch.from_numpy(X).cuda(), torch.from_numpy(B).cuda()
        Yd = torch.empty((M, N), device="cuda")
        t.tcsc_sgemm_prelu_basic(Xd, w, Bd, 0.2, Y=Yd)
        s.synchronize()
    t.lib().tsg_set_stream(None)
    assert np.array_equal(Yd.cpu().numpy(), port.tcsc_sgemm_prelu_basic(X, wo, B, 0.2))
    w.free()


def test_caller_assembled_tcsc_struct(t, port):
    """A tcsc_t the caller filled in by hand (never seen by tcsc_from_dense) must work: the mirror is built lazily."""
    import ctypes as C
    M, K, N = 64, 200, 100
    wo = port.tcsc_from_dense(port.gen_ternary(K, N, 80, 1, 4))
    s = t.tcsc_t(K, N, wo.n_elem_pos, wo.n_elem_neg,
                 wo.col_start_pos.ctypes.data_as(C.POINTER(C.c_int)), wo.col_start_neg.ctypes.data_as(C.POINTER(C.c_int)),
                 wo.row_index_pos.ctypes.data_as(C.POINTER(C.c_int)), wo.row_index_neg.ctypes.data_as(C.POINTER(C.c_int)))
    X, B = port.gen_uniform((M, K), 81), port.gen_uniform((N,), 82)
    Y = np.empty((M, N), np.float32)
    t.lib().tcsc_sgemm_prelu_basic(X.ctypes.data, C.byref(s), B.ctypes.data, 0.2, Y.ctypes.data, M, N, K)
    assert t.last_error() == ""
    assert np.array_equal(Y, port.tcsc_sgemm_prelu_basic(X, wo, B, 0.2))


def test_empty_and_degenerate(t, port):
    w = t.tcsc_from_dense(np.zeros((64, 48), np.float32))
    assert w.n_elem_pos == 0 and w.n_elem_neg == 0
    X, B = port.gen_uniform((40, 64), 1), port.gen_uniform((48,), 2)
    y = t.tcsc_sgemm_prelu_basic(X, w, B, 0.2)
    exp = np.where(B < 0, np.float32(0.2) * B, B)[None, :].repeat(40, 0)
    assert np.array_equal(y, exp)
    w.free()
    assert t.lib().tcsc_free(None) is None  # NULL-safe (tcsc.c:168)


def test_error_reporting(t, port):
    w = t.tcsc_from_dense(port.gen_ternary(64, 32, 3, 1, 2))
    X, B = port.gen_uniform((40, 60), 1), port.gen_uniform((32,), 2)
    with pytest.raises(t.TsgError, match="K=60"):
        t.tcsc_sgemm_prelu_basic(X, w, B, 0.2)  # K mismatch: the reference would read out of bounds; we refuse
    w.free()


# ---- BASELINE.json configs[1] at full size: size-independent properties -------------------------------------------------
def test_full_size_cfg2(t, port):
    import torch
    torch.backends.cuda.matmul.allow_tf32 = False
    M = K = N = 4096
    Wd = t.gen_ternary(K, N, 42, 1, 10)
    w = t.DeviceTcsc.from_dense(Wd)
    csp, csn, rip, rin = w.download()
    # structural properties of the index arrays
    assert csp[0] == 0 and csn[0] == 0 and csp[-1] == w.n_pos and csn[-1] == w.n_neg
    assert np.all(np.diff(csp) >= 0) and np.all(np.diff(csn) >= 0)
    assert w.n_pos == int((Wd == 1).sum()) and w.n_neg == int((Wd == -1).sum())
    for n in (0, 1, 2047, 4095):  # sortedness + membership of a few columns
        col = Wd[:, n].cpu().numpy()
        assert np.array_equal(rip[csp[n]:csp[n + 1]], np.nonzero(col == 1)[0])
        assert np.array_equal(rin[csn[n]:csn[n + 1]], np.nonzero(col == -1)[0])
    # (a) integer-valued X: exact in any order => must equal a dense fp32 matmul exactly
    Xi = t.gen_intvalued((M, K), 43, 512)
    B2 = torch.full((N,), 2.0, device="cuda")
    Y = torch.empty((M, N), device="cuda")
    w.gemm(Xi, B2, Y, a=0.25, use_prelu=True)
    ref = Xi @ Wd + B2
    ref = torch.where(ref < 0, 0.25 * ref, ref)
    assert torch.equal(Y, ref)
    # (b) real X: the oracle on a row slice (bit-exact), the fp64 dense check on another slice, linearity in X
    X, B = t.gen_uniform((M, K), 43), t.gen_uniform((N,), 44)
    w.gemm(X, B, Y, a=0.2, use_prelu=True)
    rows = [0, 1, 127, 128, 2048, 4095]
    wo = port.tcsc_from_dense(Wd.cpu().numpy())
    yo = port.tcsc_sgemm_prelu_basic(X[rows].cpu().numpy(), wo, B.cpu().numpy(), 0.2)
    assert np.array_equal(Y[rows].cpu().numpy(), yo)
    # fp64 dense check on another slice.  Y is bit-identical to the reference, so its distance to the fp64 result IS the
    # reference's own (strictly sequential fp32 summation of ~410 terms: ~1.5e-5 at this size); the bar is
    # max(1e-5, reference's own error), both printed.
    rel, ab = t.verify_dense_f64(X, Wd, B, Y, a=0.2, use_prelu=True, m0=1000, mrows=64)
    xs, bs = X[1000:1064].cpu().numpy(), B.cpu().numpy()
    rel_ref = rel_err(port.tcsc_sgemm_prelu_basic(xs, wo, bs, 0.2), port.tcsc_sgemm_f64(xs, wo, bs, 0.2))
    print(f"cfg2 rows 1000..1063: ours vs fp64 {rel:.3e} (abs {ab:.3e}); reference vs fp64 {rel_ref:.3e}")
    assert rel <= max(TOL, rel_ref * (1 + 1e-6)), (rel, rel_ref)
    Z0 = torch.zeros((N,), device="cuda")
    Y1, Y2 = torch.empty_like(Y), torch.empty_like(Y)
    w.gemm(X, Z0, Y1)
    w.gemm(X * 4.0, Z0, Y2)          # scaling by a power of two commutes with every rounding
    assert torch.equal(Y2, Y1 * 4.0)
    w.gemm(-X, Z0, Y2)               # negation too
    assert torch.equal(Y2, -Y1)


def test_deterministic_and_kernel_families_agree(t, port):
    """Same call twice => same bits (the persistent grid's schedule never changes a column's summation order); the
    skinny and the tiled kernel agree to the tolerance on the same input."""
    import torch
    M, K, N = 96, 2048, 1536
    Wd = t.gen_ternary(K, N, 90, 1, 10)
    w = t.DeviceTcsc.from_dense(Wd)
    X, B = t.gen_uniform((M, K), 91), t.gen_uniform((N,), 92)
    Y1, Y2, Y3 = (torch.empty((M, N), device="cuda") for _ in range(3))
    w.gemm(X, B, Y1, a=0.2, use_prelu=True)
    w.gemm(X, B, Y2, a=0.2, use_prelu=True)
    assert torch.equal(Y1, Y2)
    t.lib().tsg_tcsc_set_kernel(2)
    try:
        w.gemm(X, B, Y3, a=0.2, use_prelu=True)
    finally:
        t.lib().tsg_tcsc_set_kernel(0)
    denom = torch.clamp(Y1.abs(), min=1.0)
    assert float(((Y1 - Y3).abs() / denom).max()) <= 2e-5
    # column slab written into a wider Y (ldy > N): neighbours untouched
    Ywide = torch.full((M, N + 64), -7.0, device="cuda")
    t.lib().tsg_tcsc_gemm(w.h, t._ptr(X), t._ptr(B), 0.2, 1, 1, Ywide.data_ptr() + 4 * 32, M, N, K, N + 64)
    torch.cuda.synchronize()
    assert torch.equal(Ywide[:, 32:32 + N], Y1) and bool((Ywide[:, :32] == -7).all()) and bool((Ywide[:, 32 + N:] == -7).all())


def test_host_pointer_pipeline_large(t, port):
    """Host-pointer call big enough for the H2D / kernel / D2H slab pipeline (>= 8 MB), ragged M."""
    M, K, N = 1100, 1024, 1024
    Wd = port.gen_ternary(K, N, 95, 1, 10)
    w, wo = t.tcsc_from_dense(Wd), port.tcsc_from_dense(Wd)
    X, B = port.gen_uniform((M, K), 96), port.gen_uniform((N,), 97)
    y = t.tcsc_sgemm_prelu_basic(X, w, B, 0.2)
    assert np.array_equal(y, port.tcsc_sgemm_prelu_basic(X, wo, B, 0.2))
    y = t.tcsc_sgemm_optimized(X, w, B)
    assert np.array_equal(y, port.tcsc_sgemm_optimized(X, wo, B))
    w.free()


# ---- opt-in TSG_ORDER_FAST: tolerance contract, integer-valued X exact, reference-named entry points behind the switch -----------
@pytest.mark.parametrize("shape", [(4096, 4096, 4096, 1, 10, 42), (300, 1000, 520, 1, 2, 7), (128, 64, 40, 1, 3, 8), (64, 512, 512, 1, 2, 9), (257, 3000, 264, 1, 100, 10),
                                   (257, 300, 100, 1, 2, 11), (32, 53, 129, 2, 3, 12)])  # the last two: dense regime (FFMA2 kernel), ragged rows / K / columns
def test_fast_order(t, port, shape):
    import torch
    M, K, N, num, den, seed = shape
    Wd = t.gen_ternary(K, N, seed, num, den)
    w = t.DeviceTcsc.from_dense(Wd)
    X, B = t.gen_uniform((M, K), seed + 1), t.gen_uniform((N,), seed + 2)
    Yx, Yf = torch.empty((M, N), device="cuda"), torch.empty((M, N), device="cuda")
    w.gemm(X, B, Yx, a=0.2, use_prelu=True, order=t.ORDER_BIAS_LAST)
    w.gemm(X, B, Yf, a=0.2, use_prelu=True, order=t.ORDER_FAST)
    mr = min(M, 64)
    rel_fast, _ = t.verify_dense_f64(X, Wd, B, Yf, a=0.2, use_prelu=True, m0=0, mrows=mr)
    rel_exact, _ = t.verify_dense_f64(X, Wd, B, Yx, a=0.2, use_prelu=True, m0=0, mrows=mr)
    print(f"M{M} K{K} N{N}: fast order vs fp64 {rel_fast:.3e}; reference order vs fp64 {rel_exact:.3e}")
    assert rel_fast <= max(TOL, rel_exact * (1 + 1e-6))  # never worse than the bar the exact order is held to
    Xi, B2 = t.gen_intvalued((M, K), seed + 3, 512), torch.full((N,), 2.0, device="cuda")
    w.gemm(Xi, B2, Yx, order=t.ORDER_BIAS_LAST)
    w.gemm(Xi, B2, Yf, order=t.ORDER_FAST)
    assert torch.equal(Yx, Yf)  # integer-valued: every order is exact
    w.gemm(X, B, Yx, a=0.2, use_prelu=True, order=t.ORDER_BIAS_LAST)  # the two private streams coexist: exact order unchanged afterwards
    wo = port.tcsc_from_dense(Wd.cpu().numpy())
    rows = [0, M // 2, M - 1]
    assert np.array_equal(Yx[rows].cpu().numpy(), port.tcsc_sgemm_prelu_basic(X[rows].cpu().numpy(), wo, B.cpu().numpy(), 0.2))
    w.destroy()


def test_fast_order_switch_for_reference_entry_points(t, port):
    M, K, N = 96, 700, 300
    Wd = port.gen_ternary(K, N, 77, 1, 4)
    X, B = port.gen_uniform((M, K), 78), port.gen_uniform((N,), 79)
    W, Wo = t.tcsc_from_dense(Wd), port.tcsc_from_dense(Wd)
    exact = t.tcsc_sgemm_prelu_basic(X, W, B, 0.2)
    assert np.array_equal(exact, port.tcsc_sgemm_prelu_basic(X, Wo, B, 0.2))
    t.set_fast_order(True)
    try:
        fast = t.tcsc_sgemm_prelu_basic(X, W, B, 0.2)
    finally:
        t.set_fast_order(False)
    y64 = port.tcsc_sgemm_f64(X, Wo, B, 0.2)
    assert rel_err(fast, y64) <= max(TOL, rel_err(exact, y64))
    assert np.array_equal(t.tcsc_sgemm_prelu_basic(X, W, B, 0.2), exact)  # switch off again: the reference's bits
    W.free()


# ---- the fp64 dense checker itself: it is what the full-size tests and bench.py lean on -----------------
def test_verify_dense_f64_equals_numpy_and_catches_errors(t, port):
    import torch
    M, K, N, a = 70, 300, 90, 0.2
    Wd = port.gen_ternary(K, N, 5, 1, 3)
    X, B = port.gen_uniform((M, K), 6), port.gen_uniform((N,), 7)
    y64 = X.astype(np.float64) @ Wd.astype(np.float64) + B.astype(np.float64)
    y64 = np.where(y64 < 0, np.float64(np.float32(a)) * y64, y64)
    Y = y64.astype(np.float32)  # the correctly rounded result: what remains is the fp32 rounding of Y itself
    Xd, Wdd, Bd, Yd = (torch.from_numpy(v).cuda() for v in (X, Wd, B, Y))
    rel, ab = t.verify_dense_f64(Xd, Wdd, Bd, Yd, a=a, use_prelu=True)
    exp_ab = float(np.max(np.abs(Y.astype(np.float64) - y64)))
    exp_rel = float(np.max(np.abs(Y.astype(np.float64) - y64) / np.maximum(np.abs(y64), 1.0)))
    assert abs(ab - exp_ab) <= 1e-12 and abs(rel - exp_rel) <= 1e-12, (rel, exp_rel, ab, exp_ab)  # direct equality with numpy's fp64
    assert rel < 1e-7
    # negative: one wrong element must show up with its exact magnitude, wherever it is
    for (m, n, delta) in ((0, 0, 1e-3), (M - 1, N - 1, 0.5), (33, 17, -2.0)):
        Ybad = Yd.clone()
        Ybad[m, n] += delta
        rel_b, ab_b = t.verify_dense_f64(Xd, Wdd, Bd, Ybad, a=a, use_prelu=True)
        want = abs(float(Ybad[m, n].item()) - y64[m, n])
        assert abs(ab_b - want) <= 1e-9 and rel_b >= want / max(abs(y64[m, n]), 1.0) - 1e-9
        # a row window that excludes the bad element does not see it
        if m > 8:
            rel_w, ab_w = t.verify_dense_f64(Xd, Wdd, Bd, Ybad, a=a, use_prelu=True, m0=0, mrows=8)
            assert ab_w <= exp_ab + 1e-12
    # without the activation, and with a pitch larger than N
    Yp = torch.zeros((M, N + 6), device="cuda")
    y_lin = X.astype(np.float64) @ Wd.astype(np.float64) + B.astype(np.float64)
    Yp[:, :N] = torch.from_numpy(y_lin.astype(np.float32)).cuda()
    rel_l, ab_l = t.verify_dense_f64(Xd, Wdd, Bd, Yp, use_prelu=False, ldy=N + 6)
    assert abs(ab_l - float(np.max(np.abs(y_lin.astype(np.float32).astype(np.float64) - y_lin)))) <= 1e-12
