"""ctypes front-end to the CPU checkers.  TEST INFRASTRUCTURE ONLY.

Two back-ends with the same Python surface:

* ``Port``  -- oracle/liboracle.so, our plain-C restatement (oracle/tsg_oracle.c); always available (gcc).
* ``Ref``   -- oracle/_ref/libref_oracle*.so, the UNMODIFIED reference compiled in place from /root/reference by
               oracle/Makefile; present wherever the build container produced it (it travels to the GPU box with
               the snapshot; /root/reference itself is never read at run time).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(verbose: bool = False) -> None:
    """(Re)build liboracle.so, and oracle/_ref when /root/reference is present.  Building is not using."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        sys.stderr.write(out.stdout + out.stderr)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed")


def _as_f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Tcsc:
    """Host-side TCSC arrays (sparse/tcsc.h:6-17)."""

    def __init__(self, rows, cols, csp, csn, rip, rin):
        self.rows, self.cols = int(rows), int(cols)
        self.col_start_pos, self.col_start_neg = _as_i32(csp), _as_i32(csn)
        self.row_index_pos, self.row_index_neg = _as_i32(rip), _as_i32(rin)

    @property
    def n_elem_pos(self):
        return int(self.row_index_pos.size)

    @property
    def n_elem_neg(self):
        return int(self.row_index_neg.size)

    @property
    def nnz(self):
        return self.n_elem_pos + self.n_elem_neg

    def arrays(self):
        return self.col_start_pos, self.col_start_neg, self.row_index_pos, self.row_index_neg


class Bcsr:
    """Host-side BCSR arrays (sparse/bcsr.h:7-12)."""

    def __init__(self, r, c, br, bc, k, row_start, col_idx, values):
        self.r, self.c, self.br, self.bc, self.k = int(r), int(c), int(br), int(bc), int(k)
        self.b_row_start, self.b_col_idx = _as_i32(row_start), _as_i32(col_idx)
        self.b_values = _as_f32(values)


# =====================================================================================================================
# Port: our restatement
# =====================================================================================================================
class Port:
    kind = "port"

    def __init__(self, path: str | None = None):
        path = path or os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        i, f, d = C.c_int, C.c_float, C.c_double
        L.orc_tcsc_count_f32.argtypes = [_f32p, i, i, C.POINTER(i), C.POINTER(i)]
        L.orc_tcsc_fill_f32.argtypes = [_f32p, i, i, _i32p, _i32p, _i32p, _i32p]
        L.orc_tcsc_count_i32.argtypes = [_i32p, i, i, C.POINTER(i), C.POINTER(i)]
        L.orc_tcsc_fill_i32.argtypes = [_i32p, i, i, _i32p, _i32p, _i32p, _i32p]
        L.orc_bcsr_count.argtypes = [_f32p, i, i, i, i]
        L.orc_bcsr_count.restype = i
        L.orc_bcsr_fill.argtypes = [_f32p, i, i, i, i, i, i, _i32p, _i32p, _f32p]
        tc = [_f32p, _i32p, _i32p, _i32p, _i32p, _f32p]
        for name in ("orc_tcsc_sgemm_basic", "orc_tcsc_sgemm_optimized", "orc_sparse_gemm_f32"):
            getattr(L, name).argtypes = tc + [_f32p, i, i, i]
        for name in ("orc_tcsc_sgemm_prelu_basic", "orc_tcsc_sgemm_prelu_separate", "orc_tcsc_sgemm_prelu_onthego"):
            getattr(L, name).argtypes = tc + [f, _f32p, i, i, i]
        L.orc_sparse_gemm_prelu_f32.argtypes = tc + [_f32p, i, i, i, f]
        L.orc_tcsc_sgemm_f64.argtypes = tc + [i, d, _f64p, i, i, i]
        L.orc_tcsc_abs_mass.argtypes = tc + [_f64p, i, i, i]
        bc = [_f32p, i, i, i, _i32p, _i32p, _f32p, _f32p]
        L.orc_bcsr_sgemm_basic.argtypes = bc + [_f32p, i, i, i]
        L.orc_bcsr_sgemm_prelu_literal.argtypes = bc + [f, _f32p, i, i, i]
        L.orc_bcsr_sgemm_prelu_math.argtypes = bc + [f, _f32p, i, i, i]
        L.orc_gemm_basic.argtypes = [_f32p, _f32p, _f32p, _f32p, i, i, i]
        L.orc_compare.argtypes = [_f32p, _f32p, i, i, f]
        L.orc_compare.restype = i
        L.orc_hash64.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_hash64.restype = C.c_uint64
        L.orc_gen_ternary_f32.argtypes = [_f32p, C.c_int64, C.c_uint64, C.c_uint32, C.c_uint32]
        L.orc_gen_ternary_i32.argtypes = [_i32p, C.c_int64, C.c_uint64, C.c_uint32, C.c_uint32]
        L.orc_gen_uniform_f32.argtypes = [_f32p, C.c_int64, C.c_uint64]
        L.orc_gen_intvalued_f32.argtypes = [_f32p, C.c_int64, C.c_uint64, i]
        L.orc_gen_sparse_pattern_i32.argtypes = [_i32p, i, i, i, i, C.c_uint64]
        L.orc_time_tcsc_sgemm_prelu_basic.argtypes = tc + [f, _f32p, i, i, i, i]
        L.orc_time_tcsc_sgemm_prelu_basic.restype = d

    # ---- builders ---------------------------------------------------------------------------------------------------
    def tcsc_from_dense(self, dense) -> Tcsc:
        if np.asarray(dense).dtype.kind == "i":
            return self.sparse_format(dense)
        dense = _as_f32(dense)
        K, N = dense.shape
        p, q = C.c_int(), C.c_int()
        self.lib.orc_tcsc_count_f32(dense, K, N, C.byref(p), C.byref(q))
        csp, csn = np.empty(N + 1, np.int32), np.empty(N + 1, np.int32)
        rip, rin = np.empty(p.value, np.int32), np.empty(q.value, np.int32)
        self.lib.orc_tcsc_fill_f32(dense, K, N, csp, csn, rip, rin)
        return Tcsc(K, N, csp, csn, rip, rin)

    def sparse_format(self, dense_i32) -> Tcsc:
        dense = _as_i32(dense_i32)
        K, N = dense.shape
        p, q = C.c_int(), C.c_int()
        self.lib.orc_tcsc_count_i32(dense, K, N, C.byref(p), C.byref(q))
        csp, csn = np.empty(N + 1, np.int32), np.empty(N + 1, np.int32)
        rip, rin = np.empty(p.value, np.int32), np.empty(q.value, np.int32)
        self.lib.orc_tcsc_fill_i32(dense, K, N, csp, csn, rip, rin)
        return Tcsc(K, N, csp, csn, rip, rin)

    def bcsr_from_dense(self, dense, r, c, quirk=False, tail_fill=-1) -> Bcsr:
        dense = _as_f32(dense)
        K, N = dense.shape
        br, bc = K // r, N // c
        k = self.lib.orc_bcsr_count(dense, K, N, r, c)
        rs, ci = np.empty(br + 1, np.int32), np.empty(k, np.int32)
        vals = np.empty(k * r * c, np.float32)
        self.lib.orc_bcsr_fill(dense, K, N, r, c, int(bool(quirk)), int(tail_fill), rs, ci, vals)
        return Bcsr(r, c, br, bc, k, rs, ci, vals)

    # ---- kernels ----------------------------------------------------------------------------------------------------
    def _tc(self, name, X, W: Tcsc, B, a=None):
        X, B = _as_f32(X), _as_f32(B)
        M, K = X.shape
        N = W.cols
        Y = np.empty((M, N), np.float32)
        fn = getattr(self.lib, name)
        if a is None:
            fn(X, *W.arrays(), B, Y, M, N, K)
        else:
            fn(X, *W.arrays(), B, float(a), Y, M, N, K)
        return Y

    def tcsc_sgemm_basic(self, X, W, B):
        return self._tc("orc_tcsc_sgemm_basic", X, W, B)

    def tcsc_sgemm_optimized(self, X, W, B):
        return self._tc("orc_tcsc_sgemm_optimized", X, W, B)

    def tcsc_sgemm_prelu_basic(self, X, W, B, a):
        return self._tc("orc_tcsc_sgemm_prelu_basic", X, W, B, a)

    def tcsc_sgemm_prelu_optimized_separate(self, X, W, B, a):
        return self._tc("orc_tcsc_sgemm_prelu_separate", X, W, B, a)

    def tcsc_sgemm_prelu_optimized_onthego(self, X, W, B, a):
        return self._tc("orc_tcsc_sgemm_prelu_onthego", X, W, B, a)

    def sparse_gemm(self, X, W, b):
        return self._tc("orc_sparse_gemm_f32", X, W, b)

    def sparse_gemm_prelu(self, X, W, b, a):
        X, b = _as_f32(X), _as_f32(b)
        M, K = X.shape
        Y = np.empty((M, W.cols), np.float32)
        self.lib.orc_sparse_gemm_prelu_f32(X, *W.arrays(), b, Y, M, W.cols, K, float(a))
        return Y

    def tcsc_sgemm_f64(self, X, W, B, a=None):
        X, B = _as_f32(X), _as_f32(B)
        M, K = X.shape
        Y = np.empty((M, W.cols), np.float64)
        self.lib.orc_tcsc_sgemm_f64(X, *W.arrays(), B, 0 if a is None else 1, 0.0 if a is None else float(a), Y, M, W.cols, K)
        return Y

    def tcsc_abs_mass(self, X, W, B):
        X, B = _as_f32(X), _as_f32(B)
        M, K = X.shape
        S = np.empty((M, W.cols), np.float64)
        self.lib.orc_tcsc_abs_mass(X, *W.arrays(), B, S, M, W.cols, K)
        return S

    def _bc(self, name, X, W: Bcsr, B, N, a=None):
        X, B = _as_f32(X), _as_f32(B)
        M, K = X.shape
        Y = np.empty((M, N), np.float32)
        fn = getattr(self.lib, name)
        head = (X, W.r, W.c, W.br, W.b_row_start, W.b_col_idx, W.b_values, B)
        if a is None:
            fn(*head, Y, M, N, K)
        else:
            fn(*head, float(a), Y, M, N, K)
        return Y

    def bcsr_sgemm_basic(self, X, W, B, N):
        return self._bc("orc_bcsr_sgemm_basic", X, W, B, N)

    def bcsr_sgemm_prelu_literal(self, X, W, B, a, N):
        return self._bc("orc_bcsr_sgemm_prelu_literal", X, W, B, N, a)

    def bcsr_sgemm_prelu_math(self, X, W, B, a, N):
        return self._bc("orc_bcsr_sgemm_prelu_math", X, W, B, N, a)

    def gemm_basic(self, X, Wd, B):
        X, Wd, B = _as_f32(X), _as_f32(Wd), _as_f32(B)
        M, K = X.shape
        N = Wd.shape[1]
        Y = np.empty((M, N), np.float32)
        self.lib.orc_gemm_basic(X, Wd, B, Y, M, N, K)
        return Y

    def compare(self, result, target, tol=1e-4):
        result, target = _as_f32(result), _as_f32(target)
        return bool(self.lib.orc_compare(result, target, result.shape[0], result.shape[1], float(tol)))

    # ---- generators -------------------------------------------------------------------------------------------------
    def gen_ternary(self, K, N, seed, num, den, dtype=np.float32):
        W = np.empty((K, N), dtype)
        if dtype == np.float32:
            self.lib.orc_gen_ternary_f32(W, K * N, seed, num, den)
        else:
            self.lib.orc_gen_ternary_i32(W, K * N, seed, num, den)
        return W

    def gen_uniform(self, shape, seed):
        X = np.empty(shape, np.float32)
        self.lib.orc_gen_uniform_f32(X, X.size, seed)
        return X

    def gen_intvalued(self, shape, seed, rng=512):
        X = np.empty(shape, np.float32)
        self.lib.orc_gen_intvalued_f32(X, X.size, seed, rng)
        return X

    def gen_sparse_pattern(self, H, W, non_zero, uniform, seed):
        """generateSparseMatrix<int> (SparseGEMM.h:53-102) with counter-based draws; int32 H x W"""
        out = np.empty((H, W), np.int32)
        if self.lib.orc_gen_sparse_pattern_i32(out, H, W, non_zero, int(bool(uniform)), seed) != 0:
            raise ValueError("gen_sparse_pattern: bad arguments")
        return out

    def time_prelu_basic(self, X, W: Tcsc, B, a, reps=1):
        X, B = _as_f32(X), _as_f32(B)
        M, K = X.shape
        Y = np.empty((M, W.cols), np.float32)
        return float(self.lib.orc_time_tcsc_sgemm_prelu_basic(X, *W.arrays(), B, float(a), Y, M, W.cols, K, reps))


# =====================================================================================================================
# Ref: the unmodified reference
# =====================================================================================================================
def ref_path(variant: str = "") -> str:
    suffix = {"": "", "fm": "_fm", "native": "_native", "omp": "_omp"}[variant]
    return os.path.join(HERE, "_ref", f"libref_oracle{suffix}.so")


def ref_available(variant: str = "") -> bool:
    return os.path.exists(ref_path(variant))


def ref_native_runs() -> bool:
    """True when the -march=native build of the reference executes on THIS host (probed in a subprocess so an
    illegal-instruction fault cannot take the caller down)."""
    if not ref_available("native"):
        return False
    code = (
        "import sys; sys.path.insert(0, %r); import numpy as np; from oracle.pyoracle import Ref; r = Ref('native');"
        "W = r.tcsc_from_dense(np.eye(64, dtype=np.float32)); X = np.ones((4, 64), np.float32);"
        "r.tcsc_sgemm_prelu_basic(X, W, np.zeros(64, np.float32), 0.2);"
        "Wb = r.bcsr_from_dense(np.eye(64, dtype=np.float32), 1, 8); r.bcsr_sgemm_avx(X, Wb, np.zeros(64, np.float32), 64)"
    ) % os.path.dirname(HERE)
    try:
        return subprocess.run([sys.executable, "-c", code], capture_output=True, timeout=120).returncode == 0
    except Exception:
        return False


class _RefTcsc(Tcsc):
    """TCSC arrays copied out of a reference tcsc_t; keeps the reference object alive for kernel calls."""

    def __init__(self, ref, handle, rows, cols):
        L = ref.lib
        npos, nneg = L.ref_tcsc_n_pos(handle), L.ref_tcsc_n_neg(handle)

        def grab(fn, n):
            return np.ctypeslib.as_array(fn(handle), shape=(n,)).copy() if n else np.empty(0, np.int32)

        super().__init__(rows, cols, grab(L.ref_tcsc_col_start_pos, cols + 1), grab(L.ref_tcsc_col_start_neg, cols + 1),
                         grab(L.ref_tcsc_row_index_pos, npos), grab(L.ref_tcsc_row_index_neg, nneg))
        self._ref, self.handle = ref, handle

    def __del__(self):
        try:
            self._ref.lib.ref_tcsc_free(self.handle)
        except Exception:
            pass


class _RefBcsr(Bcsr):
    def __init__(self, ref, handle):
        L = ref.lib
        d = (C.c_int * 5)()
        L.ref_bcsr_dims(handle, d)
        r, c, br, bc, k = list(d)

        def grab(fn, n, dt):
            return np.ctypeslib.as_array(fn(handle), shape=(n,)).copy() if n else np.empty(0, dt)

        super().__init__(r, c, br, bc, k, grab(L.ref_bcsr_row_start, br + 1, np.int32), grab(L.ref_bcsr_col_idx, k, np.int32),
                         grab(L.ref_bcsr_values, k * r * c, np.float32))
        self._ref, self.handle = ref, handle

    def __del__(self):
        try:
            self._ref.lib.ref_bcsr_free(self.handle)
        except Exception:
            pass


class Ref:
    kind = "reference"

    def __init__(self, variant: str = ""):
        path = ref_path(variant)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (built only where /root/reference exists: make -C oracle ref)")
        self.variant = variant
        self.lib = L = C.CDLL(path)
        i, f, d, vp = C.c_int, C.c_float, C.c_double, C.c_void_p
        ip, fp = C.POINTER(C.c_int), C.POINTER(C.c_float)
        L.ref_tcsc_from_dense.argtypes = [_f32p, i, i]
        L.ref_tcsc_from_dense.restype = vp
        L.ref_tcsc_free.argtypes = [vp]
        for n in ("ref_tcsc_n_pos", "ref_tcsc_n_neg"):
            getattr(L, n).argtypes, getattr(L, n).restype = [vp], i
        for n in ("ref_tcsc_col_start_pos", "ref_tcsc_col_start_neg", "ref_tcsc_row_index_pos", "ref_tcsc_row_index_neg"):
            getattr(L, n).argtypes, getattr(L, n).restype = [vp], ip
        for n in ("ref_tcsc_sgemm_basic", "ref_tcsc_sgemm_optimized"):
            getattr(L, n).argtypes = [_f32p, vp, _f32p, _f32p, i, i, i]
        for n in ("ref_tcsc_sgemm_prelu_basic", "ref_tcsc_sgemm_prelu_optimized_separate", "ref_tcsc_sgemm_prelu_optimized_onthego"):
            getattr(L, n).argtypes = [_f32p, vp, _f32p, f, _f32p, i, i, i]
        L.ref_gemm_basic.argtypes = [_f32p, _f32p, _f32p, _f32p, i, i, i]
        L.ref_compare.argtypes, L.ref_compare.restype = [_f32p, _f32p, i, i], i
        L.ref_bcsr_from_dense.argtypes, L.ref_bcsr_from_dense.restype = [_f32p, i, i, i, i], vp
        L.ref_bcsr_free.argtypes = [vp]
        L.ref_bcsr_dims.argtypes = [vp, C.POINTER(C.c_int * 5)]
        L.ref_bcsr_row_start.argtypes, L.ref_bcsr_row_start.restype = [vp], ip
        L.ref_bcsr_col_idx.argtypes, L.ref_bcsr_col_idx.restype = [vp], ip
        L.ref_bcsr_values.argtypes, L.ref_bcsr_values.restype = [vp], fp
        for n in ("ref_bcsr_sgemm_basic", "ref_bcsr_sgemm_avx", "ref_bcsr_sgemm_avx2"):
            getattr(L, n).argtypes = [_f32p, vp, _f32p, _f32p, i, i, i]
        for n in ("ref_bcsr_sgemm_prelu_basic", "ref_bcsr_sgemm_prelu_avx"):
            getattr(L, n).argtypes = [_f32p, vp, _f32p, f, _f32p, i, i, i]
        L.ref_sparseformat_new.argtypes, L.ref_sparseformat_new.restype = [_i32p, i, i], vp
        L.ref_sparseformat_delete.argtypes = [vp]
        L.ref_sparseformat_sizes.argtypes = [vp, C.POINTER(C.c_int * 4)]
        L.ref_sparseformat_array.argtypes, L.ref_sparseformat_array.restype = [vp, i], ip
        L.ref_sparseGEMM_f32.argtypes = [_f32p, _i32p, _i32p, _i32p, _i32p, _f32p, _f32p, i, i, i]
        L.ref_sparseGEMM_PReLU_f32.argtypes = [_f32p, _i32p, _i32p, _i32p, _i32p, _f32p, _f32p, i, i, i, f]
        L.ref_GEMM_f32.argtypes = [_f32p, _f32p, _f32p, _f32p, i, i, i]
        L.ref_GEMM_PReLU_f32.argtypes = [_f32p, _f32p, _f32p, _f32p, i, i, i, f]
        L.ref_time_tcsc_sgemm_prelu_basic.argtypes = [_f32p, vp, _f32p, f, _f32p, i, i, i, i]
        L.ref_time_tcsc_sgemm_prelu_basic.restype = d
        L.ref_build_flags.restype = C.c_char_p

    @property
    def build_flags(self) -> str:
        return self.lib.ref_build_flags().decode()

    # ---- builders ---------------------------------------------------------------------------------------------------
    def tcsc_from_dense(self, dense) -> Tcsc:
        if np.asarray(dense).dtype.kind == "i":
            return self.sparse_format(dense)
        dense = _as_f32(dense)
        K, N = dense.shape
        return _RefTcsc(self, self.lib.ref_tcsc_from_dense(dense, K, N), K, N)

    def sparse_format(self, dense_i32) -> Tcsc:
        dense = _as_i32(dense_i32)
        K, N = dense.shape
        h = self.lib.ref_sparseformat_new(dense, K, N)
        sz = (C.c_int * 4)()
        self.lib.ref_sparseformat_sizes(h, sz)
        arrs = [np.ctypeslib.as_array(self.lib.ref_sparseformat_array(h, w), shape=(sz[w],)).copy() if sz[w] else np.empty(0, np.int32)
                for w in range(4)]
        self.lib.ref_sparseformat_delete(h)
        return Tcsc(K, N, *arrs)

    def bcsr_from_dense(self, dense, r, c) -> Bcsr:
        dense = _as_f32(dense)
        K, N = dense.shape
        return _RefBcsr(self, self.lib.ref_bcsr_from_dense(dense, K, N, r, c))

    # ---- kernels ----------------------------------------------------------------------------------------------------
    def _handle(self, W):
        if not isinstance(W, (_RefTcsc, _RefBcsr)):
            raise TypeError("Ref kernels need a W built by Ref.*_from_dense")
        return W.handle

    def _tc(self, name, X, W, B, a=None):
        X, B = _as_f32(X), _as_f32(B)
        M, K = X.shape
        Y = np.empty((M, W.cols), np.float32)
        fn = getattr(self.lib, name)
        if a is None:
            fn(X, self._handle(W), B, Y, M, W.cols, K)
        else:
            fn(X, self._handle(W), B, float(a), Y, M, W.cols, K)
        return Y

    def tcsc_sgemm_basic(self, X, W, B):
        return self._tc("ref_tcsc_sgemm_basic", X, W, B)

    def tcsc_sgemm_optimized(self, X, W, B):
        return self._tc("ref_tcsc_sgemm_optimized", X, W, B)

    def tcsc_sgemm_prelu_basic(self, X, W, B, a):
        return self._tc("ref_tcsc_sgemm_prelu_basic", X, W, B, a)

    def tcsc_sgemm_prelu_optimized_separate(self, X, W, B, a):
        return self._tc("ref_tcsc_sgemm_prelu_optimized_separate", X, W, B, a)

    def tcsc_sgemm_prelu_optimized_onthego(self, X, W, B, a):
        return self._tc("ref_tcsc_sgemm_prelu_optimized_onthego", X, W, B, a)

    def sparse_gemm(self, X, W: Tcsc, b):
        X, b = _as_f32(X), _as_f32(b)
        M, K = X.shape
        Y = np.empty((M, W.cols), np.float32)
        self.lib.ref_sparseGEMM_f32(X, *W.arrays(), b, Y, M, W.cols, K)
        return Y

    def sparse_gemm_prelu(self, X, W: Tcsc, b, a):
        X, b = _as_f32(X), _as_f32(b)
        M, K = X.shape
        Y = np.empty((M, W.cols), np.float32)
        self.lib.ref_sparseGEMM_PReLU_f32(X, *W.arrays(), b, Y, M, W.cols, K, float(a))
        return Y

    def _bc(self, name, X, W, B, N, a=None):
        X, B = _as_f32(X), _as_f32(B)
        M, K = X.shape
        # the AVX kernels need 32-byte aligned B and Y (bcsr.c:229-230)
        Y = _aligned_empty((M, N))
        Ba = _aligned_empty(B.shape)
        Ba[...] = B
        fn = getattr(self.lib, name)
        if a is None:
            fn(X, self._handle(W), Ba, Y, M, N, K)
        else:
            fn(X, self._handle(W), Ba, float(a), Y, M, N, K)
        return Y.copy()

    def bcsr_sgemm_basic(self, X, W, B, N):
        return self._bc("ref_bcsr_sgemm_basic", X, W, B, N)

    def bcsr_sgemm_avx(self, X, W, B, N):
        return self._bc("ref_bcsr_sgemm_avx", X, W, B, N)

    def bcsr_sgemm_avx2(self, X, W, B, N):
        return self._bc("ref_bcsr_sgemm_avx2", X, W, B, N)

    def bcsr_sgemm_prelu_basic(self, X, W, B, a, N):
        return self._bc("ref_bcsr_sgemm_prelu_basic", X, W, B, N, a)

    def bcsr_sgemm_prelu_avx(self, X, W, B, a, N):
        return self._bc("ref_bcsr_sgemm_prelu_avx", X, W, B, N, a)

    def gemm_basic(self, X, Wd, B):
        X, Wd, B = _as_f32(X), _as_f32(Wd), _as_f32(B)
        M, K = X.shape
        N = Wd.shape[1]
        Y = np.empty((M, N), np.float32)
        self.lib.ref_gemm_basic(X, Wd, B, Y, M, N, K)
        return Y

    def gemm_prelu(self, X, Wd, B, a):
        X, Wd, B = _as_f32(X), _as_f32(Wd), _as_f32(B)
        M, K = X.shape
        N = Wd.shape[1]
        Y = np.empty((M, N), np.float32)
        self.lib.ref_GEMM_PReLU_f32(X, Wd, B, Y, M, N, K, float(a))
        return Y

    def compare(self, result, target):
        result, target = _as_f32(result), _as_f32(target)
        return bool(self.lib.ref_compare(result, target, result.shape[0], result.shape[1]))

    def time_prelu_basic(self, X, W, B, a, reps=1):
        X, B = _as_f32(X), _as_f32(B)
        M, K = X.shape
        Y = np.empty((M, W.cols), np.float32)
        return float(self.lib.ref_time_tcsc_sgemm_prelu_basic(X, self._handle(W), B, float(a), Y, M, W.cols, K, reps))

    def omp_max_threads(self) -> int:
        return int(self.lib.ref_omp_max_threads())

    def time_sparse_gemm_prelu(self, X, W, B, a, reps=1):
        """seconds of sparseGEMM_PReLU<float> (SparseGEMM.h:151-168; `omp parallel for` over m in the omp variant)"""
        X, B = _as_f32(X), _as_f32(B)
        M, K = X.shape
        Y = np.empty((M, W.cols), np.float32)
        csp, csn, rip, rin = W.arrays()
        self.lib.ref_time_sparseGEMM_PReLU_f32.restype = C.c_double
        self.lib.ref_time_sparseGEMM_PReLU_f32.argtypes = [_f32p, _i32p, _i32p, _i32p, _i32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int,
                                                         C.c_float, C.c_int]
        rip = rip if rip.size else np.zeros(1, np.int32)
        rin = rin if rin.size else np.zeros(1, np.int32)
        return float(self.lib.ref_time_sparseGEMM_PReLU_f32(X, csp, csn, rip, rin, B, Y, M, W.cols, K, float(a), reps))


def _aligned_empty(shape, dtype=np.float32, align=32):
    n = int(np.prod(shape)) if len(shape) else 1
    raw = np.empty(n * np.dtype(dtype).itemsize + align, np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n * np.dtype(dtype).itemsize].view(dtype).reshape(shape)
