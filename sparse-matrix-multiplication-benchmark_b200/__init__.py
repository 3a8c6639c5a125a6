"""Python host-side mirror of the reference's operator interface, over libtsgemm_b200.so's C-ABI (ctypes).

The product is the C library (include/*.h); this module exists so that tests and bench.py can call it the way the
reference's own callers do -- same function names, argument order and error behaviour as sparse/tcsc.h:19-48,
sparse/bcsr.h:14-39 and SparseGEMM.h:13-40,104-168 -- with numpy arrays (host pointers) or torch CUDA tensors (device
pointers).  torch is used for device memory, streams and torch.distributed only.

There is no CPU fallback anywhere: if the shared library is missing this module raises on import of `lib()`, and
without a B200 every entry point reports the failure through sparse_last_error().

The directory name contains hyphens, so import it through `load()` in __graft_entry__.py (importlib), which
registers it as `tsgemm_b200`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TSG_LIB_PATH") or os.path.join(HERE, "libtsgemm_b200.so")  # TSG_LIB_PATH: A/B another build of the SAME library

ORDER_BIAS_FIRST, ORDER_BIAS_LAST, ORDER_SPLIT, ORDER_FAST = 0, 1, 2, 3
SKINNY_M = 32

_lib = None


class TsgError(RuntimeError):
    pass


class tcsc_t(C.Structure):  # include/sparse/tcsc.h  (reference sparse/tcsc.h:6-17)
    _fields_ = [("rows", C.c_int), ("cols", C.c_int), ("n_elem_pos", C.c_int), ("n_elem_neg", C.c_int),
                ("col_start_pos", C.POINTER(C.c_int)), ("col_start_neg", C.POINTER(C.c_int)),
                ("row_index_pos", C.POINTER(C.c_int)), ("row_index_neg", C.POINTER(C.c_int))]


class bcsr_t(C.Structure):  # include/sparse/bcsr.h  (reference sparse/bcsr.h:7-12)
    _fields_ = [("r", C.c_int), ("c", C.c_int), ("br", C.c_int), ("bc", C.c_int), ("k", C.c_int),
                ("b_row_start", C.POINTER(C.c_int)), ("b_col_idx", C.POINTER(C.c_int)), ("b_values", C.POINTER(C.c_float))]


def _build_in_tree() -> None:
    import shutil
    import subprocess
    if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        return
    out = subprocess.run(["make", "-C", HERE, "-j8", "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise TsgError("building libtsgemm_b200.so failed:\n" + out.stdout[-2000:] + out.stderr[-2000:])


def lib() -> C.CDLL:
    """Load libtsgemm_b200.so (built in-tree by `make -C sparse-matrix-multiplication-benchmark_b200`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) and "TSG_LIB_PATH" not in os.environ:
        _build_in_tree()  # a source-only checkout: compile the CUDA library (nvcc, sm_100a), never substitute for it
    if not os.path.exists(LIB_PATH):
        raise TsgError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() -- there is no CPU fallback")
    L = C.CDLL(LIB_PATH)  # RTLD_LOCAL: our reference-named symbols must not interpose on other libraries
    i, f, vp, ll = C.c_int, C.c_float, C.c_void_p, C.c_longlong
    ip = C.POINTER(C.c_int)
    # ---- reference-named entry points
    L.tcsc_from_dense.argtypes, L.tcsc_from_dense.restype = [vp, i, i], C.POINTER(tcsc_t)
    L.tcsc_free.argtypes, L.tcsc_free.restype = [C.POINTER(tcsc_t)], None
    L.tcsc_invalidate.argtypes, L.tcsc_invalidate.restype = [C.POINTER(tcsc_t)], None
    for n in ("tcsc_sgemm_basic", "tcsc_sgemm_optimized"):
        getattr(L, n).argtypes, getattr(L, n).restype = [vp, C.POINTER(tcsc_t), vp, vp, i, i, i], None
    for n in ("tcsc_sgemm_prelu_basic", "tcsc_sgemm_prelu_optimized_separate", "tcsc_sgemm_prelu_optimized_onthego"):
        getattr(L, n).argtypes, getattr(L, n).restype = [vp, C.POINTER(tcsc_t), vp, f, vp, i, i, i], None
    L.bcsr_from_dense.argtypes, L.bcsr_from_dense.restype = [vp, i, i, i, i], C.POINTER(bcsr_t)
    for n in ("bcsr_sgemm_basic", "bcsr_sgemm_avx", "bcsr_sgemm_avx2"):
        getattr(L, n).argtypes, getattr(L, n).restype = [vp, bcsr_t, vp, vp, i, i, i], None
    for n in ("bcsr_sgemm_prelu_basic", "bcsr_sgemm_prelu_avx"):
        getattr(L, n).argtypes, getattr(L, n).restype = [vp, bcsr_t, vp, f, vp, i, i, i], None
    L.bcsr_release_device.argtypes = [C.POINTER(bcsr_t)]
    L.sparse_last_error.restype = C.c_char_p
    L.tsg_sparse_gemm_f32.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, i, i, f, i]
    L.tsg_sparse_format_build_i32.argtypes = [vp, i, i, C.POINTER(vp), ip, ip]
    L.tsg_sparse_format_fetch.argtypes = [vp, vp, vp, vp, vp]
    # ---- device-level API
    L.tsg_last_error.restype = C.c_char_p
    L.tsg_version.restype = C.c_char_p
    L.tsg_set_stream.argtypes = [vp]
    L.tsg_get_stream.restype = vp
    L.tsg_launch_count.restype = ll
    L.tsg_dev_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.tsg_dev_free.argtypes = [vp]
    L.tsg_tcsc_from_dense_f32.argtypes = [vp, i, i, C.POINTER(vp)]
    L.tsg_tcsc_from_dense_i32.argtypes = [vp, i, i, C.POINTER(vp)]
    L.tsg_tcsc_from_arrays.argtypes = [vp, vp, vp, vp, i, i, C.POINTER(vp)]
    L.tsg_tcsc_destroy.argtypes, L.tsg_tcsc_destroy.restype = [vp], None
    L.tsg_tcsc_dims.argtypes = [vp, ip, ip, ip, ip]
    L.tsg_tcsc_download.argtypes = [vp, vp, vp, vp, vp]
    L.tsg_tcsc_stream_info.argtypes = [vp, C.POINTER(ll), ip, ip]
    L.tsg_tcsc_gemm.argtypes = [vp, vp, vp, f, i, i, vp, i, i, i, ll]
    L.tsg_tcsc_set_kernel.argtypes = [i]
    L.tsg_profile_enable.argtypes = [i]
    L.tsg_profile_read.argtypes = [C.POINTER(C.c_double), ip]
    L.tsg_bcsr_from_dense_f32.argtypes = [vp, i, i, i, i, C.POINTER(vp)]
    L.tsg_bcsr_from_arrays.argtypes = [vp, vp, vp, i, i, i, i, i, C.POINTER(vp)]
    L.tsg_bcsr_destroy.argtypes, L.tsg_bcsr_destroy.restype = [vp], None
    L.tsg_bcsr_dims.argtypes = [vp, ip, ip, ip, ip, ip]
    L.tsg_bcsr_download.argtypes = [vp, vp, vp, vp]
    L.tsg_bcsr_gemm.argtypes = [vp, vp, vp, f, i, vp, i, i, i, ll]
    L.tsg_bcsr_set_kernel.argtypes = [i]
    L.tsg_bcsr_set_prelu_literal.argtypes = [i]
    L.tsg_bcsr_get_prelu_literal.argtypes = []
    L.tsg_gen_ternary_f32.argtypes = [vp, ll, C.c_uint64, C.c_uint32, C.c_uint32]
    L.tsg_gen_ternary_i32.argtypes = [vp, ll, C.c_uint64, C.c_uint32, C.c_uint32]
    L.tsg_gen_ternary_slice_f32.argtypes = [vp, i, i, i, i, C.c_uint64, C.c_uint32, C.c_uint32]
    L.tsg_gen_uniform_f32.argtypes = [vp, ll, C.c_uint64]
    L.tsg_gen_sparse_pattern_i32.argtypes = [vp, i, i, i, i, C.c_uint64]
    L.tsg_gen_sparse_pattern_f32.argtypes = [vp, i, i, i, i, C.c_uint64]
    L.tsg_gen_intvalued_f32.argtypes = [vp, ll, C.c_uint64, i]
    L.tsg_verify_dense_f64.argtypes = [vp, vp, vp, f, i, vp, i, i, i, ll, i, i, C.POINTER(C.c_double * 2)]
    L.tsg_dist_unique_id.argtypes = [C.POINTER(C.c_ubyte * 128)]
    L.tsg_dist_create.argtypes = [C.POINTER(C.c_ubyte * 128), i, i, C.POINTER(vp)]
    L.tsg_dist_destroy.argtypes, L.tsg_dist_destroy.restype = [vp], None
    L.tsg_dist_partition.argtypes, L.tsg_dist_partition.restype = [i, i, i, ip, ip], None
    L.tsg_dist_alloc_y.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.tsg_dist_barrier.argtypes = [vp]
    L.tsg_dist_has_multicast.argtypes = [vp]
    L.tsg_dist_gemm.argtypes = [vp, vp, vp, i, vp, f, i, i, vp, i, i, i, i]
    L.tsg_dist_gemm_host.argtypes = [vp, vp, vp, vp, f, i, i, vp, i, i, i, i]
    L.tsg_host_register.argtypes = [vp, C.c_size_t]
    L.tsg_host_unregister.argtypes = [vp]
    _lib = L
    return L


def last_error() -> str:
    return lib().sparse_last_error().decode()


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise TsgError(f"{what} failed (code {rc}): {last_error()}")


# ---- pointer plumbing -------------------------------------------------------------------------------------------------
def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _ptr(x, dtype=None) -> int:
    """Raw address of a numpy array (host) or torch tensor (host or device); the object must be contiguous."""
    if x is None:
        return 0
    if _is_torch(x):
        if not x.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return x.data_ptr()
    if not isinstance(x, np.ndarray):
        raise TypeError(f"expected numpy array or torch tensor, got {type(x)}")
    if dtype is not None and x.dtype != dtype:
        raise TypeError(f"expected dtype {dtype}, got {x.dtype}")
    if not x.flags["C_CONTIGUOUS"]:
        raise ValueError("array must be C-contiguous")
    return x.ctypes.data


def use_torch_stream() -> None:
    """Point the library at torch's current CUDA stream (call after torch.cuda.set_device / inside a stream ctx)."""
    import torch
    lib().tsg_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))


def set_fast_order(on: bool) -> None:
    """opt-in: the reference-named entry points use TSG_ORDER_FAST (tolerance contract) instead of the exact order"""
    lib().tsg_set_fast_order(int(bool(on)))


def profile_enable(on: bool) -> None:
    lib().tsg_profile_enable(int(on))


def profile_read():
    """(total device ms, launches) of the tiled GEMM kernel since the last call."""
    ms, n = C.c_double(), C.c_int()
    lib().tsg_profile_read(C.byref(ms), C.byref(n))
    return ms.value, n.value


def launch_count(reset: bool = False) -> int:
    n = int(lib().tsg_launch_count())
    if reset:
        lib().tsg_launch_count_reset()
    return n


# =====================================================================================================================
# reference-named host API (include/sparse/tcsc.h, include/sparse/bcsr.h, include/SparseGEMM.h)
# =====================================================================================================================
class Tcsc:
    """Owner of a `tcsc_t*` returned by tcsc_from_dense; exposes the struct's host arrays as numpy views."""

    def __init__(self, handle):
        self.handle = handle

    @property
    def s(self) -> tcsc_t:
        return self.handle.contents

    rows = property(lambda self: self.s.rows)
    cols = property(lambda self: self.s.cols)
    n_elem_pos = property(lambda self: self.s.n_elem_pos)
    n_elem_neg = property(lambda self: self.s.n_elem_neg)

    def _view(self, p, n):
        return np.ctypeslib.as_array(p, shape=(n,)) if n else np.empty(0, np.int32)

    col_start_pos = property(lambda self: self._view(self.s.col_start_pos, self.s.cols + 1))
    col_start_neg = property(lambda self: self._view(self.s.col_start_neg, self.s.cols + 1))
    row_index_pos = property(lambda self: self._view(self.s.row_index_pos, self.s.n_elem_pos))
    row_index_neg = property(lambda self: self._view(self.s.row_index_neg, self.s.n_elem_neg))

    def arrays(self):
        return self.col_start_pos, self.col_start_neg, self.row_index_pos, self.row_index_neg

    def free(self):
        if self.handle:
            lib().tcsc_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def tcsc_from_dense(dense, rows=None, cols=None) -> Tcsc:
    """tcsc_t *tcsc_from_dense(dense_t dense, int rows, int cols) -- reference sparse/tcsc.c:6-66.  Raises on NULL."""
    if rows is None:
        rows, cols = dense.shape
    h = lib().tcsc_from_dense(_ptr(dense, np.float32), rows, cols)
    if not h:
        raise TsgError("tcsc_from_dense returned NULL: " + last_error())
    return Tcsc(h)


def _out(Y, X, M, N):
    if Y is not None:
        return Y
    if _is_torch(X):
        import torch
        return torch.empty((M, N), dtype=torch.float32, device=X.device)
    return np.empty((M, N), np.float32)


def _gemm(name, X, W: Tcsc, B, a, Y, M, N, K):
    M = X.shape[0] if M is None else M
    K = X.shape[1] if K is None else K
    N = W.cols if N is None else N
    Y = _out(Y, X, M, N)
    fn = getattr(lib(), name)
    if a is None:
        fn(_ptr(X, np.float32), W.handle, _ptr(B, np.float32), _ptr(Y, np.float32), M, N, K)
    else:
        fn(_ptr(X, np.float32), W.handle, _ptr(B, np.float32), float(a), _ptr(Y, np.float32), M, N, K)
    err = last_error()
    if err:
        raise TsgError(f"{name}: {err}")
    return Y


def tcsc_sgemm_basic(X, W, B, Y=None, M=None, N=None, K=None):
    return _gemm("tcsc_sgemm_basic", X, W, B, None, Y, M, N, K)


def tcsc_sgemm_optimized(X, W, B, Y=None, M=None, N=None, K=None):
    return _gemm("tcsc_sgemm_optimized", X, W, B, None, Y, M, N, K)


def tcsc_sgemm_prelu_basic(X, W, B, a, Y=None, M=None, N=None, K=None):
    return _gemm("tcsc_sgemm_prelu_basic", X, W, B, a, Y, M, N, K)


def tcsc_sgemm_prelu_optimized_separate(X, W, B, a, Y=None, M=None, N=None, K=None):
    return _gemm("tcsc_sgemm_prelu_optimized_separate", X, W, B, a, Y, M, N, K)


def tcsc_sgemm_prelu_optimized_onthego(X, W, B, a, Y=None, M=None, N=None, K=None):
    return _gemm("tcsc_sgemm_prelu_optimized_onthego", X, W, B, a, Y, M, N, K)


class SparseFormat:
    """SparseFormat(int* matrix, K, N) -- reference SparseGEMM.h:13-40 (int32 matrix, predicates >=1 / <=-1)."""

    def __init__(self, matrix, K=None, N=None):
        if K is None:
            K, N = matrix.shape
        h, npos, nneg = C.c_void_p(), C.c_int(), C.c_int()
        _check(lib().tsg_sparse_format_build_i32(_ptr(matrix, np.int32), K, N, C.byref(h), C.byref(npos), C.byref(nneg)), "SparseFormat")
        self.col_start_pos, self.col_start_neg = np.empty(N + 1, np.int32), np.empty(N + 1, np.int32)
        self.row_index_pos, self.row_index_neg = np.empty(npos.value, np.int32), np.empty(nneg.value, np.int32)
        _check(lib().tsg_sparse_format_fetch(h, _ptr(self.col_start_pos), _ptr(self.col_start_neg), _ptr(self.row_index_pos),
                                             _ptr(self.row_index_neg)), "SparseFormat")

    def arrays(self):
        return self.col_start_pos, self.col_start_neg, self.row_index_pos, self.row_index_neg


def sparseGEMM(X, col_start_pos, col_start_neg, row_index_pos, row_index_neg, b, Y, M, N, K):
    """reference SparseGEMM.h:104-119, same argument order"""
    _check(lib().tsg_sparse_gemm_f32(_ptr(X), _ptr(col_start_pos), _ptr(col_start_neg), _ptr(row_index_pos), _ptr(row_index_neg),
                                     _ptr(b), _ptr(Y), M, N, K, 0.0, 0), "sparseGEMM")
    return Y


def sparseGEMM_PReLU(X, col_start_pos, col_start_neg, row_index_pos, row_index_neg, b, Y, M, N, K, a):
    """reference SparseGEMM.h:151-168, same argument order"""
    _check(lib().tsg_sparse_gemm_f32(_ptr(X), _ptr(col_start_pos), _ptr(col_start_neg), _ptr(row_index_pos), _ptr(row_index_neg),
                                     _ptr(b), _ptr(Y), M, N, K, float(a), 1), "sparseGEMM_PReLU")
    return Y


class Bcsr:
    """Owner of a `bcsr_t*`; the caller-frees rule of test/test_bcsr.cpp:48-51 is applied in free()."""

    def __init__(self, handle):
        self.handle = handle
        self._libc = C.CDLL(None)
        self._libc.free.argtypes = [C.c_void_p]

    @property
    def s(self) -> bcsr_t:
        return self.handle.contents

    r = property(lambda self: self.s.r)
    c = property(lambda self: self.s.c)
    br = property(lambda self: self.s.br)
    bc = property(lambda self: self.s.bc)
    k = property(lambda self: self.s.k)
    b_row_start = property(lambda self: np.ctypeslib.as_array(self.s.b_row_start, shape=(self.s.br + 1,)))
    b_col_idx = property(lambda self: np.ctypeslib.as_array(self.s.b_col_idx, shape=(self.s.k,)) if self.s.k else np.empty(0, np.int32))
    b_values = property(lambda self: np.ctypeslib.as_array(self.s.b_values, shape=(self.s.k * self.s.r * self.s.c,))
                        if self.s.k else np.empty(0, np.float32))

    def free(self):
        if self.handle:
            lib().bcsr_release_device(self.handle)
            s = self.s
            for p in (s.b_values, s.b_row_start, s.b_col_idx):
                self._libc.free(C.cast(p, C.c_void_p))
            self._libc.free(C.cast(self.handle, C.c_void_p))
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def bcsr_from_dense(dense, r, c, rows=None, cols=None) -> Bcsr:
    if rows is None:
        rows, cols = dense.shape
    h = lib().bcsr_from_dense(_ptr(dense, np.float32), rows, cols, r, c)
    if not h:
        raise TsgError("bcsr_from_dense returned NULL: " + last_error())
    return Bcsr(h)


def _bgemm(name, X, W: Bcsr, B, a, Y, N):
    M, K = X.shape
    Y = _out(Y, X, M, N)
    fn = getattr(lib(), name)
    if a is None:
        fn(_ptr(X, np.float32), W.s, _ptr(B, np.float32), _ptr(Y, np.float32), M, N, K)
    else:
        fn(_ptr(X, np.float32), W.s, _ptr(B, np.float32), float(a), _ptr(Y, np.float32), M, N, K)
    err = last_error()
    if err:
        raise TsgError(f"{name}: {err}")
    return Y


def bcsr_sgemm_basic(X, W, B, N, Y=None):
    return _bgemm("bcsr_sgemm_basic", X, W, B, None, Y, N)


def bcsr_sgemm_avx(X, W, B, N, Y=None):
    return _bgemm("bcsr_sgemm_avx", X, W, B, None, Y, N)


def bcsr_sgemm_avx2(X, W, B, N, Y=None):
    return _bgemm("bcsr_sgemm_avx2", X, W, B, None, Y, N)


def bcsr_sgemm_prelu_basic(X, W, B, a, N, Y=None):
    return _bgemm("bcsr_sgemm_prelu_basic", X, W, B, a, Y, N)


def bcsr_sgemm_prelu_avx(X, W, B, a, N, Y=None):
    return _bgemm("bcsr_sgemm_prelu_avx", X, W, B, a, Y, N)


def bcsr_set_prelu_literal(on: bool) -> None:
    """bcsr_sgemm_prelu_basic / _avx on this thread: False (default) = PReLU(X*W + B); True = the reference's literal loop
    (activation after every partial update, sparse/bcsr.c:177-218), bit-identical to the reference."""
    _check(lib().tsg_bcsr_set_prelu_literal(1 if on else 0), "tsg_bcsr_set_prelu_literal")


def bcsr_set_kernel(which: int) -> None:
    """0 = default, 1 = plain kernel, 2 = shared-memory ring kernel (include/tsgemm_b200.h); both give the same bits."""
    _check(lib().tsg_bcsr_set_kernel(int(which)), "tsg_bcsr_set_kernel")


# =====================================================================================================================
# device-level API (include/tsgemm_b200.h) -- used by bench.py and the multi-GPU path
# =====================================================================================================================
class DeviceTcsc:
    """tsg_tcsc handle: device-resident TCSC arrays + the private gather stream."""

    def __init__(self, handle: C.c_void_p):
        self.h = handle
        r, c, p, n = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _check(lib().tsg_tcsc_dims(self.h, C.byref(r), C.byref(c), C.byref(p), C.byref(n)), "tsg_tcsc_dims")
        self.rows, self.cols, self.n_pos, self.n_neg = r.value, c.value, p.value, n.value

    @property
    def nnz(self):
        return self.n_pos + self.n_neg

    @classmethod
    def from_dense(cls, dense_dev):
        """dense_dev: torch CUDA tensor (K x N) of float32 (==+-1.0f) or int32 (>=1 / <=-1)."""
        import torch
        K, N = dense_dev.shape
        h = C.c_void_p()
        fn = lib().tsg_tcsc_from_dense_i32 if dense_dev.dtype == torch.int32 else lib().tsg_tcsc_from_dense_f32
        _check(fn(_ptr(dense_dev), K, N, C.byref(h)), "tsg_tcsc_from_dense")
        return cls(h)

    @classmethod
    def from_arrays(cls, csp, csn, rip, rin, rows, cols):
        h = C.c_void_p()
        _check(lib().tsg_tcsc_from_arrays(_ptr(csp), _ptr(csn), _ptr(rip), _ptr(rin), rows, cols, C.byref(h)), "tsg_tcsc_from_arrays")
        return cls(h)

    def download(self):
        csp, csn = np.empty(self.cols + 1, np.int32), np.empty(self.cols + 1, np.int32)
        rip, rin = np.empty(self.n_pos, np.int32), np.empty(self.n_neg, np.int32)
        _check(lib().tsg_tcsc_download(self.h, _ptr(csp), _ptr(csn), _ptr(rip), _ptr(rin)), "tsg_tcsc_download")
        return csp, csn, rip, rin

    def stream_info(self):
        b, kc, nc = C.c_longlong(), C.c_int(), C.c_int()
        _check(lib().tsg_tcsc_stream_info(self.h, C.byref(b), C.byref(kc), C.byref(nc)), "tsg_tcsc_stream_info")
        return {"bytes": b.value, "kc": kc.value, "nchunk": nc.value}

    def gemm(self, X, B, Y, a=0.0, use_prelu=False, order=ORDER_BIAS_LAST, ldy=None):
        M, K = X.shape
        N = self.cols
        _check(lib().tsg_tcsc_gemm(self.h, _ptr(X), _ptr(B), float(a), int(use_prelu), order, _ptr(Y), M, N, K,
                                   N if ldy is None else ldy), "tsg_tcsc_gemm")
        return Y

    def destroy(self):
        if self.h:
            lib().tsg_tcsc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def gen_ternary(K, N, seed, num, den, device="cuda", dtype=None):
    import torch
    dtype = dtype or torch.float32
    W = torch.empty((K, N), dtype=dtype, device=device)
    fn = lib().tsg_gen_ternary_i32 if dtype == torch.int32 else lib().tsg_gen_ternary_f32
    _check(fn(_ptr(W), K * N, seed, num, den), "tsg_gen_ternary")
    return W


def gen_ternary_slice(K, N, col0, ncols, seed, num, den, device="cuda"):
    import torch
    W = torch.empty((K, ncols), dtype=torch.float32, device=device)
    _check(lib().tsg_gen_ternary_slice_f32(_ptr(W), K, N, col0, ncols, seed, num, den), "tsg_gen_ternary_slice")
    return W


def gen_sparse_pattern(H, W, non_zero, uniform, seed, device="cuda", dtype=None):
    """generateSparseMatrix (reference SparseGEMM.h:53-102) on the device: the window pattern (uniform=True) or the
    row-skewed pattern (uniform=False); int32 (what SparseFormat takes) or float32."""
    import torch
    dtype = dtype or torch.int32
    out = torch.empty((H, W), device=device, dtype=dtype)
    fn = lib().tsg_gen_sparse_pattern_i32 if dtype == torch.int32 else lib().tsg_gen_sparse_pattern_f32
    _check(fn(_ptr(out), H, W, int(non_zero), int(bool(uniform)), seed), "tsg_gen_sparse_pattern")
    return out


def gen_uniform(shape, seed, device="cuda"):
    import torch
    X = torch.empty(shape, dtype=torch.float32, device=device)
    _check(lib().tsg_gen_uniform_f32(_ptr(X), X.numel(), seed), "tsg_gen_uniform")
    return X


def gen_intvalued(shape, seed, rng=512, device="cuda"):
    import torch
    X = torch.empty(shape, dtype=torch.float32, device=device)
    _check(lib().tsg_gen_intvalued_f32(_ptr(X), X.numel(), seed, rng), "tsg_gen_intvalued")
    return X


def verify_dense_f64(X, Wdense, B, Y, a=0.0, use_prelu=False, m0=0, mrows=None, ldy=None):
    """max |Y - y64| / max(|y64|, 1) and max |Y - y64| against a double-precision dense evaluation on the device."""
    M, K = X.shape
    N = Wdense.shape[1]
    mrows = M - m0 if mrows is None else mrows
    out = (C.c_double * 2)()
    _check(lib().tsg_verify_dense_f64(_ptr(X), _ptr(Wdense), _ptr(B), float(a), int(use_prelu), _ptr(Y), M, N, K,
                                      N if ldy is None else ldy, m0, mrows, C.byref(out)), "tsg_verify_dense_f64")
    return float(out[0]), float(out[1])


# ---- column-partitioned multi-GPU path ---------------------------------------------------------------------------------
def partition(N: int, rank: int, world: int):
    """Columns [col0, col0+ncols) owned by `rank` (tsg_dist_partition; pure arithmetic, no device needed)."""
    c0, nc = C.c_int(), C.c_int()
    lib().tsg_dist_partition(N, rank, world, C.byref(c0), C.byref(nc))
    return c0.value, nc.value


class Dist:
    """One rank of the column-partitioned path.  Bootstraps its own NCCL communicator: rank 0's 128-byte unique id is
    shipped to the other ranks with torch.distributed (any backend)."""

    def __init__(self, rank: int, world: int, group=None):
        import torch
        import torch.distributed as dist
        self.rank, self.world = rank, world
        uid = (C.c_ubyte * 128)()
        if rank == 0:
            _check(lib().tsg_dist_unique_id(C.byref(uid)), "tsg_dist_unique_id")
        if world > 1:
            obj = [bytes(uid)]
            dist.broadcast_object_list(obj, src=0, group=group)
            uid = (C.c_ubyte * 128).from_buffer_copy(obj[0])
        self.h = C.c_void_p()
        _check(lib().tsg_dist_create(C.byref(uid), rank, world, C.byref(self.h)), "tsg_dist_create")
        self._torch = torch

    def partition(self, N):
        return partition(N, self.rank, self.world)

    def alloc_y(self, M, N):
        """Symmetric Y (fused mode): returns a torch tensor view of this rank's buffer."""
        p = C.c_void_p()
        _check(lib().tsg_dist_alloc_y(self.h, M * N * 4, C.byref(p)), "tsg_dist_alloc_y")
        return _tensor_from_ptr(p.value, (M, N))

    def barrier(self):
        _check(lib().tsg_dist_barrier(self.h), "tsg_dist_barrier")

    def has_multicast(self) -> bool:
        """the symmetric Y of the last alloc_y carries an NVSwitch multicast mapping (mode 5 usable)"""
        return bool(lib().tsg_dist_has_multicast(self.h))

    def gemm(self, W_local: DeviceTcsc, X, B, Y, N, a=0.0, use_prelu=False, order=ORDER_BIAS_LAST, root=0, mode=0):
        M, K = X.shape
        _check(lib().tsg_dist_gemm(self.h, W_local.h, _ptr(X), root, _ptr(B), float(a), int(use_prelu), order, _ptr(Y), M, N, K, mode),
               "tsg_dist_gemm")
        return Y

    def gemm_host(self, W_local: DeviceTcsc, X_host, B, Y_host, N, a=0.0, use_prelu=False, order=ORDER_BIAS_LAST, mode=0):
        """tsg_dist_gemm_host: X_host / Y_host are numpy arrays over memory shared by all ranks; every rank moves its row
        block over its own PCIe link.  Returns when this rank's rows of Y_host are complete."""
        M, K = X_host.shape
        _check(lib().tsg_dist_gemm_host(self.h, W_local.h, _ptr(X_host, np.float32), _ptr(B), float(a), int(use_prelu), order,
                                        _ptr(Y_host, np.float32), M, N, K, mode), "tsg_dist_gemm_host")
        return Y_host

    def destroy(self):
        if self.h:
            lib().tsg_dist_destroy(self.h)
            self.h = None


def host_register(arr) -> None:
    _check(lib().tsg_host_register(_ptr(arr), arr.nbytes), "tsg_host_register")


def host_unregister(arr) -> None:
    _check(lib().tsg_host_unregister(_ptr(arr)), "tsg_host_unregister")


def _tensor_from_ptr(ptr: int, shape):
    """Wrap raw device memory owned by the library as a torch tensor (no copy) via __cuda_array_interface__."""
    import torch

    class _Holder:
        pass

    hld = _Holder()
    hld.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}
    return torch.as_tensor(hld, device="cuda")
