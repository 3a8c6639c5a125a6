"""GPU tests of the host-side mirror tables and the import checks (csrc/api_tcsc.c, api_bcsr.c, api_cxx.cpp, handles.cu).

The reference hands out plain structs and raw arrays that callers free(), rebuild at the same address or edit in place
(test/test_bcsr.cpp:48-51 frees the BCSR arrays itself; SparseGEMM.cpp:149-156 passes std::vector storage).  The library
caches a device mirror per struct, so every one of those life cycles must end up multiplying by the CURRENT matrix.  All
comparisons are bit-exact against the oracle."""
import ctypes as C

import numpy as np
import pytest

import __graft_entry__ as ge

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def t():
    import torch
    assert torch.cuda.is_available()
    mod = ge.load()
    mod.lib()
    return mod


def _libc():
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    return libc


def test_bcsr_free_and_rebuild_at_same_address(t, port):
    """build, multiply, free() with plain free (no bcsr_release_device), rebuild the same shape: malloc hands the same
    addresses out again; with 4x4 blocks at 50 % every block is kept, so k is equal too -- only the contents differ"""
    L, libc = t.lib(), _libc()
    M, K, N, r, c = 64, 256, 256, 4, 4
    X = port.gen_uniform((M, K), 5)
    B = port.gen_uniform((N,), 6)
    seen, reused = set(), 0
    for it in range(6):
        Wd = port.gen_ternary(K, N, 100 + it, 1, 2)
        h = L.bcsr_from_dense(Wd.ctypes.data, K, N, r, c)
        assert h, t.last_error()
        s = h.contents
        key = C.cast(s.b_values, C.c_void_p).value
        reused += key in seen
        seen.add(key)
        Y = np.zeros((M, N), np.float32)
        L.bcsr_sgemm_basic(X.ctypes.data, s, B.ctypes.data, Y.ctypes.data, M, N, K)
        assert t.last_error() == ""
        wo = port.bcsr_from_dense(Wd, r, c)
        assert np.array_equal(Y, port.bcsr_sgemm_basic(X, wo, B, N)), f"iteration {it}: stale device mirror"
        for p in (s.b_values, s.b_row_start, s.b_col_idx):  # test/test_bcsr.cpp:48-51
            libc.free(C.cast(p, C.c_void_p))
        libc.free(C.cast(h, C.c_void_p))
    print(f"b_values address re-used by malloc in {reused} of 5 rebuilds")


def test_tcsc_rebuild_without_tcsc_free(t, port):
    """a caller that releases a tcsc_t with plain free() (bypassing tcsc_free) and builds another one"""
    L, libc = t.lib(), _libc()
    M, K, N = 40, 128, 96
    X = port.gen_uniform((M, K), 7)
    B = port.gen_uniform((N,), 8)
    for it in range(5):
        Wd = port.gen_ternary(K, N, 200 + it, 1, 4)
        h = L.tcsc_from_dense(Wd.ctypes.data, K, N)
        assert h, t.last_error()
        Y = np.zeros((M, N), np.float32)
        L.tcsc_sgemm_prelu_basic(X.ctypes.data, h, B.ctypes.data, 0.2, Y.ctypes.data, M, N, K)
        assert t.last_error() == ""
        assert np.array_equal(Y, port.tcsc_sgemm_prelu_basic(X, port.tcsc_from_dense(Wd), B, 0.2)), f"iteration {it}"
        s = h.contents
        for p in (s.col_start_pos, s.col_start_neg, s.row_index_pos, s.row_index_neg):
            libc.free(C.cast(p, C.c_void_p))
        libc.free(C.cast(h, C.c_void_p))


def test_tcsc_edited_in_place(t, port):
    """caller-visible arrays are plain memory (main.cpp:296 reads them): an in-place edit must not hit a stale mirror"""
    M, K, N = 48, 200, 64
    X = port.gen_uniform((M, K), 9)
    B = port.gen_uniform((N,), 10)
    Wd = port.gen_ternary(K, N, 300, 1, 4)
    W = t.tcsc_from_dense(Wd)
    Wo = port.tcsc_from_dense(Wd)
    assert np.array_equal(t.tcsc_sgemm_basic(X, W, B), port.tcsc_sgemm_basic(X, Wo, B))
    # move the first +1 of column 0 to another free row, keeping the list ascending
    lo, hi = W.col_start_pos[0], W.col_start_pos[1]
    assert hi - lo >= 2
    rows = W.row_index_pos
    old = int(rows[lo])
    new = next(k for k in range(int(rows[lo + 1])) if k != old and Wd[k, 0] == 0)
    rows[lo] = new
    Wd2 = Wd.copy()
    Wd2[old, 0], Wd2[new, 0] = 0.0, 1.0
    Wo2 = port.tcsc_from_dense(Wd2)
    assert np.array_equal(W.row_index_pos, Wo2.row_index_pos)
    assert np.array_equal(t.tcsc_sgemm_basic(X, W, B), port.tcsc_sgemm_basic(X, Wo2, B))
    # explicit invalidation (for edits the sampled hash cannot see)
    rows[lo] = old
    t.lib().tcsc_invalidate(W.handle)
    assert np.array_equal(t.tcsc_sgemm_basic(X, W, B), port.tcsc_sgemm_basic(X, Wo, B))
    W.free()


def test_raw_arrays_reused_storage(t, port):
    """sparseGEMM's raw arrays (SparseGEMM.h:104-119): same addresses and column counts, different interior rows"""
    M, K, N = 36, 64, 32
    X = port.gen_intvalued((M, K), 11, 512)
    b = np.full(N, 2.0, np.float32)
    rng = np.random.default_rng(0)
    csp = np.arange(0, 4 * (N + 1), 4, dtype=np.int32)  # 4 entries +1 and 4 entries -1 per column
    csn = csp.copy()
    rip, rin = np.empty(4 * N, np.int32), np.empty(4 * N, np.int32)
    Y = np.empty((M, N), np.float32)
    for it in range(4):
        Wd = np.zeros((K, N), np.float32)
        for n in range(N):
            ks = rng.choice(K, 8, replace=False)
            p, q = np.sort(ks[:4]), np.sort(ks[4:])
            rip[4 * n:4 * n + 4], rin[4 * n:4 * n + 4] = p, q  # in place: the pointers never change
            Wd[p, n], Wd[q, n] = 1.0, -1.0
        t.sparseGEMM(X, csp, csn, rip, rin, b, Y, M, N, K)
        assert np.array_equal(Y, X @ Wd + b), f"iteration {it}: stale raw-array mirror"  # integer-valued: exact


def test_import_rejects_bad_arrays(t):
    K, N = 16, 4
    csp = np.array([0, 2, 4, 6, 8], np.int32)
    csn = np.zeros(N + 1, np.int32)
    rin = np.zeros(0, np.int32)
    good = np.array([0, 5, 1, 2, 3, 9, 4, 15], np.int32)
    h = t.DeviceTcsc.from_arrays(csp, csn, good, rin, K, N)
    h.destroy()
    for bad, what in ((np.array([5, 0, 1, 2, 3, 9, 4, 15], np.int32), "ascending"), (np.array([0, 5, 1, 2, 3, 9, 4, 16], np.int32), "outside"),
                      (np.array([0, -1, 1, 2, 3, 9, 4, 15], np.int32), "outside")):
        with pytest.raises(t.TsgError) as e:
            t.DeviceTcsc.from_arrays(csp, csn, bad, rin, K, N)
        assert what in str(e.value)
    with pytest.raises(t.TsgError) as e:
        t.DeviceTcsc.from_arrays(np.array([0, 4, 2, 6, 8], np.int32), csn, good, rin, K, N)
    assert "monotone" in str(e.value)


def test_pinned_x_device_y_returns_after_the_copy(t, port):
    """host X (pinned) + device Y: the call must not return before X has been read (the caller may overwrite it)"""
    import torch
    M, K, N = 96, 4096, 128
    Wd = port.gen_ternary(K, N, 400, 1, 10)
    W = t.tcsc_from_dense(Wd)
    Wo = port.tcsc_from_dense(Wd)
    Xp = torch.from_numpy(port.gen_uniform((M, K), 12)).pin_memory()
    Xref = Xp.numpy().copy()
    Bd = torch.from_numpy(port.gen_uniform((N,), 13)).cuda()
    Yd = torch.empty((M, N), dtype=torch.float32, device="cuda")
    t.tcsc_sgemm_basic(Xp, W, Bd, Yd)
    Xp.zero_()  # allowed as soon as the call has returned
    torch.cuda.synchronize()
    assert np.array_equal(Yd.cpu().numpy(), port.tcsc_sgemm_basic(Xref, Wo, Bd.cpu().numpy()))
    W.free()
