"""GPU parity of the two BCSR kernels (gemm_bcsr_ring.cu, the default, and gemm_bcsr.cu): the same FFMA sequence, hence the
same bits as each other and as the oracle (sparse/bcsr.c:141-175).  Kept in its own file, after the TCSC parity tests."""
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


@pytest.mark.parametrize("shape", [(200, 520, 530, 1, 4, 1, 2, 17), (128, 300, 300, 3, 2, 1, 2, 18), (96, 256, 256, 1, 1, 9, 10, 19),
                                   (300, 2048, 1024, 4, 16, 1, 2, 20), (128, 512, 512, 1, 8, 0, 1, 21), (1, 512, 512, 1, 8, 1, 2, 7048),
                                   (129, 448, 264, 1, 8, 1, 2, 7049), (64, 225, 512, 1, 16, 1, 2, 7050), (64, 1000, 40, 5, 2, 1, 3, 7051)])
def test_bcsr_ring_and_plain_kernels_agree(t, port, shape):
    """the shared-memory ring kernel (default) and the plain kernel perform the same FFMA sequence: same bits as each
    other and as the oracle -- r > 1 (block rows become consecutive entries), ragged K/N, a single row, an empty W"""
    M, K, N, r, c, num, den, seed = shape
    Wd = port.gen_ternary(K, N, seed, num, den)
    X, B = port.gen_uniform((M, K), seed + 1), port.gen_uniform((N,), seed + 2)
    want = port.bcsr_sgemm_basic(X, port.bcsr_from_dense(Wd, r, c), B, N)
    try:
        for which in (1, 2):
            t.bcsr_set_kernel(which)
            w = t.bcsr_from_dense(Wd, r, c)
            try:
                assert np.array_equal(t.bcsr_sgemm_basic(X, w, B, N), want), f"kernel {which}"
            finally:
                w.free()
    finally:
        t.bcsr_set_kernel(0)
