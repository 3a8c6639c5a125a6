#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libref_oracle.so).

Run in the build container only (needs /root/reference to have been compiled by `make -C oracle ref`):

    python tests/golden/make_golden.py

The reference has no seeds and no fixtures of its own (all RNG is std::random_device / time(0): dense/utils.h:11-12,
SparseGEMM.h:45,71), so the inputs here are (a) the one fixed matrix of test/test.c:6-11, (b) hand-written edge
cases, (c) tensors from the oracle's counter-based generators, regenerated from the recorded seeds at test time
(a sample of each generator's output is frozen too, so a generator change cannot silently re-define the inputs).
Outputs are whatever the reference returned.  Files are small (< 1 MB in total) and committed.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Port, Ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# (name, M, K, N, num, den, seedW)   -- sparsity = 1 - num/den
TCSC_CASES = [
    ("cfg1_like", 8, 512, 512, 1, 2, 42),     # BASELINE.json configs[0] shape at M=8 (50 %)
    ("s90", 8, 256, 192, 1, 10, 1042),
    ("s99", 5, 300, 130, 1, 100, 2042),       # many empty columns
    ("ragged", 3, 37, 29, 1, 3, 3042),        # nothing divisible by anything
    ("one_col", 4, 64, 1, 1, 2, 4042),
    ("one_row", 2, 1, 40, 1, 2, 5042),
    ("dense_ish", 6, 96, 80, 9, 10, 6042),    # 10 % sparsity
]

BCSR_CASES = [
    ("r1c8", 4, 64, 128, 1, 8, 1, 2, 7042),   # the only blocking the reference tests (test_bcsr.cpp:16-17)
    ("r2c2", 3, 32, 32, 2, 2, 1, 2, 7043),
    ("r8c8", 5, 64, 64, 8, 8, 1, 4, 7044),
    ("r4c4", 2, 64, 96, 4, 4, 1, 2, 7045),
    ("r1c8_s90", 3, 128, 256, 1, 8, 1, 10, 7046),  # sparse enough for dropped blocks, every block-row non-empty
]


def main():
    port, ref, ref_fm = Port(), Ref(""), Ref("fm")
    g = {}

    # ---- (1) the fixed 4x4 matrix of test/test.c:6-11 -------------------------------------------------------------
    m44 = np.array([[-1, -1, 0, -1], [0, -1, 0, 0], [0, 0, -1, -1], [0, 0, -1, 0]], np.float32)
    w = ref.tcsc_from_dense(m44)
    g["kat_test_c.dense"] = m44
    for nm, arr in zip(("csp", "csn", "rip", "rin"), w.arrays()):
        g[f"kat_test_c.tcsc.{nm}"] = arr
    b = ref.bcsr_from_dense(m44, 2, 2)
    g["kat_test_c.bcsr22.row_start"], g["kat_test_c.bcsr22.col_idx"] = b.b_row_start, b.b_col_idx
    g["kat_test_c.bcsr22.values"], g["kat_test_c.bcsr22.k"] = b.b_values, np.int32(b.k)

    # ---- (2) non-ternary values are ignored by the float builder, +-2 are kept by the int builder ----------------
    odd = np.array([[1, .5, -1], [-1, 1, 2], [0, -1, 1], [1, 1, -0.0]], np.float32)
    w = ref.tcsc_from_dense(odd)
    g["kat_odd.dense"] = odd
    for nm, arr in zip(("csp", "csn", "rip", "rin"), w.arrays()):
        g[f"kat_odd.tcsc.{nm}"] = arr
    special = np.array([[np.nan, np.inf, -np.inf, 1], [-1, 1e-45, -1.0000001, 0.99999994], [1, -1, 3, -3]], np.float32)
    w = ref.tcsc_from_dense(special)
    g["kat_special.dense"] = special
    for nm, arr in zip(("csp", "csn", "rip", "rin"), w.arrays()):
        g[f"kat_special.tcsc.{nm}"] = arr
    oddi = np.array([[1, 0, -1, 2], [-2, 1, 5, 0], [0, -1, 1, -7], [1, 1, 0, 0]], np.int32)
    w = ref.sparse_format(oddi)
    g["kat_oddi.dense"] = oddi
    for nm, arr in zip(("csp", "csn", "rip", "rin"), w.arrays()):
        g[f"kat_oddi.tcsc.{nm}"] = arr
    # all-zero and single-sign matrices (malloc(0) paths, tcsc.c:32-33)
    for nm, mat in (("zeros", np.zeros((5, 7), np.float32)), ("allpos", np.ones((3, 4), np.float32)),
                    ("allneg", -np.ones((4, 3), np.float32))):
        w = ref.tcsc_from_dense(mat)
        g[f"kat_{nm}.dense"] = mat
        for an, arr in zip(("csp", "csn", "rip", "rin"), w.arrays()):
            g[f"kat_{nm}.tcsc.{an}"] = arr

    # ---- (3) BCSR empty-block-row quirk (bcsr.c:114-117): only the defined prefix is frozen -------------------------
    q = np.zeros((4, 4), np.float32)
    q[0, 0] = 1
    q[3, 2] = -1
    bq = ref.bcsr_from_dense(q, 1, 2)
    g["kat_quirk.dense"] = q
    g["kat_quirk.row_start_defined_prefix"] = bq.b_row_start[:3]  # [0, 1, k]; entries 3..4 are uninitialised memory
    g["kat_quirk.col_idx"], g["kat_quirk.values"], g["kat_quirk.k"] = bq.b_col_idx, bq.b_values, np.int32(bq.k)

    # ---- generator samples -----------------------------------------------------------------------------------------
    g["gen.ternary_1_10.seed42"] = port.gen_ternary(16, 16, 42, 1, 10)
    g["gen.ternary_i32_1_2.seed7"] = port.gen_ternary(8, 8, 7, 1, 2, np.int32)
    g["gen.uniform.seed43"] = port.gen_uniform((4, 16), 43)
    g["gen.intvalued.seed43"] = port.gen_intvalued((4, 16), 43, 512)

    # ---- (4)(5) seeded TCSC cases: index arrays + every kernel variant ----------------------------------------------
    for name, M, K, N, num, den, seed in TCSC_CASES:
        Wd = port.gen_ternary(K, N, seed, num, den)
        w = ref.tcsc_from_dense(Wd)
        wfm = ref_fm.tcsc_from_dense(Wd)
        for an, arr in zip(("csp", "csn", "rip", "rin"), w.arrays()):
            g[f"tcsc.{name}.{an}"] = arr
        wi = ref.sparse_format(port.gen_ternary(K, N, seed, num, den, np.int32))
        for an, arr, arr_f in zip(("csp", "csn", "rip", "rin"), wi.arrays(), w.arrays()):
            assert np.array_equal(arr, arr_f)
        Bu = port.gen_uniform((N,), seed + 2)
        # integer-valued X, b=2, a=0.25 (SparseGEMM.cpp:81,95,99): exact, order-independent
        Xi = port.gen_intvalued((M, K), seed + 1, 512)
        B2 = np.full(N, 2.0, np.float32)
        g[f"tcsc.{name}.int.Y_bias"] = ref.tcsc_sgemm_basic(Xi, w, B2)
        g[f"tcsc.{name}.int.Y_prelu"] = ref.tcsc_sgemm_prelu_basic(Xi, w, B2, 0.25)
        assert np.array_equal(g[f"tcsc.{name}.int.Y_bias"], ref.tcsc_sgemm_optimized(Xi, w, B2))
        assert np.array_equal(g[f"tcsc.{name}.int.Y_prelu"], ref.tcsc_sgemm_prelu_optimized_onthego(Xi, w, B2, 0.25))
        assert np.array_equal(g[f"tcsc.{name}.int.Y_prelu"], ref.tcsc_sgemm_prelu_optimized_separate(Xi, w, B2, 0.25))
        assert np.array_equal(g[f"tcsc.{name}.int.Y_bias"], ref.sparse_gemm(Xi, w, B2))
        assert np.array_equal(g[f"tcsc.{name}.int.Y_prelu"], ref.sparse_gemm_prelu(Xi, w, B2, 0.25))
        # real-valued X ~ U[-1,1), b ~ U[-1,1), a = 0.2 (main.cpp:268,279-280): the reference's own roundings
        Xu = port.gen_uniform((M, K), seed + 1)
        g[f"tcsc.{name}.real.basic"] = ref.tcsc_sgemm_basic(Xu, w, Bu)
        g[f"tcsc.{name}.real.optimized"] = ref.tcsc_sgemm_optimized(Xu, w, Bu)
        g[f"tcsc.{name}.real.prelu_basic"] = ref.tcsc_sgemm_prelu_basic(Xu, w, Bu, 0.2)
        g[f"tcsc.{name}.real.prelu_separate"] = ref.tcsc_sgemm_prelu_optimized_separate(Xu, w, Bu, 0.2)
        g[f"tcsc.{name}.real.prelu_onthego"] = ref.tcsc_sgemm_prelu_optimized_onthego(Xu, w, Bu, 0.2)
        g[f"tcsc.{name}.real.sparse_gemm"] = ref.sparse_gemm(Xu, w, Bu)
        g[f"tcsc.{name}.real.sparse_gemm_prelu"] = ref.sparse_gemm_prelu(Xu, w, Bu, 0.2)
        g[f"tcsc.{name}.real.prelu_basic_fastmath"] = ref_fm.tcsc_sgemm_prelu_basic(Xu, wfm, Bu, 0.2)
        g[f"tcsc.{name}.real.gemm_basic"] = ref.gemm_basic(Xu, Wd, Bu)

    # ---- BCSR cases --------------------------------------------------------------------------------------------------
    for name, M, K, N, r, c, num, den, seed in BCSR_CASES:
        Wd = port.gen_ternary(K, N, seed, num, den)
        bw = ref.bcsr_from_dense(Wd, r, c)
        # the reference's row pointers are only well defined when every block-row owns a block
        std = port.bcsr_from_dense(Wd, r, c)
        assert np.all(np.diff(std.b_row_start) > 0), f"{name}: pick a seed without empty block-rows"
        g[f"bcsr.{name}.row_start"], g[f"bcsr.{name}.col_idx"] = bw.b_row_start, bw.b_col_idx
        g[f"bcsr.{name}.values"], g[f"bcsr.{name}.k"] = bw.b_values, np.int32(bw.k)
        Xi = port.gen_intvalued((M, K), seed + 1, 512)
        B2 = np.full(N, 2.0, np.float32)
        Xu = port.gen_uniform((M, K), seed + 1)
        Bu = port.gen_uniform((N,), seed + 2)
        g[f"bcsr.{name}.int.Y_bias"] = ref.bcsr_sgemm_basic(Xi, bw, B2, N)
        g[f"bcsr.{name}.real.basic"] = ref.bcsr_sgemm_basic(Xu, bw, Bu, N)
        g[f"bcsr.{name}.real.prelu_basic_literal"] = ref.bcsr_sgemm_prelu_basic(Xu, bw, Bu, 0.2, N)
        if c == 8:
            g[f"bcsr.{name}.real.avx"] = ref.bcsr_sgemm_avx(Xu, bw, Bu, N)
            g[f"bcsr.{name}.real.prelu_avx_literal"] = ref.bcsr_sgemm_prelu_avx(Xu, bw, Bu, 0.2, N)
            assert np.array_equal(g[f"bcsr.{name}.int.Y_bias"], ref.bcsr_sgemm_avx(Xi, bw, B2, N))
        if c == 8 and r == 8:
            g[f"bcsr.{name}.real.avx2"] = ref.bcsr_sgemm_avx2(Xu, bw, Bu, N)

    path = os.path.join(OUT, "reference_vectors.npz")
    np.savez_compressed(path, **g)
    print(f"wrote {path}: {len(g)} arrays, {os.path.getsize(path)/1024:.1f} KiB; reference flags: {ref.build_flags}")


if __name__ == "__main__":
    main()
