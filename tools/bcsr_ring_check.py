#!/usr/bin/env python
"""A/B of the two BCSR kernels on the GPU box, without torch (numpy + ctypes only, so it starts in seconds):
bitwise comparison ring vs plain vs the CPU oracle on the shapes of tests/test_gpu_bcsr.py, then timings of both
kernels with device-resident operands.  Writes gpurun_out/bcsr_ring_check.json.

    gpurun --timeout 120 -- 'timeout 100 python tools/bcsr_ring_check.py'
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle.pyoracle import Port  # noqa: E402

SHAPES = [(64, 512, 2048, 1, 8, 1, 2, 11), (130, 96, 100, 2, 4, 1, 4, 12), (33, 64, 64, 8, 8, 1, 10, 13), (5, 128, 256, 1, 16, 1, 3, 15),
          (256, 1024, 512, 1, 8, 1, 10, 16), (200, 520, 530, 1, 4, 1, 2, 17), (128, 300, 300, 3, 2, 1, 2, 18), (96, 256, 256, 1, 1, 9, 10, 19),
          (300, 2048, 1024, 4, 16, 1, 2, 20), (128, 512, 512, 1, 8, 0, 1, 21),
          # the golden cases of tests/golden/make_golden.py and the shape of the reference's own test (test_bcsr.cpp:16-17)
          (4, 64, 128, 1, 8, 1, 2, 7042), (3, 32, 32, 2, 2, 1, 2, 7043), (5, 64, 64, 8, 8, 1, 4, 7044), (2, 64, 96, 4, 4, 1, 2, 7045),
          (3, 128, 256, 1, 8, 1, 10, 7046), (32, 1024, 4096, 1, 8, 1, 2, 7047), (1, 512, 512, 1, 8, 1, 2, 7048),
          (129, 448, 264, 1, 8, 1, 2, 7049), (64, 225, 512, 1, 16, 1, 2, 7050), (64, 1000, 40, 5, 2, 1, 3, 7051)]


def main():
    t = ge.load()
    L = t.lib()
    port = Port()
    out = {"parity": [], "timing": []}
    ok_all = True
    for (M, K, N, r, c, num, den, seed) in SHAPES:
        Wd = port.gen_ternary(K, N, seed, num, den)
        X, B = port.gen_uniform((M, K), seed + 1), port.gen_uniform((N,), seed + 2)
        wo = port.bcsr_from_dense(Wd, r, c)
        want = port.bcsr_sgemm_basic(X, wo, B, N)
        res = {}
        for name, which in (("plain", 1), ("ring", 0)):  # 0 = the default, which is the ring kernel
            t.bcsr_set_kernel(which)
            w = t.bcsr_from_dense(Wd, r, c)
            y = t.bcsr_sgemm_basic(X, w, B, N)
            yp = t.bcsr_sgemm_prelu_basic(X, w, B, 0.2, N)
            res[name] = bool(np.array_equal(y, want)) and bool(np.array_equal(yp, np.where(y < 0, np.float32(0.2) * y, y)))
            if not res[name]:
                res[name + "_maxdiff"] = float(np.nanmax(np.abs(y - want)))
            w.free()
        ok_all &= res["plain"] and res["ring"]
        out["parity"].append({"shape": [M, K, N, r, c, f"{num}/{den}"], **res})
        print(out["parity"][-1], flush=True)
    t.bcsr_set_kernel(0)
    if "--parity-only" in sys.argv:
        print("ALL OK" if ok_all else "MISMATCH")
        return 0 if ok_all else 1

    # timings: device-resident operands, generators and kernels through the device-level API
    vp = C.c_void_p
    cudart = C.CDLL("libcudart.so.12")
    for (M, K, N, r, c, num, den) in [(4096, 4096, 4096, 1, 8, 1, 2), (4096, 4096, 4096, 1, 8, 1, 10), (4096, 4096, 4096, 4, 4, 1, 10),
                                      (4096, 4096, 4096, 1, 16, 1, 2)]:
        bufs = {}
        for name, n in (("W", K * N), ("X", M * K), ("B", N), ("Y", M * N)):
            p = vp()
            assert L.tsg_dev_alloc(C.byref(p), C.c_size_t(n * 4)) == 0, t.last_error()
            bufs[name] = p
        L.tsg_gen_ternary_f32(bufs["W"], C.c_longlong(K * N), C.c_uint64(1), num, den)
        L.tsg_gen_uniform_f32(bufs["X"], C.c_longlong(M * K), C.c_uint64(2))
        L.tsg_gen_uniform_f32(bufs["B"], C.c_longlong(N), C.c_uint64(3))
        h = vp()
        assert L.tsg_bcsr_from_dense_f32(bufs["W"], K, N, r, c, C.byref(h)) == 0, t.last_error()
        k = C.c_int()
        L.tsg_bcsr_dims(h, None, None, None, None, C.byref(k))
        row = {"shape": [M, K, N, r, c, f"{num}/{den}"], "blocks": k.value}
        ys = {}
        for name, which in (("plain", 1), ("ring", 2)):
            t.bcsr_set_kernel(which)
            for _ in range(2):
                assert L.tsg_bcsr_gemm(h, bufs["X"], bufs["B"], 0.2, 1, bufs["Y"], M, N, K, C.c_longlong(N)) == 0, t.last_error()
            L.tsg_synchronize()
            t0 = time.perf_counter()
            reps = 5
            for _ in range(reps):
                L.tsg_bcsr_gemm(h, bufs["X"], bufs["B"], 0.2, 1, bufs["Y"], M, N, K, C.c_longlong(N))
            L.tsg_synchronize()
            row[name + "_ms"] = (time.perf_counter() - t0) / reps * 1e3
            y = np.empty((64, N), np.float32)  # first rows are enough for an equality check
            cudart.cudaMemcpy(y.ctypes.data_as(vp), bufs["Y"], C.c_size_t(y.nbytes), 2)
            ys[name] = y
        row["same_bits"] = bool(np.array_equal(ys["plain"], ys["ring"]))
        ok_all &= row["same_bits"]
        flop = 2.0 * M * k.value * r * c
        row["ring_tflops"] = flop / row["ring_ms"] / 1e9
        out["timing"].append(row)
        print(row, flush=True)
        L.tsg_bcsr_destroy(h)
        for p in bufs.values():
            L.tsg_dev_free(p)
        t.bcsr_set_kernel(0)
    out["ok"] = bool(ok_all)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bcsr_ring_check.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("ALL OK" if ok_all else "MISMATCH")
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
