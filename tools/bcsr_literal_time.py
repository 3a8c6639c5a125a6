#!/usr/bin/env python
"""Device time of tsg_bcsr_gemm with use_prelu = 1 (PReLU(X*W+b), ring kernel) and use_prelu = 2 (the reference's literal
bcsr_sgemm_prelu_* loop, plain sequential kernel) on the same operands."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

torch.cuda.set_device(0)
t = ge.load()
L = t.lib()
t.use_torch_stream()
out = []
for (M, K, N, r, c, keep) in [(4096, 4096, 4096, 1, 8, 10), (4096, 4096, 4096, 1, 8, 2), (256, 1024, 4096, 1, 8, 2)]:
    Wd = t.gen_ternary(K, N, 42, 1, keep)
    h = C.c_void_p()
    t._check(L.tsg_bcsr_from_dense_f32(t._ptr(Wd), K, N, r, c, C.byref(h)), "tsg_bcsr_from_dense_f32")
    X = t.gen_uniform((M, K), 43)
    B = t.gen_uniform((N,), 44)
    Y = torch.empty((M, N), device="cuda")
    row = {"M": M, "K": K, "N": N, "r": r, "c": c, "sparsity": 1 - 1 / keep}
    for mode, name in ((1, "prelu_math_ms"), (2, "prelu_literal_ms")):
        call = lambda: t._check(L.tsg_bcsr_gemm(h, t._ptr(X), t._ptr(B), 0.2, mode, t._ptr(Y), M, N, K, N), "tsg_bcsr_gemm")
        for _ in range(3):
            call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            call()
        e1.record()
        torch.cuda.synchronize()
        row[name] = e0.elapsed_time(e1) / 10
    out.append(row)
    L.tsg_bcsr_destroy(h)
print(json.dumps(out))
