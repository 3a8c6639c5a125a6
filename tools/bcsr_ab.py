#!/usr/bin/env python
"""Time tsg_bcsr_gemm (device-resident operands, CUDA events) on a few shapes; run once plain and once with
TSG_BCSR_NO_SPLIT=1 to see what dealing the last round as column slices buys (gemm_bcsr_ring.cu: bcsr_ring_plan)."""
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
for (M, K, N, r, c, keep) in [(4096, 4096, 4096, 1, 8, 10), (4096, 4096, 4096, 1, 8, 2), (4096, 4096, 4096, 8, 8, 10), (256, 4096, 4096, 1, 8, 10),
                              (8192, 4096, 4096, 1, 8, 10), (2048, 4096, 4096, 1, 8, 10), (4096, 4096, 4096, 4, 4, 10)]:
    Wd = t.gen_ternary(K, N, 42, 1, keep)
    h = C.c_void_p()
    t._check(L.tsg_bcsr_from_dense_f32(t._ptr(Wd), K, N, r, c, C.byref(h)), "tsg_bcsr_from_dense_f32")
    X = t.gen_uniform((M, K), 43)
    B = t.gen_uniform((N,), 44)
    Y = torch.empty((M, N), device="cuda")
    call = lambda: t._check(L.tsg_bcsr_gemm(h, t._ptr(X), t._ptr(B), 0.2, 1, t._ptr(Y), M, N, K, N), "tsg_bcsr_gemm")
    for _ in range(5):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        call()
    e1.record()
    torch.cuda.synchronize()
    out.append({"M": M, "K": K, "N": N, "r": r, "c": c, "sparsity": 1 - 1 / keep, "ms": e0.elapsed_time(e1) / 20,
                "checksum": float(Y.double().sum().item())})
    L.tsg_bcsr_destroy(h)
print(json.dumps({"no_split": bool(os.environ.get("TSG_BCSR_NO_SPLIT")), "results": out}))
