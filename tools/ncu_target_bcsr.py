#!/usr/bin/env python
"""Short target for ncu: the BCSR ring kernel at 4096^3, 1x8 blocks, 90 % sparsity (a few launches)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

torch.cuda.set_device(0)
t = ge.load()
L = t.lib()
t.use_torch_stream()
M = K = N = 4096
Wd = t.gen_ternary(K, N, 42, 1, 10)
h = C.c_void_p()
t._check(L.tsg_bcsr_from_dense_f32(t._ptr(Wd), K, N, 1, 8, C.byref(h)), "tsg_bcsr_from_dense_f32")
X = t.gen_uniform((M, K), 43)
B = t.gen_uniform((N,), 44)
Y = torch.empty((M, N), device="cuda")
for _ in range(3):
    t._check(L.tsg_bcsr_gemm(h, t._ptr(X), t._ptr(B), 0.2, 1, t._ptr(Y), M, N, K, N), "tsg_bcsr_gemm")
torch.cuda.synchronize()
print("ok")
