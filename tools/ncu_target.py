#!/usr/bin/env python
"""Short target for ncu: the headline shape (BASELINE.json configs[1]), a few launches of the tiled kernel.
    python tools/ncu_target.py [M K N den [order]]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

a = [int(x) for x in sys.argv[1:]]
M, K, N, den = (a + [4096, 4096, 4096, 10])[:4] if len(a) >= 4 else (4096, 4096, 4096, 10)
order = a[4] if len(a) > 4 else 1
torch.cuda.set_device(0)
t = ge.load()
t.lib()
t.use_torch_stream()
W = t.DeviceTcsc.from_dense(t.gen_ternary(K, N, 42, 1, den))
Xs = [t.gen_uniform((M, K), 43 + i) for i in range(2)]
B = t.gen_uniform((N,), 44)
Y = torch.empty((M, N), device="cuda")
for i in range(4):
    W.gemm(Xs[i % 2], B, Y, a=0.2, use_prelu=True, order=order)
torch.cuda.synchronize()
print("ok")
