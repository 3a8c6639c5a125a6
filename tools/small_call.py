#!/usr/bin/env python
"""Fixed cost of a small host-pointer call (the reference driver's first shape, main.cpp:259: M=1, K=512, N=2048, 50 %)
and of a few neighbours, through the C entry point tcsc_sgemm_prelu_basic with malloc'ed numpy buffers."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

torch.cuda.set_device(0)
t = ge.load()
L = t.lib()
out = {"env": {k: v for k, v in os.environ.items() if k.startswith("TSG_")}, "calls": []}
for (M, K, N) in [(1, 512, 2048), (1, 4096, 4096), (4, 512, 2048), (16, 512, 512)]:
    Wd = t.gen_ternary(K, N, 42, 1, 2).cpu().numpy()
    W = t.tcsc_from_dense(Wd)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((M, K), dtype=np.float32)
    b = rng.standard_normal((N,), dtype=np.float32)
    y = np.empty((M, N), np.float32)
    for _ in range(30):
        t.tcsc_sgemm_prelu_basic(x, W, b, 0.2, Y=y)
    ref = np.where((v := x.astype(np.float64) @ Wd.astype(np.float64) + b) < 0, 0.2 * v, v)
    err = float(np.abs(y - ref).max() / max(1.0, np.abs(ref).max()))
    fn, a = L.tcsc_sgemm_prelu_basic, (x.ctypes.data, W.handle, b.ctypes.data, 0.2, y.ctypes.data, M, N, K)
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        for _ in range(300):
            fn(*a)
        best = min(best, (time.perf_counter() - t0) * 1e6 / 300)
    out["calls"].append({"M": M, "K": K, "N": N, "us_per_call": best, "max_rel_err_vs_f64": err})
    W.free()
print(json.dumps(out))
