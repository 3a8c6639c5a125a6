#!/usr/bin/env python
"""End-to-end time of the host-pointer call tcsc_sgemm_prelu_basic (pinned X and Y in host memory) at a few M, as
bench.py's `e2e` measures it.  TSG_HOST_SLAB_ROWS=256 gives the uniform slab schedule for an A/B against the ramp."""
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
t.lib()
K = N = 4096
Wd = t.gen_ternary(K, N, 42, 1, 10)
Wh = t.tcsc_from_dense(Wd.cpu().numpy())
Wdev = t.DeviceTcsc.from_dense(Wd)
B = t.gen_uniform((N,), 44)
Bh = B.cpu().numpy()
out = {"env": {k: v for k, v in os.environ.items() if k.startswith("TSG_")}, "runs": []}
for M in (4096, 1024, 8192, 5000):
    Xd = t.gen_uniform((M, K), 43)
    Xh = torch.empty((M, K), dtype=torch.float32).pin_memory()
    Xh.copy_(Xd.cpu())
    Yh = torch.empty((M, N), dtype=torch.float32).pin_memory()
    for _ in range(3):
        t.tcsc_sgemm_prelu_basic(Xh.numpy(), Wh, Bh, 0.2, Y=Yh.numpy())
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            t.tcsc_sgemm_prelu_basic(Xh.numpy(), Wh, Bh, 0.2, Y=Yh.numpy())
        best = min(best, (time.perf_counter() - t0) * 1e3 / 10)
    Yd = torch.empty((M, N), device="cuda")
    Wdev.gemm(Xd, B, Yd, a=0.2, use_prelu=True, order=t.ORDER_BIAS_LAST)
    same = bool(torch.equal(Yh, Yd.cpu()))
    out["runs"].append({"M": M, "ms": best, "bit_identical_to_device_resident": same, "pcie_bytes": 4 * (M * K + M * N)})
print(json.dumps(out))
