#!/usr/bin/env python
"""How long does the X broadcast of tsg_dist_gemm take at a given size?  (torchrun, one rank per GPU.)  Times mode 5 with
root=0 and with root=-1 on the same operands (W without non-zeros, few columns: the difference is the broadcast) and
NCCL's broadcast of the same buffer beside it.  Environment switches of dist.cu apply (TSG_DIST_NCCL_BCAST,
TSG_MC_BCAST_UNROLL, TSG_MC_BCAST_CTAS)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
t = ge.load()
t.lib()
t.use_torch_stream()
D = t.Dist(rank, world)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    v = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    return float(v)


out = {"world": world, "env": {k: v for k, v in os.environ.items() if k.startswith("TSG_")}, "sizes": []}
N = 256 * world
c0, nc = D.partition(N)
for (M, K) in [(4096, 4096), (8192, 8192), (16384, 16384)]:
    W0 = t.DeviceTcsc.from_dense(torch.zeros((K, nc), device="cuda"))
    X = t.gen_uniform((M, K), 43)
    B = t.gen_uniform((N,), 44)
    Y = D.alloc_y(M, N)
    mode = 5 if D.has_multicast() else 3
    with_b = timeit(lambda: D.gemm(W0, X, B, Y, N, a=0.2, use_prelu=True, root=0, mode=mode))
    without = timeit(lambda: D.gemm(W0, X, B, Y, N, a=0.2, use_prelu=True, root=-1, mode=mode))
    nccl = timeit(lambda: dist.broadcast(X, src=0))
    mb = M * K * 4 / 2**20
    out["sizes"].append({"X_MiB": mb, "mode": mode, "gemm_with_bcast_ms": with_b, "gemm_without_ms": without, "bcast_ms": with_b - without,
                         "bcast_GBps": M * K * 4 / (with_b - without) / 1e6, "nccl_bcast_ms": nccl, "nccl_GBps": M * K * 4 / nccl / 1e6})
    del W0, X
if rank == 0:
    print(json.dumps(out))
torch.cuda.synchronize()
D.destroy()
dist.destroy_process_group()
