#!/usr/bin/env python
"""Debug helper: one exchange mode, one shape, bit-exact check against a single-GPU recompute (torchrun, 2+ ranks)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

mode, M, K, N = [int(x) for x in sys.argv[1:5]]
rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
t = ge.load()
t.lib()
t.use_torch_stream()
D = t.Dist(rank, world)
c0, nc = D.partition(N)
W = t.DeviceTcsc.from_dense(t.gen_ternary_slice(K, N, c0, nc, 42, 1, 4))
X = t.gen_uniform((M, K), 43) if rank == 0 else torch.zeros((M, K), device="cuda")
B = t.gen_uniform((N,), 44)
Y = D.alloc_y(M, N) if mode >= 1 else torch.empty((M, N), device="cuda")
print(f"rank {rank}: mode {mode} M{M} K{K} N{N} multicast={D.has_multicast()}", flush=True)
for i in range(3):
    D.gemm(W, X, B, Y, N, a=0.2, use_prelu=True, root=0, mode=mode)
    torch.cuda.synchronize()
    print(f"rank {rank}: call {i} done", flush=True)
Wf = t.DeviceTcsc.from_dense(t.gen_ternary(K, N, 42, 1, 4))
Yf = torch.empty((M, N), device="cuda")
Wf.gemm(X, B, Yf, a=0.2, use_prelu=True)
torch.cuda.synchronize()
print(f"rank {rank}: mode {mode} equal={bool(torch.equal(Yf, Y))} X_is_roots={bool(torch.equal(X, t.gen_uniform((M, K), 43)))}", flush=True)
dist.barrier()
D.destroy()
dist.destroy_process_group()
