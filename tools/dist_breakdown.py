#!/usr/bin/env python
"""Where does a multi-GPU step go?  Times the pieces of tsg_dist_gemm separately (torchrun, one rank per GPU)."""
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
M = K = 4096
Ng = 4096
N = Ng * world
D = t.Dist(rank, world)
c0, nc = D.partition(N)
W = t.DeviceTcsc.from_dense(t.gen_ternary_slice(K, N, c0, nc, 42, 1, 10))
X = t.gen_uniform((M, K), 43)
B = t.gen_uniform((N,), 44)
Y = D.alloc_y(M, N)


def timeit(fn, iters=20):
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


res = {}
res["torch_bcast_X_64MB"] = timeit(lambda: dist.broadcast(X, src=0))
res["tsg_barrier"] = timeit(lambda: D.barrier())
Bl = B[c0:c0 + nc].contiguous()
res["local_gemm_only(world=1 path)"] = timeit(lambda: t.lib().tsg_tcsc_gemm(W.h, t._ptr(X), t._ptr(Bl), 0.2, 1, 1, Y.data_ptr() + 4 * c0, M, nc, K, N))
modes = [int(m) for m in os.environ.get("TSG_BREAKDOWN_MODES", "5,3,2").split(",")]
print_mc = D.has_multicast()
for mode in modes:
    if mode == 5 and not print_mc:
        continue
    Yb = Y if mode else torch.empty((M, N), device="cuda")
    res[f"dist_gemm_mode{mode}_nobcast"] = timeit(lambda: D.gemm(W, X, B, Yb, N, a=0.2, use_prelu=True, root=-1, mode=mode))
    res[f"dist_gemm_mode{mode}_bcast"] = timeit(lambda: D.gemm(W, X, B, Yb, N, a=0.2, use_prelu=True, root=0, mode=mode))
# exchange only: a W without non-zeros leaves the epilogue (bias + PReLU + store + exchange) and nothing else
W0 = t.DeviceTcsc.from_dense(torch.zeros((K, nc), device="cuda"))
for mode in modes:
    if mode == 0 or (mode == 5 and not print_mc):
        continue
    res[f"exchange_only_mode{mode}(empty W)"] = timeit(lambda: D.gemm(W0, X, B, Y, N, a=0.2, use_prelu=True, root=-1, mode=mode))
res["local_gemm_empty_W"] = timeit(lambda: t.lib().tsg_tcsc_gemm(W0.h, t._ptr(X), t._ptr(Bl), 0.2, 1, 1, Y.data_ptr() + 4 * c0, M, nc, K, N))
# raw peer pushes, no compute: every rank pushes its slab to every peer, (a) strided 2-D copies into the row-major Y,
# (b) the same bytes as contiguous 1-D copies, each on one stream per peer
import ctypes as C
cudart = C.CDLL("libcudart.so.12")
cudart.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
cudart.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
peer_ptrs = (C.c_void_p * 8)()
t.lib().tsg_dist_peer_ptrs.argtypes = [C.c_void_p, C.POINTER(C.c_void_p * 8)]
t.lib().tsg_dist_peer_ptrs(D.h, C.byref(peer_ptrs))
streams = [torch.cuda.Stream() for _ in range(world)]


def push(two_d, rows=M):
    cur = torch.cuda.current_stream()
    evs = []
    for p in range(1, world):
        q = (rank + p) % world
        s = streams[q]
        s.wait_stream(cur)
        src = Y.data_ptr() + 4 * c0
        if two_d:
            cudart.cudaMemcpy2DAsync(peer_ptrs[q] + 4 * c0, N * 4, src, N * 4, nc * 4, rows, 3, C.c_void_p(s.cuda_stream))
        else:  # contiguous: same byte count, lands in the first bytes of the peer buffer region of this rank
            cudart.cudaMemcpyAsync(peer_ptrs[q] + 4 * M * nc * rank, Y.data_ptr(), nc * 4 * rows, 3, C.c_void_p(s.cuda_stream))
        cur.wait_stream(s)


res["push_all_peers_2D_strided"] = timeit(lambda: push(True), 10)
res["push_all_peers_1D_contiguous"] = timeit(lambda: push(False), 10)
# raw peer push of the whole slab with 2-D DMA copies (no overlap)
peer = (rank + 1) % world
Y2 = torch.empty((M, N), device="cuda")
res["memcpy2d_local_slab_64MB"] = timeit(lambda: Y2[:, c0:c0 + nc].copy_(Y[:, c0:c0 + nc]))
if rank == 0:
    for k, v in res.items():
        print(f"{k:40s} {v:8.4f} ms")
torch.cuda.synchronize()
D.destroy()
dist.destroy_process_group()
