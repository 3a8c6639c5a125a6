"""NVLink write bandwidth of SM-issued stores into a peer's memory: scattered 64-byte row segments (what the fused
epilogue of dist mode 1 does) vs 1 KB rows through bulk async (TMA) stores.  torchrun --nproc-per-node >= 2."""
import ctypes as C, os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
os.environ.setdefault("NCCL_DEBUG", "WARN")
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
t = ge.load(); L = t.lib(); t.use_torch_stream()
M, N = 4096, 4096 * world
D = t.Dist(rank, world)
Y = D.alloc_y(M, N)
ptrs = (C.c_void_p * 8)()
L.tsg_dist_peer_ptrs.argtypes = [C.c_void_p, C.POINTER(C.c_void_p * 8)]
L.tsg_dist_peer_ptrs(D.h, C.byref(ptrs))
L.tsg_dbg_peer_store.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int]
c0 = 4096 * rank
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for target, name in ((rank, "local"), ((rank + 1) % world, "peer")):
    for mode, mname in ((0, "scattered 64B/lane st.v4"), (1, "1KB rows, bulk async (TMA) store")):
        dst = ptrs[target] + 4 * c0
        iters = 10
        assert L.tsg_dbg_peer_store(dst, N, M, 4096, 1, mode) == 0, t.last_error()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0.record()
        assert L.tsg_dbg_peer_store(dst, N, M, 4096, iters, mode) == 0
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        v = torch.tensor([ms], device="cuda"); dist.all_reduce(v, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"{name:6s} {mname:34s} 64 MiB slab in {float(v):.4f} ms = {64 * 1.048576 / float(v):.0f} GB/s (all ranks writing at once)", flush=True)
torch.cuda.synchronize(); D.destroy(); dist.destroy_process_group()
