import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
os.environ.setdefault("NCCL_DEBUG", "WARN")
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
t = ge.load(); t.lib(); t.use_torch_stream()
M = K = 4096; N = 4096 * world
D = t.Dist(rank, world); c0, nc = D.partition(N)
W = t.DeviceTcsc.from_dense(t.gen_ternary_slice(K, N, c0, nc, 42, 1, 10))
X = t.gen_uniform((M, K), 43); B = t.gen_uniform((N,), 44); Y = D.alloc_y(M, N)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for ncs in (os.environ.get("SWEEP_NCS", "1,2,3,7").split(",")):
  os.environ["TSG_DIST_COPY_STREAMS"] = ncs
  for i in range(6):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0.record()
    h0 = time.perf_counter()
    D.gemm(W, X, B, Y, N, a=0.2, use_prelu=True, root=-1, mode=2)
    h1 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    if i >= 4 and rank == 0:
        print(f"[single call rank {rank} copy_streams={ncs}] device {e0.elapsed_time(e1):.3f} ms, host enqueue {1e3 * (h1 - h0):.3f} ms", flush=True)
# back-to-back
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0.record(); h0 = time.perf_counter()
for i in range(20):
    D.gemm(W, X, B, Y, N, a=0.2, use_prelu=True, root=-1, mode=2)
h1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
if rank == 0:
    print(f"[back-to-back x20 rank 0] device {e0.elapsed_time(e1) / 20:.3f} ms/call, host enqueue {1e3 * (h1 - h0) / 20:.3f} ms/call", flush=True)
D.destroy(); dist.destroy_process_group()
