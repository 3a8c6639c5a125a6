"""Time the dense->TCSC conversion and the gather-stream build in steady state (CUDA events)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
t = ge.load(); t.lib(); torch.cuda.set_device(0)
K = N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
den = int(sys.argv[2]) if len(sys.argv) > 2 else 10
Wd = t.gen_ternary(K, N, 42, 1, den)
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(3):
    torch.cuda.synchronize()
    e[0].record(); W = t.DeviceTcsc.from_dense(Wd); e[1].record(); info = W.stream_info(); e[2].record(); torch.cuda.synchronize()
    print(f"K=N={K} 1/{den}: dense->TCSC {e[0].elapsed_time(e[1]):.3f} ms, gather-stream build {e[1].elapsed_time(e[2]):.3f} ms, {info}, nnz={W.nnz}", flush=True)
    W.destroy()
