import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
t = ge.load(); t.lib(); torch.cuda.set_device(0); t.use_torch_stream()
M = K = N = 16384
Wd = t.gen_ternary(K, N, 42, 1, 10); X = t.gen_uniform((M, K), 43); B = t.gen_uniform((N,), 44); Y = torch.empty((M, N), device="cuda")
W = t.DeviceTcsc.from_dense(Wd); W.gemm(X, B, Y, a=0.2, use_prelu=True); torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
for it in range(4):
    torch.cuda.synchronize(); h0 = time.perf_counter()
    ev[0].record(); W.destroy(); ev[1].record(); W = t.DeviceTcsc.from_dense(Wd); ev[2].record(); h1 = time.perf_counter()
    W.gemm(X, B, Y, a=0.2, use_prelu=True); ev[3].record(); h2 = time.perf_counter(); torch.cuda.synchronize()
    print(f"destroy {ev[0].elapsed_time(ev[1]):.3f} ms | from_dense {ev[1].elapsed_time(ev[2]):.3f} ms (host {1e3*(h1-h0):.2f}) | gemm incl. stream build {ev[2].elapsed_time(ev[3]):.3f} ms (host {1e3*(h2-h1):.2f})", flush=True)
