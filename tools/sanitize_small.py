"""Small end-to-end run for compute-sanitizer (memcheck): conversion, gather-stream build, tiled + skinny TCSC GEMM in all
three orders, BCSR conversion + GEMM, host-pointer entry points."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
from oracle.pyoracle import Port
t = ge.load(); t.lib(); port = Port()
for (M, K, N, den) in ((70, 300, 200, 4), (5, 257, 129, 10), (129, 64, 513, 2)):
    Wd = port.gen_ternary(K, N, 42, 1, den)
    X, B = port.gen_uniform((M, K), 43), port.gen_uniform((N,), 44)
    w, wo = t.tcsc_from_dense(Wd), port.tcsc_from_dense(Wd)
    for a, b in zip(w.arrays(), wo.arrays()):
        assert np.array_equal(a, b)
    ok = np.array_equal(t.tcsc_sgemm_basic(X, w, B), port.tcsc_sgemm_basic(X, wo, B)) if M >= 32 else True
    ok &= np.array_equal(t.tcsc_sgemm_prelu_basic(X, w, B, 0.2), port.tcsc_sgemm_prelu_basic(X, wo, B, 0.2)) if M >= 32 else True
    ok &= np.array_equal(t.tcsc_sgemm_optimized(X, w, B), port.tcsc_sgemm_optimized(X, wo, B)) if M >= 32 else True
    y = t.tcsc_sgemm_prelu_basic(X, w, B, 0.2)
    bw, bo = t.bcsr_from_dense(Wd, 1, 8), port.bcsr_from_dense(Wd, 1, 8)
    ok &= np.array_equal(t.bcsr_sgemm_basic(X, bw, B, N), port.bcsr_sgemm_basic(X, bo, B, N))
    print(M, K, N, den, "ok" if ok else "MISMATCH", flush=True)
    w.free(); bw.free()
print("done")
