#!/usr/bin/env python
"""BASELINE.json configs[2]: sparsity sweep 50/75/90/95/99 % at K=N=4096, M = 1..8192, TCSC vs BCSR (r=1,c=8 and r=c=8),
device-resident, CUDA events, best of a few runs.  Writes CSV to stdout (kept under profiles/).  Also times the dense->TCSC
/ dense->BCSR conversions in steady state (second call) and the private gather-stream build."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402
import ctypes as C  # noqa: E402


def timed(fn, iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--K", type=int, default=4096)
    ap.add_argument("--N", type=int, default=4096)
    ap.add_argument("--Ms", default="1,2,4,8,16,32,64,128,256,512,1024,2048,4096,8192")
    ap.add_argument("--dens", default="2,4,10,20,100")
    ap.add_argument("--bcsr", action="store_true")
    ap.add_argument("--pattern", default="iid", choices=["iid", "window", "skewed"],
                    help="W generator: iid ternary (rands_sparse), or generateSparseMatrix's window / row-skewed pattern (SparseGEMM.h:53-102)")
    a = ap.parse_args()
    t = ge.load()
    L = t.lib()
    torch.cuda.set_device(0)
    K, N = a.K, a.N
    peak = 148 * 128 * 1.965e9
    print("format,sparsity,M,K,N,nnz,ms,gadd_per_s,gflops_equiv,frac_fp32_add_peak,frac_smem_ceiling,alg_GBps,kernel")
    for den in [int(x) for x in a.dens.split(",")]:
        if a.pattern == "iid":
            Wd = t.gen_ternary(K, N, 42, 1, den)
        else:  # den = nonZero of generateSparseMatrix: density 1/den like the iid case
            Wd = t.gen_sparse_pattern(K, N, den, a.pattern == "window", 42, dtype=torch.float32)
        t_conv = timed(lambda: t.DeviceTcsc.from_dense(Wd).destroy(), 3)
        W = t.DeviceTcsc.from_dense(Wd)
        info = W.stream_info()
        B = t.gen_uniform((N,), 44)
        print(f"# sparsity {1 - 1 / den:.2f}: nnz={W.nnz} dense->TCSC {t_conv:.3f} ms (steady state, incl. the sizing sync), gather stream {info}", file=sys.stderr)
        bc = {}
        if a.bcsr:
            for (r, c) in ((1, 8), (8, 8)):
                h = C.c_void_p()
                t_b = timed(lambda: (L.tsg_bcsr_from_dense_f32(t._ptr(Wd), K, N, r, c, C.byref(h)), L.tsg_bcsr_destroy(h)), 2)
                L.tsg_bcsr_from_dense_f32(t._ptr(Wd), K, N, r, c, C.byref(h))
                bc[(r, c)] = C.c_void_p(h.value)
                print(f"# dense->BCSR(r={r},c={c}) {t_b:.3f} ms", file=sys.stderr)
        for M in [int(x) for x in a.Ms.split(",")]:
            X = t.gen_uniform((M, K), 43)
            Y = torch.empty((M, N), device="cuda")
            ms = timed(lambda: W.gemm(X, B, Y, a=0.2, use_prelu=True), 5 if M <= 2048 else 3)
            adds = M * W.nnz
            byts = 4 * M * K + 4 * M * N + 4 * W.nnz + 8 * (N + 1) + 4 * N
            print(f"TCSC,{1 - 1 / den:.2f},{M},{K},{N},{W.nnz},{ms:.5f},{adds / ms / 1e6:.1f},{(2 * adds + M * N) / ms / 1e6:.1f},"
                  f"{adds / ms / 1e-3 / peak:.4f},{adds / ms / 1e-3 / (peak / 4):.4f},{byts / ms / 1e6:.1f},{'skinny' if M < 32 else 'tiled'}")
            for (r, c), h in bc.items():
                msb = timed(lambda: L.tsg_bcsr_gemm(h, t._ptr(X), t._ptr(B), 0.2, 1, t._ptr(Y), M, N, K, N), 3)
                print(f"BCSR_r{r}c{c},{1 - 1 / den:.2f},{M},{K},{N},{W.nnz},{msb:.5f},{adds / msb / 1e6:.1f},{(2 * adds + M * N) / msb / 1e6:.1f},"
                      f"{adds / msb / 1e-3 / peak:.4f},{adds / msb / 1e-3 / (peak / 4):.4f},{byts / msb / 1e6:.1f},bcsr")
            sys.stdout.flush()
        for h in bc.values():
            L.tsg_bcsr_destroy(h)
        W.destroy()


if __name__ == "__main__":
    main()
