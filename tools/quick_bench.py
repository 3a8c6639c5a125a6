#!/usr/bin/env python
"""Developer loop: time the device-resident TCSC GEMM at one shape (CUDA events) and print rates.  Not the contract
bench (that is bench.py); used between kernel edits."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=4096)
    ap.add_argument("--K", type=int, default=4096)
    ap.add_argument("--N", type=int, default=4096)
    ap.add_argument("--den", type=int, default=10)
    ap.add_argument("--num", type=int, default=1)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--order", type=int, default=1)
    ap.add_argument("--kernel", type=int, default=0)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    t = ge.load()
    t.lib()
    torch.cuda.set_device(0)
    Wd = t.gen_ternary(a.K, a.N, 42, a.num, a.den)
    X = t.gen_uniform((a.M, a.K), 43)
    B = t.gen_uniform((a.N,), 44)
    Y = torch.empty((a.M, a.N), device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    W = t.DeviceTcsc.from_dense(Wd)
    e1.record()
    torch.cuda.synchronize()
    t_conv = e0.elapsed_time(e1)
    e0.record()
    info = W.stream_info()
    e1.record()
    torch.cuda.synchronize()
    t_stream = e0.elapsed_time(e1)
    t.lib().tsg_tcsc_set_kernel(a.kernel)
    for _ in range(3):
        W.gemm(X, B, Y, a=0.2, use_prelu=True, order=a.order)
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.iters):
        e0.record()
        W.gemm(X, B, Y, a=0.2, use_prelu=True, order=a.order)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    adds = a.M * W.nnz
    flops = 2 * adds + a.M * a.N
    peak = 148 * 128 * 1.965e9
    print(f"M={a.M} K={a.K} N={a.N} density={a.num}/{a.den} nnz={W.nnz} kc={info['kc']} nchunk={info['nchunk']} stream_bytes={info['bytes']}")
    print(f"convert {t_conv:.3f} ms, stream build {t_stream:.3f} ms")
    print(f"gemm best {ms:.4f} ms  median {sorted(ts)[len(ts)//2]:.4f} ms   {adds/ms/1e9:.3f} Tadd/s = {adds/ms/1e-3/peak*100:.1f}% of FP32-add peak "
          f"({adds/ms/1e-3/(peak/4)*100:.1f}% of smem-gather ceiling)   {flops/ms/1e6:.1f} GFLOP/s-equiv")
    if a.check:
        rel, ab = t.verify_dense_f64(X, Wd, B, Y, a=0.2, use_prelu=True, m0=0, mrows=min(a.M, 64))
        print(f"verify rows 0..63 vs fp64 dense: rel {rel:.3e} abs {ab:.3e}")


if __name__ == "__main__":
    main()
