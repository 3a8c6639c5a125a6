#!/usr/bin/env python
"""Stage-by-stage bring-up with flushed prints (run under `timeout`): shows where a hang or mismatch happens."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402
from oracle.pyoracle import Port  # noqa: E402


def say(*a):
    print(*a, flush=True)


t = ge.load()
t.lib()
port = Port()
torch.cuda.set_device(0)
say("device ok", t.lib().tsg_device_check(), torch.cuda.get_device_name(0))
M, K, N = [int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (64, 512, 512))]
den = int(sys.argv[4]) if len(sys.argv) > 4 else 2
Wd = port.gen_ternary(K, N, 42, 1, den)
X = port.gen_uniform((M, K), 43)
B = port.gen_uniform((N,), 44)
Wdev = torch.from_numpy(Wd).cuda()
say("gen ok")
W = t.DeviceTcsc.from_dense(Wdev)
torch.cuda.synchronize()
say("convert ok", W.n_pos, W.n_neg)
wo = port.tcsc_from_dense(Wd)
for a, b in zip(W.download(), wo.arrays()):
    assert np.array_equal(a, b)
say("convert matches oracle")
say("stream", W.stream_info())
Xd, Bd = torch.from_numpy(X).cuda(), torch.from_numpy(B).cuda()
Y = torch.zeros((M, N), device="cuda")
for kern in (2, 1):
    t.lib().tsg_tcsc_set_kernel(kern)
    W.gemm(Xd, Bd, Y, a=0.2, use_prelu=True, order=1)
    torch.cuda.synchronize()
    yo = port.tcsc_sgemm_prelu_basic(X, wo, B, 0.2)
    d = np.abs(Y.cpu().numpy() - yo).max()
    say(f"kernel {kern}: gemm done, max |d| vs oracle = {d:.3e}, bit-exact = {np.array_equal(Y.cpu().numpy(), yo)}")
t.lib().tsg_tcsc_set_kernel(0)
w = t.tcsc_from_dense(Wd)
say("host tcsc_from_dense ok")
y = t.tcsc_sgemm_prelu_basic(X, w, B, 0.2)
say("host gemm ok, bit-exact =", np.array_equal(y, port.tcsc_sgemm_prelu_basic(X, wo, B, 0.2)))
