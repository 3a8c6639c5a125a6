#!/usr/bin/env python
"""Conversion kernels under ncu (launch list):  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum ... python tools/convert_probe.py K N den"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

K, N, den = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
torch.cuda.set_device(0)
t = ge.load()
t.lib()
t.use_torch_stream()
Wd = t.gen_ternary(K, N, 42, 1, den)
for _ in range(2):
    t.DeviceTcsc.from_dense(Wd).destroy()
torch.cuda.synchronize()
