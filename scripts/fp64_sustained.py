#!/usr/bin/env python
"""FP64 FMA rate of parrm_fp64_fma_burn as a function of how long the kernel runs."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyparrm_b200 import _native  # noqa: E402

sink = torch.zeros(8, dtype=torch.float64, device="cuda")
flops = ctypes.c_double(0.0)
stream = torch.cuda.current_stream()
for shift in (13, 15, 17, 19, 20):
    iters = 1 << shift
    _native.lib.parrm_fp64_fma_burn(iters, sink.data_ptr(), ctypes.byref(flops), stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _native.lib.parrm_fp64_fma_burn(iters, sink.data_ptr(), ctypes.byref(flops), stream.cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"iters 2^{shift}: {ms:8.2f} ms  {flops.value / ms / 1e9:.2f} TFLOP/s")
# back-to-back short kernels for ~0.5 s
iters = 1 << 15
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 120
for _ in range(n):
    _native.lib.parrm_fp64_fma_burn(iters, sink.data_ptr(), ctypes.byref(flops), stream.cuda_stream)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"{n} back-to-back 2^15 kernels: {ms:.1f} ms total  {n * flops.value / ms / 1e9:.2f} TFLOP/s")
