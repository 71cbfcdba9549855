#!/usr/bin/env python
"""Time magat_gso_scan alone (CUDA events) on the bench GSO: python tools/time_scan.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from magat_pathplanning_b200 import _cabi  # noqa: E402

dev = torch.device("cuda:0")
B, N = 512, 1000
gen = torch.Generator(device=dev).manual_seed(1337)
S = bench.synth_gso(B, N, 200, dev, gen)
L = _cabi.lib()
W = (N + 31) // 32
rowbits = torch.empty((B, N, W), dtype=torch.int32, device=dev)
colbits = torch.empty((B, N, W), dtype=torch.int32, device=dev)
stats = torch.zeros(4, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream


def run():
    _cabi.check(L.magat_gso_scan(S.data_ptr(), 0, B, N, rowbits.data_ptr(), colbits.data_ptr(), stats.data_ptr(), st))


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"scan+stats {ms:.3f} ms  -> {S.numel() * 4 / ms / 1e6:.0f} GB/s of GSO", os.environ.get("MAGAT_SCAN_ONE_COPY"))
