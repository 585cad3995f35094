#!/usr/bin/env python
"""Experiment: the NCHW sampling kernels gathering straight from PINNED HOST feature maps (unified addressing: the kernel
reads the taps over PCIe) against cudaMemcpyAsync of the whole maps followed by the device-side gather.
B = 256, 67 points, 256 channels, the three configs[1] levels."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from whmr_b200 import _lib, ops  # noqa: E402
from whmr_b200._lib import check  # noqa: E402

dev = torch.device("cuda:0")
B, C, N = 256, 256, int(os.environ.get("N", 67))
g = torch.Generator().manual_seed(0)
pts = (torch.rand(B, N, 2, generator=g) * 1.9 - 0.95).to(dev)
lib = _lib.lib()


def sample(ptr, H, W, out, layout=0):
    check(lib.whmr_sample_bilinear(ptr, layout, B, C, H, W, pts.data_ptr(), 0, N, out.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream))


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for H, W in ((32, 24), (64, 48), (128, 96)):
    h = torch.empty(B, C, H, W, pin_memory=True)
    h.normal_(generator=g)
    d = torch.empty(B, C, H, W, device=dev)
    out_d = torch.empty(B, C, N, device=dev)
    out_h = torch.empty(B, C, N, device=dev)
    t_copy = timed(lambda: d.copy_(h, non_blocking=True), 5)
    t_dev = timed(lambda: sample(d.data_ptr(), H, W, out_d), 20)
    t_zero = timed(lambda: sample(h.data_ptr(), H, W, out_h), 5)
    same = bool(torch.equal(out_d, out_h))
    print("%3dx%-3d map %.0f MB: H2D copy %.2f ms (%.1f GB/s) + device gather %.3f ms | zero-copy gather %.2f ms (%.1f GB/s of "
          "whole-map bytes) | identical %s" % (H, W, h.numel() * 4 / 1e6, t_copy, h.numel() * 4 / t_copy / 1e6, t_dev, t_zero,
                                            h.numel() * 4 / t_zero / 1e6, same), flush=True)
