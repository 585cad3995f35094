"""configs[3] sampler timing for the dense-kernel plans (WHMR_DENSE_CG / WHMR_SAMPLE_DENSE set by the caller)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whmr_b200  # noqa: E402,F401
import whmr_b200.synthetic as syn  # noqa: E402
from whmr_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, N, C = 1024, 431, 256
pts = torch.from_numpy(syn.make_sample_points(B, N, seed=2)).to(dev)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for hw in (14, 28):
    feat = torch.randn(B, C, hw, hw, device=dev)
    fl = feat.contiguous(memory_format=torch.channels_last)
    alg = B * (4 * C * (min(4 * N, hw * hw) + N) + 8 * N)
    t1 = timed(lambda: ops.sample_bilinear(feat, pts, ops.LAYOUT_NCHW))
    t2 = timed(lambda: ops.sample_bilinear(fl, pts, ops.LAYOUT_NCHW))
    print("env CG=%s DENSE=%s  %dx%d  nchw %.4f ms (%.3f)  channels_last %.4f ms (%.3f)" % (
        os.environ.get("WHMR_DENSE_CG"), os.environ.get("WHMR_SAMPLE_DENSE"), hw, hw, t1, alg / t1 / 1e6 / 6457.4, t2,
        alg / t2 / 1e6 / 6457.4), flush=True)
