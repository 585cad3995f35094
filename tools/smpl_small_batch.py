"""SMPL forward (chain + fused kernel, no read-outs) as a CUDA-graph replay at small batch sizes: how the launch time
scales below the plateau.  usage: smpl_small_batch.py [B ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whmr_b200.synthetic as syn  # noqa: E402
from whmr_b200.loop import RegressorLoop  # noqa: E402

dev = torch.device("cuda:0")
model = syn.make_smpl_model(seed=0)
loop = RegressorLoop(model, dev)
h, _ = loop.smpl._state(dev)
for B in [int(x) for x in (sys.argv[1:] or ["16", "64", "128", "256", "512", "1024", "2048"])]:
    b = syn.make_bodies(B, seed=5)
    betas = torch.from_numpy(b["betas"]).to(dev)
    rot = torch.from_numpy(b["rotmat"]).to(dev)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            h.forward(betas, rot, True)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            out = h.forward(betas, rot, True)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 200 * 1e3
    print("B=%5d: %.1f us per SMPL forward (chain + fused, graph of 10) = %.2f M bodies/s" % (B, us, B / us), flush=True)
