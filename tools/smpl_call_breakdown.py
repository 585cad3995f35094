#!/usr/bin/env python
"""Per-call cost of the SMPL pieces at the loop's batch size, as graph replays of 5 back-to-back calls."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whmr_b200.synthetic as syn  # noqa: E402
from whmr_b200 import ops  # noqa: E402
from whmr_b200.loop import RegressorLoop, make_loop_inputs  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
model = syn.make_smpl_model(seed=0)
loop = RegressorLoop(model, dev)
feats, params, bbox = make_loop_inputs(B, dev)
h, _ = loop.smpl._state(dev)
ro = loop.head._readout(dev, True)


def graph_us(fn, n=300):
    s = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream(dev).wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = fn()
    for _ in range(10):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    del keep
    return a.elapsed_time(b) / n * 1000.0


def calls(fn):
    return lambda: [fn(p) for p in params]


ws, n = h.workspace(B)
t_chain = graph_us(calls(lambda p: h.stage_chain(p['betas'], p['rotmat'], True, ws, n)))
t_smpl = graph_us(calls(lambda p: ops.smpl_lbs(h.id, p['betas'], p['rotmat'], True)))
t_ro = graph_us(calls(lambda p: ops.smpl_lbs_readout(h.id, ro.id, p['betas'], p['rotmat'], True)))
print("B=%d per call: chain alone %.1f us | chain+fused (no read-outs) %.1f us | chain+fused(emits)+reduce %.1f us"
      % (B, t_chain / 5, t_smpl / 5, t_ro / 5))
