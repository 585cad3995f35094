#!/usr/bin/env python
"""Quick timing of the SMPL forward (chain + fused kernel) alone and with the BodyModelHead read-out table, plus a
parity spot check against the fp32 oracle.  usage: quick_smpl.py [B ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whmr_b200.synthetic as syn  # noqa: E402
from whmr_b200 import ops  # noqa: E402
from whmr_b200.loop import RegressorLoop  # noqa: E402

dev = torch.device("cuda:0")
model = syn.make_smpl_model(seed=0)
loop = RegressorLoop(model, dev)
h, _ = loop.smpl._state(dev)
ro = loop.head._readout(dev, True)
from oracle.smpl_oracle import SMPLOracle  # noqa: E402
orc = SMPLOracle(model)
for B in [int(x) for x in (sys.argv[1:] or ["256", "4096", "16384"])]:
    b = syn.make_bodies(B, seed=5)
    betas = torch.from_numpy(b["betas"]).to(dev)
    rot = torch.from_numpy(b["rotmat"]).to(dev)
    v, j, _ = h.forward(betas, rot, True)
    n = min(B, 24)
    sel = list(range(n // 2)) + list(range(B - n // 2, B))
    ref = orc(b["betas"][sel], b["rotmat"][sel][:, 1:], b["rotmat"][sel][:, :1], pose2rot=False)
    err = float((v[sel].cpu() - ref["vertices"]).abs().max())
    res = []
    for name, fn in (("smpl", lambda: h.forward(betas, rot, True)),
                     ("smpl+readouts", lambda: ops.smpl_lbs_readout(h.id, ro.id, betas, rot, True))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        reps = 20 if B <= 4096 else 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res.append("%s %.1f us = %.2f M bodies/s" % (name, ms * 1e3, B / ms / 1e3))
    print("B=%d: %s | max vertex error %.2e m" % (B, " | ".join(res), err), flush=True)
