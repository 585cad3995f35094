#!/usr/bin/env python
"""WHMR_FUSED_DEBUG driver: SMPL forward at a few batch sizes, per-role wait cycles printed by the library."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whmr_b200.synthetic as syn  # noqa: E402
from whmr_b200.smpl import SMPL  # noqa: E402

dev = torch.device("cuda:0")
model = syn.make_smpl_model(seed=0)
smpl = SMPL(model=model).to(dev)
for B in [int(x) for x in (sys.argv[1:] or ["256", "768"])]:
    b = syn.make_bodies(B, seed=5)
    betas = torch.from_numpy(b["betas"]).to(dev)
    rot = torch.from_numpy(b["rotmat"]).to(dev)
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = smpl(betas=betas, body_pose=rot[:, 1:], global_orient=rot[:, :1], pose2rot=False)
        torch.cuda.synchronize()
        print("B=%d rep %d wall %.1f us" % (B, rep, (time.perf_counter() - t0) * 1e6), flush=True)
