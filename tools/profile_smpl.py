#!/usr/bin/env python
"""Small driver for ncu / compute-sanitizer runs: SMPL forward + read-outs + one sampling pass at a
chosen batch, outside CUDA graphs so every kernel is a plain launch."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whmr_b200.synthetic as syn  # noqa: E402
from whmr_b200.loop import RegressorLoop, make_loop_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16384)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--gemm-mode", default=None)
ap.add_argument("--loop-batch", type=int, default=0, help="also run the full regressor loop at this batch")
ap.add_argument("--weights", default="random")
a = ap.parse_args()
dev = torch.device("cuda:0")
model = syn.make_smpl_model(seed=0, weights=a.weights)
loop = RegressorLoop(model, dev, gemm_mode=a.gemm_mode)
b = syn.make_bodies(a.batch, seed=5)
betas = torch.from_numpy(b["betas"]).to(dev)
rot = torch.from_numpy(b["rotmat"]).to(dev)
cam = torch.from_numpy(b["cam"]).to(dev)
if a.loop_batch:
    feats, params, bbox = make_loop_inputs(a.loop_batch, dev)
for rep in range(a.reps):
    if rep == a.reps - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()     # ncu --profile-from-start off: only the last pass is captured
    out = loop.head(rot, betas, cam, J_regressor=True)
    if a.loop_batch:
        loop.step(feats, params, bbox)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", out["verts"].shape)
