#!/usr/bin/env python
"""Small driver for compute-sanitizer (memcheck / racecheck / synccheck): SMPL forward with read-outs at ragged batch
sizes, sampling in both regimes, projection."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whmr_b200.synthetic as syn  # noqa: E402
from whmr_b200 import ops  # noqa: E402
from whmr_b200.loop import RegressorLoop, make_loop_inputs  # noqa: E402

dev = torch.device("cuda:0")
model = syn.make_smpl_model(seed=0)
loop = RegressorLoop(model, dev)
for B in (1, 17, 80):
    feats, params, bbox = make_loop_inputs(B, dev, seed=3)
    out = loop.step(feats, params, bbox)
    torch.cuda.synchronize()
    print("B=%d ok" % B, float(out["verts"].abs().max()))
pts = torch.from_numpy(syn.make_sample_points(4, 431, seed=2)).to(dev)
for hw in ((14, 14), (56, 56)):
    f = torch.randn(4, 256, *hw, device=dev)
    o = ops.sample_bilinear(f, pts, ops.LAYOUT_NCHW)
    torch.cuda.synchronize()
    print("sample", hw, float(o.abs().max()))
# backward: SMPL (skin, pose-blend transpose, chain), transposed read-out
from whmr_b200.smpl import SMPL  # noqa: E402
smpl = SMPL(model=model).to(dev)
for B in (1, 13):
    b = syn.make_bodies(B, seed=5)
    rm = torch.from_numpy(b['rotmat']).to(dev).requires_grad_(True)
    be = torch.from_numpy(b['betas']).to(dev).requires_grad_(True)
    o = smpl(betas=be, body_pose=rm[:, 1:], global_orient=rm[:, :1], pose2rot=False)
    (o.vertices.sum() + o.joints.pow(2).sum()).backward()
    torch.cuda.synchronize()
    print("smpl backward B=%d" % B, float(rm.grad.abs().max()), float(be.grad.abs().max()))
