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
# channels_last maps: NHWC gather + the dense-regime kernel; fused sampling + reduce_dim MLP (tcgen05, both layouts, with
# and without the projection, tile tail + more than one tile per CTA is covered by the GPU tests)
from whmr_b200.maf_extractor import MAF_Extractor  # noqa: E402
ext = MAF_Extractor(mesh_downsampling=None).to(dev).eval()
for hw in ((14, 14), (64, 48)):
    f = torch.randn(4, 256, *hw, device=dev)
    fl = f.contiguous(memory_format=torch.channels_last)
    o = ops.sample_bilinear(fl, pts, ops.LAYOUT_NCHW)
    with torch.no_grad():
        m1, p1 = ext.sampling(pts, im_feat=f)
        m2, p2 = ext.sampling(pts, im_feat=fl)
        ext.im_feat, ext.cam = f, torch.tensor([[0.9, 0.0, 0.1]], device=dev).repeat(4, 1)
        m3, _ = ext(torch.randn(4, 67, 3, device=dev) * 0.3, None, None, None, None)
    torch.cuda.synchronize()
    print("maf fused", hw, float(o.abs().max()), float((m1 - m2).abs().max()), float(m3.abs().max()))
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
# last session: the kind::tf32 instantiation of the fused SMPL kernel (3xTF32) and sampling from pinned host maps
smpl3 = SMPL(model=model, gemm_mode="3xtf32").to(dev)
assert smpl3._state(dev)[0].is_fused()
for B in (1, 37, 130):
    b = syn.make_bodies(B, seed=7)
    o = smpl3(betas=torch.from_numpy(b['betas']).to(dev), body_pose=torch.from_numpy(b['rotmat'][:, 1:]).to(dev),
              global_orient=torch.from_numpy(b['rotmat'][:, :1]).to(dev), pose2rot=False)
    torch.cuda.synchronize()
    print("smpl 3xtf32 fused B=%d" % B, float(o.vertices.abs().max()))
hf = torch.randn(3, 64, 128, 96).pin_memory()
pts3 = torch.from_numpy(syn.make_sample_points(3, 67, seed=4)).to(dev)
oh = ops.sample_bilinear(hf, pts3)
od = ops.sample_bilinear(hf.to(dev), pts3)
torch.cuda.synchronize()
print("host-map sampling identical", bool(torch.equal(oh, od)))
