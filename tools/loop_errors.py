#!/usr/bin/env python
"""Per-tensor max error of RegressorLoop.step (deferred and immediate schedules) against the CPU oracle loop."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whmr_b200.synthetic as syn  # noqa: E402
from oracle.loop_oracle import LoopOracle, to_cpu_inputs  # noqa: E402
from whmr_b200.loop import RegressorLoop, make_loop_inputs  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 12
model = syn.make_smpl_model(seed=0)
loop = RegressorLoop(model, dev)
feats, params, bbox = make_loop_inputs(B, dev, seed=4)
ref = LoopOracle(model).step(*to_cpu_inputs(feats, params, bbox))
for name, defer in (("deferred", True), ("immediate", False)):
    loop.defer = defer
    got = loop.step(feats, params, bbox)
    torch.cuda.synchronize()
    print("==", name)
    for k in sorted(ref):
        if k == "point_feats":
            e = max(float((a.cpu() - b).abs().max()) for a, b in zip(got[k], ref[k]))
        elif k in got and torch.is_tensor(ref[k]):
            e = float((got[k].cpu().double() - ref[k].double()).abs().max())
        else:
            continue
        print("  %-14s max abs err %.3e" % (k, e))


def report(tag, got):
    print("==", tag)
    for k in sorted(ref):
        if k == "point_feats":
            e = max(float((a.cpu() - b).abs().max()) for a, b in zip(got[k], ref[k]))
        elif k in got and torch.is_tensor(ref[k]):
            e = float((got[k].cpu().double() - ref[k].double()).abs().max())
        else:
            continue
        if e > 1e-4:
            print("  %-14s max abs err %.3e  <-- " % (k, e))


loop.defer = True
for overlap in (True, False):
    loop.overlap = overlap
    report("eager overlap=%s" % overlap, loop.step(feats, params, bbox))
    g, outs = loop.capture(feats, params, bbox)
    g.replay()
    torch.cuda.synchronize()
    report("graph overlap=%s (no zeroing)" % overlap, outs)
    for k in ('verts', 'global_verts', 'kp_3d', 'markers', 'kp_2d', 'kp_2d_w', 'pose', 'theta', 'rotmat', 'global_pose'):
        outs[k].zero_()
        g.replay()
        torch.cuda.synchronize()
        bad = [kk for kk in sorted(ref) if kk != 'point_feats' and kk in outs and torch.is_tensor(ref[kk])
               and float((outs[kk].cpu().double() - ref[kk].double()).abs().max()) > 1e-4]
        if bad:
            print("   after zeroing %-12s: wrong %s" % (k, bad))
