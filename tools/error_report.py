#!/usr/bin/env python
"""Max-abs error of the SMPL forward vs the fp64 oracle, per pose-blend arithmetic mode and skinning kernel."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whmr_b200.synthetic as syn  # noqa: E402
from oracle.smpl_oracle import SMPLOracle  # noqa: E402
from whmr_b200.smpl import SMPL  # noqa: E402

dev = torch.device("cuda:0")
B = 128
for weights in ("random", "skeleton"):
    model = syn.make_smpl_model(seed=0, weights=weights)
    b = syn.make_bodies(B, seed=1)
    ref = SMPLOracle(model, torch.float64)(b["betas"].astype(np.float64), b["rotmat"][:, 1:].astype(np.float64),
                                           b["rotmat"][:, :1].astype(np.float64), pose2rot=False)
    ref32 = SMPLOracle(model, torch.float32)(b["betas"], b["rotmat"][:, 1:], b["rotmat"][:, :1], pose2rot=False)
    print("%s: fp32 oracle vs fp64 oracle: verts %.2e" % (weights, float((ref32["vertices"].double() - ref["vertices"]).abs().max())))
    for skin in ("tc", "simt"):
        os.environ["WHMR_SKIN"] = skin
        for mode in ("fp32_simt", "bf16x3", "3xtf32"):
            smpl = SMPL(model=model, gemm_mode=mode).to(dev)
            T = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
            out = smpl(betas=T(b["betas"]), body_pose=T(b["rotmat"][:, 1:]), global_orient=T(b["rotmat"][:, :1]), pose2rot=False)
            ev = float((out.vertices.double().cpu() - ref["vertices"]).abs().max())
            ej = float((out.joints.double().cpu() - ref["joints"]).abs().max())
            print("  skin=%-4s gemm=%-9s verts %.2e m   joints %.2e m" % (skin, mode, ev, ej))
