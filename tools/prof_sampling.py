"""One call of each sampler variant at BASELINE configs[3] sizes (for `ncu -k regex:sample_bilinear|maf_fused`).
python tools/prof_sampling.py [HW ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whmr_b200  # noqa: E402,F401
import whmr_b200.synthetic as syn  # noqa: E402
from whmr_b200 import ops  # noqa: E402
from whmr_b200.maf_extractor import MAF_Extractor  # noqa: E402

dev = torch.device("cuda:0")
B, N, C = 1024, 431, 256
sizes = [int(a) for a in sys.argv[1:]] or [14, 28]
pts = torch.from_numpy(syn.make_sample_points(B, N, seed=2)).to(dev)
ext = MAF_Extractor(mesh_downsampling=None).to(dev).eval()
ext.return_point_feat = False
for hw in sizes:
    feat = torch.randn(B, C, hw, hw, device=dev)
    fl = feat.contiguous(memory_format=torch.channels_last)
    for _ in range(2):
        ops.sample_bilinear(feat, pts, ops.LAYOUT_NCHW)
        ops.sample_bilinear(fl, pts, ops.LAYOUT_NCHW)
        with torch.no_grad():
            ext.sampling(pts, im_feat=feat)
            ext.sampling(pts, im_feat=fl)
    torch.cuda.synchronize()
