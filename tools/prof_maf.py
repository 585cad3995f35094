"""One call of the fused sampling + MLP kernel at BASELINE configs[3] sizes and at a live-path level (for ncu -k regex:maf_fused)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whmr_b200  # noqa: E402,F401
import whmr_b200.synthetic as syn  # noqa: E402
from whmr_b200.maf_extractor import MAF_Extractor  # noqa: E402

dev = torch.device("cuda:0")
ext = MAF_Extractor(mesh_downsampling=None).to(dev).eval()
ext.return_point_feat = False
for (B, N, hw) in ((1024, 431, (14, 14)), (1024, 431, (56, 56)), (256, 67, (128, 96))):
    pts = torch.from_numpy(syn.make_sample_points(B, N, seed=2)).to(dev)
    feat = torch.randn(B, 256, *hw, device=dev)
    fl = feat.contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        for _ in range(2):
            ext.sampling(pts, im_feat=fl)
            ext.sampling(pts, im_feat=feat)
    torch.cuda.synchronize()
    del feat, fl
