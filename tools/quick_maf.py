"""Quick GPU check + timing of the fused sampling + reduce_dim kernel (maf_fused_tc.cuh) against the two-step path
(sampling kernel + PyTorch MLP).  python tools/quick_maf.py [B N H W]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whmr_b200  # noqa: E402,F401
from oracle.sampling_oracle import grid_sample_points, reduce_dim  # noqa: E402
from whmr_b200.maf_extractor import MAF_Extractor  # noqa: E402
import whmr_b200.synthetic as syn  # noqa: E402


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run(B, N, H, W, check=True):
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(0)
    ext = MAF_Extractor(mesh_downsampling=None)
    with torch.no_grad():
        for p in ext.parameters():
            p.copy_(torch.randn(p.shape, generator=gen) * (0.5 if p.dim() == 1 else 2.0 / np.sqrt(p.shape[1])))
    convs = [(c.weight.detach().double(), c.bias.detach().double()) for c in ext.filters]
    ext = ext.to(dev).eval()
    feat = torch.randn(B, 256, H, W, device=dev)
    pts = torch.from_numpy(syn.make_sample_points(B, N, seed=1)).to(dev)
    res = {}
    for lay in ("nchw", "channels_last"):
        f_in = feat if lay == "nchw" else feat.contiguous(memory_format=torch.channels_last)
        with torch.no_grad():
            ext.fused = True
            maf, pf = ext.sampling(pts, im_feat=f_in)
            torch.cuda.synchronize()
            if check:
                nb = min(B, 8)
                ref_p = grid_sample_points(feat[:nb].double().cpu(), pts[:nb].double().cpu())
                ref_m = reduce_dim(ref_p, convs)
                ep = float((pf[:nb].double().cpu() - ref_p).abs().max() / ref_p.abs().max())
                em = float((maf[:nb].double().cpu() - ref_m).abs().max() / ref_m.abs().max())
            else:
                ep = em = float('nan')
            t_f = timed(lambda: ext.sampling(pts, im_feat=f_in))
            ext.return_point_feat = False
            t_f0 = timed(lambda: ext.sampling(pts, im_feat=f_in))
            ext.return_point_feat = True
            ext.fused = False
            maf_u, _ = ext.sampling(pts, im_feat=f_in)
            eu = float((maf_u - maf).abs().max() / maf_u.abs().max())
            t_u = timed(lambda: ext.sampling(pts, im_feat=f_in))
            from whmr_b200 import ops
            t_s = timed(lambda: ops.sample_bilinear(f_in, pts, ops.LAYOUT_NCHW))
        res[lay] = dict(err_point_feat=ep, err_mesh_align=em, fused_vs_torch_mlp=eu, ms_fused=t_f, ms_fused_no_pf=t_f0,
                        ms_sampling_plus_torch_mlp=t_u, ms_sampling_alone=t_s)
        print(B, N, H, W, lay, {k: (round(v, 7) if isinstance(v, float) else v) for k, v in res[lay].items()}, flush=True)
    return res


if __name__ == "__main__":
    if len(sys.argv) >= 5:
        run(*[int(a) for a in sys.argv[1:5]])
    else:
        run(2, 24, 8, 6)
        run(256, 63, 32, 24)
        run(256, 67, 64, 48)
        run(256, 67, 128, 96)
        for hw in (14, 28, 56):
            run(1024, 431, hw, hw)
