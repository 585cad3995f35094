"""Mesh-aligned feature sampling oracle -- TEST INFRASTRUCTURE.

`grid_sample_points` is the reference's call itself (models/maf_extractor.py:119:
F.grid_sample(im_feat, points.unsqueeze(2), align_corners=True)[..., 0], default mode
bilinear, default padding zeros).  `bilinear_points_np` is an independent float64 NumPy
derivation of the same semantics (pixel = (g+1)/2*(size-1); points[...,0] <-> W; taps outside
the map contribute zero) used to pin the torch call against a hand computation.
"""
import numpy as np
import torch


def grid_sample_points(im_feat, points):
    """im_feat [B,C,H,W], points [B,N,2] in [-1,1] (x<->W, y<->H) -> [B,C,N]."""
    return torch.nn.functional.grid_sample(im_feat, points.unsqueeze(2), align_corners=True)[..., 0]


def bilinear_points_np(im_feat, points):
    f = np.asarray(im_feat, dtype=np.float64)
    p = np.asarray(points, dtype=np.float64)
    B, C, H, W = f.shape
    N = p.shape[1]
    out = np.zeros((B, C, N), dtype=np.float64)
    x = (p[..., 0] + 1) * 0.5 * (W - 1)
    y = (p[..., 1] + 1) * 0.5 * (H - 1)
    x0 = np.floor(x).astype(np.int64)
    y0 = np.floor(y).astype(np.int64)
    for dy in (0, 1):
        for dx in (0, 1):
            xi = x0 + dx
            yi = y0 + dy
            w = (1 - np.abs(x - xi)) * (1 - np.abs(y - yi))
            ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
            xi_c = np.clip(xi, 0, W - 1)
            yi_c = np.clip(yi, 0, H - 1)
            for b in range(B):
                tap = f[b][:, yi_c[b], xi_c[b]]            # [C,N]
                out[b] += tap * (w[b] * ok[b])[None, :]
    return out


def reduce_dim(point_feat, convs):
    """models/maf_extractor.py:75-101 with num_views == 1: Conv1d(k=1) MLP with skip-concat of the
    input, leaky_relu between layers, ReLU at the end, flatten to [B, C_p*N].
    convs: list of (weight [Co,Ci,1], bias [Co])."""
    y = point_feat
    n = len(convs)
    for i, (w, b) in enumerate(convs):
        x = y if i == 0 else torch.cat([y, point_feat], 1)
        y = torch.nn.functional.conv1d(x, w, b)
        if i != n - 1:
            y = torch.nn.functional.leaky_relu(y)
    y = torch.relu(y)
    return y.reshape(y.shape[0], -1)
