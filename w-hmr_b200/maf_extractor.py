"""Drop-in for models/maf_extractor.py:MAF_Extractor (bound at models/whmr.py:333-335 as
`WHMR.maf_extractor[i]`; driven at :564,593,597,606).

Kept: constructor spelling, mutable `im_feat` / `cam` attributes, parameter names `conv0..2`, the
`Dmap` buffer, `sampling(points, im_feat=None, z_feat=None)`, `forward(p, center, scale, img_focal,
img_center, s_feat=None, cam=None)`, `project`, `get_trans`, `perspective_projection`, `reduce_dim`.

Inference (under `torch.no_grad()`, or when neither the parameters nor the inputs require a gradient): `sampling` / `forward` are ONE launch -- grid_sample, the weak
projection and the whole `reduce_dim` MLP run in `maf_fused_kernel` (3xTF32 on tcgen05; weights are
this module's `conv0..2` parameters, so checkpoints load unchanged), and with `return_point_feat =
False` the [B,256,N] point features the reference returns but never reads (models/whmr.py:597,606)
are not written at all.  Under autograd (training) the sampling custom op (which has a backward) and
the PyTorch `reduce_dim` below are used, as in the reference.  `fused = False` forces that path.
"""
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import constants, ops


class MAF_Extractor(nn.Module):
    def __init__(self, device=torch.device('cuda'), filter_channels=constants.MLP_DIM,
                 mesh_downsampling='data/mesh_downsampling.npz', Dmap=None, n_verts=constants.NUM_VERTS):
        super().__init__()
        self.device = device
        self.filters = []
        self.num_views = 1
        self.last_op = nn.ReLU(True)
        for l in range(0, len(filter_channels) - 1):
            cin = filter_channels[l] + (filter_channels[0] if l != 0 else 0)
            self.filters.append(nn.Conv1d(cin, filter_channels[l + 1], 1))
            self.add_module("conv%d" % l, self.filters[l])
        self.im_feat = None
        self.cam = None
        # models/maf_extractor.py:53-71: Dmap = D[1] @ D[0]  (6890 -> 431), kept as a buffer for
        # checkpoint compatibility; the hot path applies it as a sparse read-out (ops.Readout).
        if Dmap is None and mesh_downsampling and os.path.exists(mesh_downsampling):
            import scipy.sparse
            g = np.load(mesh_downsampling, allow_pickle=True, encoding='latin1')
            D = g['D']
            Dmap = (scipy.sparse.csr_matrix(D[1]) @ scipy.sparse.csr_matrix(D[0])).toarray()
        if Dmap is None:
            Dmap = np.zeros((431, n_verts), dtype=np.float32)
        self.register_buffer('Dmap', torch.as_tensor(np.asarray(Dmap), dtype=torch.float32))
        self.crop_size = constants.IMG_RES_WIDTH
        self.layout = ops.LAYOUT_NCHW
        self.filter_channels = tuple(int(c) for c in filter_channels)
        self.fused = True                # one-launch sampling + reduce_dim when no gradient is needed
        self.return_point_feat = True    # False: the fused path skips the [B,C_s,N] output and returns None for it
        self._mlp = None                 # ops.MafMlp, built lazily on the module's device

    def __getstate__(self):
        """copy.deepcopy / pickle carry the module without the native handle of the fused kernel (rebuilt on first use)."""
        d = self.__dict__.copy()
        d['_mlp'] = None
        return d

    def _fused_mlp(self, im_feat, *others):
        """The fused-kernel state, or None when this call must take the sampling op + PyTorch MLP path."""
        if not self.fused or self.num_views != 1 or not ops.MafMlp.supported(self.filter_channels):
            return None
        if not (torch.is_tensor(im_feat) and im_feat.is_cuda and im_feat.dtype == torch.float32):
            return None
        if torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad
                                           for t in tuple(self.parameters()) + (im_feat,) + others):
            return None      # the result must carry a grad_fn: sampling op with backward + PyTorch MLP
        if self._mlp is None or self._mlp.device != im_feat.device:
            self._mlp = ops.MafMlp(self.filter_channels, im_feat.device)
        self._mlp.set_weights(self.filters)
        return self._mlp

    def reduce_dim(self, feature):
        """models/maf_extractor.py:75-101 (the PyTorch MLP: training / autograd path and unsupported widths)."""
        y = feature
        tmpy = feature
        for i, f in enumerate(self.filters):
            y = self._modules['conv' + str(i)](y if i == 0 else torch.cat([y, tmpy], 1))
            if i != len(self.filters) - 1:
                y = F.leaky_relu(y)
            if self.num_views > 1 and i == len(self.filters) // 2:
                y = y.view(-1, self.num_views, y.shape[1], y.shape[2]).mean(dim=1)
                tmpy = feature.view(-1, self.num_views, feature.shape[1], feature.shape[2]).mean(dim=1)
        y = self.last_op(y)
        return y.view(y.shape[0], -1)

    def sampling(self, points, im_feat=None, z_feat=None):
        """models/maf_extractor.py:103-124.  points [B,N,2]; -> (mesh_align_feat [B,C_p*N], point_feat [B,C_s,N])"""
        if im_feat is None:
            im_feat = self.im_feat
        mlp = self._fused_mlp(im_feat, points)
        if mlp is not None:
            return mlp.sample(im_feat, points, self.layout, self.return_point_feat)
        if ops.is_host_map(im_feat):     # pinned host map: taps gathered in place over PCIe (no gradient to the map)
            point_feat = ops.sample_bilinear(im_feat, points.detach(), self.layout)
        else:
            point_feat = ops.sample_bilinear_op(im_feat, points, self.layout)
        return self.reduce_dim(point_feat), point_feat

    def forward(self, p, center, scale, img_focal, img_center, s_feat=None, cam=None, **kwargs):
        """models/maf_extractor.py:126-143: projection(p, cam) then sampling."""
        if cam is None:
            cam = self.cam
        im_feat = self.im_feat if s_feat is None else s_feat
        mlp = self._fused_mlp(im_feat, p, cam)
        if mlp is not None:
            return mlp.project_sample(im_feat, p, cam, constants.FOCAL_LENGTH, float(constants.IMG_RES_WIDTH),
                                      float(constants.IMG_RES_HEIGHT), self.layout, self.return_point_feat)[:2]
        if ops.is_host_map(im_feat):     # the reference detaches p and cam here (models/whmr.py:586-591)
            point_feat, _ = ops.project_sample(im_feat, p.detach(), cam.detach(), constants.FOCAL_LENGTH,
                                               float(constants.IMG_RES_WIDTH), float(constants.IMG_RES_HEIGHT), self.layout)
        else:
            point_feat, _ = ops.project_sample_op(im_feat, p, cam, constants.FOCAL_LENGTH,
                                                  float(constants.IMG_RES_WIDTH), float(constants.IMG_RES_HEIGHT),
                                                  self.layout)
        return self.reduce_dim(point_feat), point_feat

    def project(self, points, pred_cam, center, scale, img_focal, img_center, return_full=False):
        """models/maf_extractor.py:145-173 (not called on the live path; forward-only)."""
        if torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad for t in (points, pred_cam)):
            raise NotImplementedError("MAF_Extractor.project has no backward (the reference never differentiates it: "
                                      "models/whmr.py drives MAF_Extractor.forward); call it under torch.no_grad()")
        full, crop = ops.project_crop(points, pred_cam, center, scale, img_focal, img_center, self.crop_size,
                                      constants.IMG_RES_WIDTH, constants.IMG_RES_HEIGHT)
        return (full, crop) if return_full else crop

    def get_trans(self, pred_cam, center, scale, img_focal, img_center):
        """models/maf_extractor.py:175-190 (host-side glue on [B] vectors; not on the live path)."""
        b = scale * 200
        s, tx, ty = pred_cam.unbind(-1)
        bs = b * s
        return torch.stack([tx + 2 * (center[:, 0] - img_center[:, 0]) / bs,
                            ty + 2 * (center[:, 1] - img_center[:, 1]) / bs,
                            2 * img_focal / bs], dim=-1).unsqueeze(1)

    def perspective_projection(self, points, rotation, translation, focal_length, camera_center, distortion=None):
        """models/maf_extractor.py:192-235 (rotation / translation / distortion all optional)."""
        return ops.perspective_projection(points, rotation, translation, focal_length, camera_center,
                                          distortion=distortion)
