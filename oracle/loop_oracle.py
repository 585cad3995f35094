"""CPU restatement of the body-model part of WHMR.forward's regressor loop -- TEST INFRASTRUCTURE
(also the `cpu_baseline` / `--impl reference` arm of bench.py: it executes the path the way the
reference does on its CPU code path -- dense SMPL matmuls, dense [1723,6890] / [431,1723] Dmap
matmuls, expanded H36M regressor, F.grid_sample -- with torch CPU kernels on all host threads).

Follows models/whmr.py:550-651 with the MLPs factored out exactly as whmr_b200.loop.RegressorLoop
does: per iteration the regressor outputs (rotmat, betas, cam) are inputs."""
import numpy as np
import torch

from . import geometry_oracle as G
from .sampling_oracle import grid_sample_points
from .smpl_oracle import SMPLOracle, regressor_readouts


class LoopOracle:
    def __init__(self, model, backbone='vitpose', with_h36m=True, device='cpu', convs=None):
        """convs: optional 3 x [(weight, bias)] x 3 of the extractors' Conv1d MLPs; then `step` also returns
        'ref_features' = reduce_dim(point_feats[i]) (models/maf_extractor.py:75-101)."""
        self.convs = convs
        import importlib
        syn = importlib.import_module('whmr_b200.synthetic')
        self.model = model
        self.smpl = SMPLOracle(model, torch.float32, device=device)
        self.grid = torch.from_numpy(syn.grid_points(backbone)).to(device)
        self.with_h36m = with_h36m

    def regressor_outputs(self, p, bbox=None, is_train=False, train_stage=None):
        """Regressor.forward / forward_init after the MLP (models/whmr.py:128-209, 225-269).  Regressor.forward (bbox
        given) orthonormalises the predicted rotations in eval mode (:129-130); both compute `pose` / `theta` (:174,190)."""
        rotmat = p['rotmat']
        if bbox is not None and not is_train:
            rotmat = G.unbiased_gram_schmidt(rotmat)
        o = self.smpl(p['betas'], rotmat[:, 1:], rotmat[:, :1], pose2rot=False)
        verts, joints = o['vertices'], o['joints']
        r = regressor_readouts(self.model, verts)
        pose = G.rotation_matrix_to_angle_axis(rotmat.reshape(-1, 3, 3)).reshape(-1, 72)
        # models/whmr.py:142-165: with cfg.TRAIN.STAGE given, one of the two projections sees detached joints and the
        # predicted-focal block a detached camera (forward values are unchanged)
        jw = joints if train_stage in (None, 1) else joints.detach()
        out = {'verts': verts, 'joints49': joints, 'kp_2d': G.projection(jw, p['cam']), 'rotmat': rotmat, 'pose': pose,
               'theta': torch.cat([p['cam'], p['betas'], pose], dim=1),
               'sub_verts': r['sub_verts'], 'temp_verts': r['temp_verts'], 'markers': r['markers'],
               'smpl_kp_3d': r['smpl_kp_3d'], 'kp_3d': r['kp_3d_h36m'] if self.with_h36m else joints}
        if bbox is not None:
            jf = joints if train_stage in (None, 2) else joints.detach()
            camf = p['cam'] if train_stage is None else p['cam'].detach()
            kpn, focal, cam_t, _ = G.full_projection(jf, camf, bbox['bbox_height'], bbox['center'],
                                                     bbox['orig_shape'], bbox['Tz'])
            out.update(kp_2d_w=kpn, focal_length=focal, pred_cam_t=cam_t)
        return out

    def step(self, feats, params, bbox):
        B = feats[0].shape[0]
        out = self.regressor_outputs(params[0])
        point_feats = []
        for it in range(3):
            if it == 0:
                pts = self.grid.unsqueeze(0).expand(B, -1, -1)
            else:
                pts = G.projection(out['markers'], params[it]['cam'])
            point_feats.append(grid_sample_points(feats[it], pts))
            out = self.regressor_outputs(params[it + 1], bbox)
        g = self.smpl(params[4]['betas'], params[4]['rotmat'][:, 1:], params[4]['rotmat'][:, :1], pose2rot=False)
        res = dict(out)
        # models/whmr.py:632-633: global_pose = cat(axis-angle of the re-estimated global rotation, pose[:, 3:])
        res['global_pose'] = G.rotation_matrix_to_angle_axis(params[4]['rotmat'].reshape(-1, 3, 3)).reshape(-1, 72)
        res['point_feats'] = point_feats
        if self.convs is not None:
            from .sampling_oracle import reduce_dim
            res['ref_features'] = [reduce_dim(pf, cv) for pf, cv in zip(point_feats, self.convs)]
        res['global_verts'] = g['vertices']
        if self.with_h36m:
            res['global_kp_3d'] = regressor_readouts(self.model, g['vertices'])['kp_3d_h36m']
        else:
            res['global_kp_3d'] = g['joints']
        return res


def to_cpu_inputs(feats, params, bbox, n=None):
    """first n bodies of device inputs -> CPU tensors"""
    s = slice(0, n)
    f = [x[s].detach().cpu() for x in feats]
    p = [{k: v[s].detach().cpu() for k, v in q.items()} for q in params]
    b = {k: v[s].detach().cpu() for k, v in bbox.items()}
    return f, p, b


def make_cpu_inputs(B, backbone='vitpose', seed=1, channels=256):
    """CPU-only twin of whmr_b200.loop.make_loop_inputs (same parameter streams; feature maps from a
    CPU generator) for boxes without a GPU (`bench.py --impl reference`)."""
    import importlib
    syn = importlib.import_module('whmr_b200.synthetic')
    levels = ((32, 24), (64, 48), (128, 96)) if backbone == 'vitpose' else ((14, 14), (28, 28), (56, 56))
    g = torch.Generator().manual_seed(1000 * seed)
    feats = [torch.randn(B, channels, h, w, generator=g) for h, w in levels]
    base = syn.make_bodies(B, seed=seed)
    params = []
    for i in range(5):
        b = syn.make_bodies(B, seed=seed + 17 * (i + 1), with_real_rows=(i == 0))
        params.append({k: torch.from_numpy(b[k]) for k in ('rotmat', 'betas', 'cam')})
    bbox = {k: torch.from_numpy(np.ascontiguousarray(base[k])) for k in ('bbox_height', 'center', 'orig_shape', 'Tz')}
    return feats, params, bbox
