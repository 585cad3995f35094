#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the REFERENCE's own code.

Run in the authoring container only (needs /root/reference, which never travels to the GPU
box):  python tests/golden/make_golden.py

What is imported from the reference (unmodified, read-only):
  * utils/geometry.py     -- projection, perspective_projection, convert_pare_to_full_img_cam,
                             batch_rodrigues (quaternion variant), rot6d_to_rotmat,
                             unbiased_gram_schmidt, rotation_matrix_to_angle_axis
  * models/maf_extractor.py (loaded by file path; the package import drags in `pare`)
                          -- MAF_Extractor.sampling / forward / project
  * utils/pose_utils.py   -- compute_similarity_transform_batch, reconstruction_error
`yacs` is not installed, so a ~25-line CfgNode stand-in is put on sys.modules and cfg is
loaded from the reference's configs/pymaf_config.yaml.  MAF_Extractor.__init__ reads
data/mesh_downsampling.npz from the cwd (absent from the reference tree): a synthetic one
with the same schema is written to a temp dir first.

Also extracts the real pose/shape rows of the reference's vendored test fixtures
(models/ViTPose/tests/data/{mosh/test_mosh.npz,h36m/test_h36m.npz,smpl/smpl_mean_params.npz})
into real_pose_shape.npz (inputs only).

SMPL itself cannot be run: smplx / pare / the SMPL weights are absent (parity unpinned).
"""
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np
import torch
import yaml

REF = os.environ.get('WHMR_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))


class CfgNode(dict):
    def __init__(self, init=None, new_allowed=False):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def merge_from_file(self, path):
        with open(path) as fh:
            self._merge(yaml.safe_load(fh))

    def _merge(self, d):
        for k, v in d.items():
            if isinstance(v, dict):
                if k not in self or not isinstance(self[k], CfgNode):
                    self[k] = CfgNode()
                self[k]._merge(v)
            else:
                self[k] = v

    def merge_from_list(self, lst):
        raise NotImplementedError

    def clone(self):
        return CfgNode(self)


def install_shims():
    yacs = types.ModuleType('yacs')
    yc = types.ModuleType('yacs.config')
    yc.CfgNode = CfgNode
    yacs.config = yc
    sys.modules['yacs'] = yacs
    sys.modules['yacs.config'] = yc
    sys.path.insert(0, REF)


def load_reference():
    install_shims()
    from core.cfgs import cfg
    cfg.merge_from_file(os.path.join(REF, 'configs', 'pymaf_config.yaml'))
    import utils.geometry as geo
    import utils.pose_utils as pu
    spec = importlib.util.spec_from_file_location('ref_maf_extractor',
                                                  os.path.join(REF, 'models', 'maf_extractor.py'))
    maf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(maf)
    return cfg, geo, pu, maf


def synthetic_mesh_downsampling(path, rng):
    import scipy.sparse as sp
    keep0 = np.sort(rng.choice(6890, 1723, replace=False))
    keep1 = np.sort(rng.choice(1723, 431, replace=False))
    D0 = sp.csr_matrix((np.ones(1723), (np.arange(1723), keep0)), shape=(1723, 6890))
    D1 = sp.csr_matrix((np.ones(431), (np.arange(431), keep1)), shape=(431, 1723))
    D = np.empty(2, dtype=object)
    D[0], D[1] = D0, D1
    A = np.empty(1, dtype=object); A[0] = sp.eye(2).tocsr()
    U = np.empty(1, dtype=object); U[0] = sp.eye(2).tocsr()
    np.savez(path, A=A, U=U, D=D)


def main():
    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    cfg, geo, pu, maf = load_reference()
    f32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))  # noqa: E731
    out = {}

    # ---- projection (utils/geometry.py:289-307) ----
    B, N = 6, 49
    pts = f32(rng.normal(0, 0.4, size=(B, N, 3)))
    cam = f32(np.stack([rng.uniform(0.6, 1.2, B), rng.uniform(-0.2, 0.2, B), rng.uniform(-0.2, 0.2, B)], 1))
    out['proj_points'] = pts.numpy(); out['proj_cam'] = cam.numpy()
    out['proj_out'] = geo.projection(pts, cam).numpy()

    # ---- full-image projection block (models/whmr.py:147-173) via the reference's functions ----
    bbox_h = f32(rng.uniform(100, 600, B)); Tz = f32(rng.uniform(1, 10, B))
    orig_shape = f32(np.array([[1080, 1920], [720, 1280]])[rng.integers(0, 2, B)])
    center = f32(rng.uniform(0.1, 0.9, size=(B, 2))) * orig_shape[:, [1, 0]]
    s = cam[:, 0]
    focal = s * bbox_h * Tz / 2.
    cc = orig_shape[:, [1, 0]] / 2.
    cam_t = geo.convert_pare_to_full_img_cam(cam, bbox_h, center, orig_shape[:, 1], orig_shape[:, 0], Tz=Tz)
    kp = geo.perspective_projection(pts, rotation=torch.eye(3).unsqueeze(0).expand(1, -1, -1),
                                    translation=cam_t, focal_length=focal, camera_center=cc)
    out.update(full_bbox_h=bbox_h.numpy(), full_Tz=Tz.numpy(), full_orig_shape=orig_shape.numpy(),
               full_center=center.numpy(), full_focal=focal.numpy(), full_cam_t=cam_t.numpy(),
               full_kp_px=kp.numpy(), full_kp_norm=(kp / cc.unsqueeze(1) - 1).numpy())
    # convert_pare_to_full_img_cam with an explicit focal length (:149-150 branch)
    out['full_cam_t_f5000'] = geo.convert_pare_to_full_img_cam(
        cam, bbox_h, center, orig_shape[:, 1], orig_shape[:, 0], focal_length=5000.).numpy()
    # perspective_projection with a general rotation and retain_z
    R = geo.batch_rodrigues(f32(rng.normal(0, 0.5, size=(B, 3))))
    out['pp_rot'] = R.numpy()
    out['pp_retain_z'] = geo.perspective_projection(pts, R, cam_t, focal, cc, retain_z=True).numpy()

    # ---- rotation helpers ----
    aa = f32(np.concatenate([np.zeros((1, 3)), rng.normal(0, 1.0, size=(15, 3))]))
    out['rodq_in'] = aa.numpy(); out['rodq_out'] = geo.batch_rodrigues(aa).numpy()
    x6 = f32(rng.normal(0, 1, size=(16, 6)))
    out['rot6d_in'] = x6.numpy(); out['rot6d_out'] = geo.rot6d_to_rotmat(x6).numpy()
    noisy = geo.batch_rodrigues(f32(rng.normal(0, 1.2, size=(24, 3)))) + f32(rng.normal(0, 0.05, size=(24, 3, 3)))
    noisy = noisy.reshape(2, 12, 3, 3)
    out['ugs_in'] = noisy.numpy(); out['ugs_out'] = geo.unbiased_gram_schmidt(noisy).numpy()
    Rm = geo.batch_rodrigues(f32(np.concatenate([np.zeros((1, 3)), rng.normal(0, 1.5, size=(63, 3)),
                                                 np.array([[np.pi, 0, 0], [0, 3.0, 0.2]])])))
    out['r2aa_in'] = Rm.numpy(); out['r2aa_out'] = geo.rotation_matrix_to_angle_axis(Rm.clone()).numpy()

    # ---- MAF_Extractor (models/maf_extractor.py) ----
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, 'data'))
        synthetic_mesh_downsampling(os.path.join(td, 'data', 'mesh_downsampling.npz'), rng)
        cwd = os.getcwd(); os.chdir(td)
        try:
            ext = maf.MAF_Extractor(device=torch.device('cpu'))
        finally:
            os.chdir(cwd)
    ext.eval()
    for k, v in ext.state_dict().items():
        if k.startswith('conv'):
            out['maf_' + k.replace('.', '_')] = v.numpy()
    Bm = 2
    for tag, (H, W), Np in (('a', (8, 6), 24), ('b', (5, 7), 17)):
        feat = f32(rng.normal(0, 1, size=(Bm, 256, H, W)))
        p = rng.uniform(-1.0, 1.0, size=(Bm, Np, 2))
        p[:, :3] *= 1.4                      # a few outside [-1,1] -> zero padding
        p[0, 3] = (-1.0, 1.0); p[1, 3] = (1.0, -1.0)   # exact corners
        p = f32(p)
        with torch.no_grad():
            maf_feat, pf = ext.sampling(p, im_feat=feat)
        out['samp_%s_feat' % tag] = feat.numpy(); out['samp_%s_points' % tag] = p.numpy()
        out['samp_%s_point_feat' % tag] = pf.numpy(); out['samp_%s_mesh_align' % tag] = maf_feat.numpy()
    # forward = projection + sampling (:126-143), through self.im_feat / self.cam state
    feat = f32(rng.normal(0, 1, size=(Bm, 256, 8, 6)))
    p3 = f32(rng.normal(0, 0.35, size=(Bm, 19, 3)))
    camm = cam[:Bm]
    ext.im_feat = feat; ext.cam = camm
    with torch.no_grad():
        maf_feat, pf = ext(p3, None, None, None, None)
    out.update(fwd_feat=feat.numpy(), fwd_p=p3.numpy(), fwd_cam=camm.numpy(),
               fwd_point_feat=pf.numpy(), fwd_mesh_align=maf_feat.numpy())
    # project (:145-173) incl. get_trans and the crop normalisation
    scale = bbox_h[:Bm] / 200.
    img_center = cc[:Bm]
    full2d, crop2d = ext.project(p3, camm, center[:Bm], scale, focal[:Bm], img_center, return_full=True)
    out.update(mproj_scale=scale.numpy(), mproj_center=center[:Bm].numpy(), mproj_focal=focal[:Bm].numpy(),
               mproj_img_center=img_center.numpy(), mproj_full=full2d.numpy(), mproj_crop=crop2d.numpy())
    kc = f32(rng.normal(0, 0.05, size=(Bm, 5)))
    tr = ext.get_trans(camm, center[:Bm], scale, focal[:Bm], img_center)
    out['mproj_kc'] = kc.numpy()
    out['mproj_distorted'] = ext.perspective_projection(p3 + tr, None, None, focal[:Bm], img_center,
                                                        distortion=kc).numpy()

    # ---- Procrustes (utils/pose_utils.py:10-75) ----
    S1 = rng.normal(0, 0.3, size=(5, 14, 3)); S2 = rng.normal(0, 0.3, size=(5, 14, 3))
    S2[0] = 0.5 * S1[0] + rng.uniform(0, 1, size=(1, 3))     # the vendored KAT's construction
    out['pa_S1'] = S1.astype(np.float32); out['pa_S2'] = S2.astype(np.float32)
    out['pa_S1_hat'] = pu.compute_similarity_transform_batch(S1.astype(np.float32), S2.astype(np.float32))
    re, _ = pu.reconstruction_error(S1.astype(np.float32), S2.astype(np.float32), reduction=None)
    out['pa_err'] = re

    # ---- estimate_translation (utils/geometry.py:344-408; host NumPy in the reference) ----
    Be = 9
    S3 = f32(rng.normal(0, 0.35, size=(Be, 49, 3)))
    true_t = f32(np.stack([rng.uniform(-0.4, 0.4, Be), rng.uniform(-0.4, 0.4, Be), rng.uniform(2.5, 12.0, Be)], axis=1))
    pe = S3 + true_t[:, None, :]
    kp = 5000. * pe[..., :2] / pe[..., 2:] + 112. + f32(rng.normal(0, 1.5, size=(Be, 49, 2)))   # noisy pixels
    conf = f32(rng.uniform(0, 1, size=(Be, 49, 1)))
    conf[1, 30:40] = 0.0                                      # undetected joints
    j2 = torch.cat([kp, conf], dim=-1)
    out['et_S'] = S3.numpy(); out['et_joints_2d'] = j2.numpy()
    out['et_out'] = geo.estimate_translation(S3, j2, focal_length=5000., img_size=[224., 224.]).numpy()
    out['et_out_f1000'] = geo.estimate_translation(S3, j2, focal_length=1000., img_size=[256., 192.]).numpy()

    np.savez_compressed(os.path.join(HERE, 'reference_outputs.npz'), **out)
    print('wrote reference_outputs.npz with', len(out), 'arrays,',
          os.path.getsize(os.path.join(HERE, 'reference_outputs.npz')) // 1024, 'KiB')

    # ---- real pose/shape rows (inputs only) ----
    td = os.path.join(REF, 'models', 'ViTPose', 'tests', 'data')
    mosh = np.load(os.path.join(td, 'mosh', 'test_mosh.npz'))
    h36m = np.load(os.path.join(td, 'h36m', 'test_h36m.npz'), allow_pickle=True)
    mean = np.load(os.path.join(td, 'smpl', 'smpl_mean_params.npz'))
    mean_R = geo.rot6d_to_rotmat(torch.from_numpy(mean['pose'][:]).reshape(1, 24, 6)).reshape(24, 3, 3)
    mean_aa = geo.rotation_matrix_to_angle_axis(mean_R.clone()).reshape(1, 72).numpy()
    pose = np.concatenate([mosh['pose'], h36m['pose'], mean_aa, np.zeros((1, 72))]).astype(np.float32)
    betas = np.concatenate([mosh['shape'], h36m['shape'], mean['shape'][None], np.zeros((1, 10))]).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, 'real_pose_shape.npz'), pose_aa=pose, betas=betas,
                        mean_pose_rot6d=mean['pose'].astype(np.float32), mean_cam=mean['cam'].astype(np.float32),
                        mean_rotmat=mean_R.numpy())
    print('wrote real_pose_shape.npz', pose.shape, betas.shape)


if __name__ == '__main__':
    main()
