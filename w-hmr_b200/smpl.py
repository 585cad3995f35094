"""Drop-in for `pare.models.SMPL` (= `smplx.SMPL` + the extra-joint wrapper whose twin is
commented at models/smpl.py:61-83 of the reference), bound at models/whmr.py:59 as
`Regressor.smpl` and at core/trainer.py:54-66 as `Trainer.smpl`.

Same constructor spelling (`SMPL(model_dir, batch_size=..., create_transl=False, gender=...)`),
same forward signature and output attributes (`vertices, joints, global_orient, body_pose, betas,
full_pose`), same buffer / parameter names as smplx (so reference checkpoints load with
strict=True), `.faces`, `.J_regressor`, `.to(device)`.  The arithmetic runs in the sm_100a kernels
behind `whmr_smpl_forward`; calling it on CPU tensors raises (no fallback).
"""
import os
from collections import namedtuple

import numpy as np
import torch
import torch.nn as nn

from . import constants, ops
from .synthetic import load_smpl_pkl

ModelOutput = namedtuple('ModelOutput', ['vertices', 'joints', 'full_pose', 'betas', 'global_orient',
                                         'body_pose', 'smpl_joints', 'rel_transforms'])
ModelOutput.__new__.__defaults__ = (None,) * len(ModelOutput._fields)


class SMPL(nn.Module):
    NUM_JOINTS = 23
    NUM_BODY_JOINTS = 23
    NUM_BETAS = 10

    def __init__(self, model_path=None, batch_size=1, create_transl=False, gender='neutral',
                 create_betas=True, create_global_orient=True, create_body_pose=True,
                 J_regressor_extra=None, vertex_ids=None, joint_map=None, model=None,
                 gemm_mode=None, dtype=torch.float32, **kwargs):
        """model_path: directory holding SMPL_{GENDER}.pkl or a .pkl path (as smplx); or pass the
        arrays directly as `model` (dict in smplx's in-memory layout, see synthetic.make_smpl_model).
        J_regressor_extra: [9,V] array or .npy path (default: model['J_regressor_extra'] or
        data/J_regressor_extra.npy, core/path_config.py:11)."""
        super().__init__()
        if model is None:
            if model_path is None:
                raise ValueError("SMPL: give model_path (dir or .pkl) or model=dict")
            path = model_path
            if os.path.isdir(path):
                path = os.path.join(path, 'SMPL_%s.pkl' % gender.upper())
            model = load_smpl_pkl(path)
        self.gender = gender
        self.batch_size = batch_size
        self.dtype = dtype
        if gemm_mode is None:   # production default: tcgen05 bf16x3; WHMR_GEMM_MODE overrides (fp32_simt | bf16x3 | 3xtf32)
            gemm_mode = os.environ.get('WHMR_GEMM_MODE', 'bf16x3')
        self.gemm_mode = ops.GEMM_MODES[gemm_mode] if isinstance(gemm_mode, str) else int(gemm_mode)
        V = model['v_template'].shape[0]
        t = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32)))  # noqa: E731
        self.faces = np.asarray(model['f']) if 'f' in model else None
        if self.faces is not None:
            self.register_buffer('faces_tensor', torch.from_numpy(self.faces.astype(np.int64)))
        self.register_buffer('v_template', t(model['v_template']))
        self.register_buffer('shapedirs', t(model['shapedirs'])[:, :, :self.NUM_BETAS].contiguous())
        self.register_buffer('J_regressor', t(model['J_regressor']))
        self.register_buffer('posedirs', t(model['posedirs']))
        parents = np.asarray(model['parents']).astype(np.int64).copy()
        parents[0] = -1
        self.register_buffer('parents', torch.from_numpy(parents))
        self.register_buffer('lbs_weights', t(model['weights']))

        if J_regressor_extra is None:
            J_regressor_extra = model.get('J_regressor_extra')
        if J_regressor_extra is None and os.path.exists('data/J_regressor_extra.npy'):
            J_regressor_extra = 'data/J_regressor_extra.npy'
        if isinstance(J_regressor_extra, str):
            J_regressor_extra = np.load(J_regressor_extra)
        if J_regressor_extra is None:
            raise ValueError("SMPL: J_regressor_extra ([9,V]) is required (models/smpl.py:66-69)")
        self.register_buffer('J_regressor_extra', t(J_regressor_extra))
        if vertex_ids is None:
            vertex_ids = model.get('vertex_ids')
        if vertex_ids is None:
            vertex_ids = constants.vertex_joint_selector_ids()
        elif isinstance(vertex_ids, dict):
            vertex_ids = constants.vertex_joint_selector_ids(vertex_ids)
        # smplx keeps these under vertex_joint_selector.extra_joints_idxs
        self.vertex_joint_selector = nn.Module()
        self.vertex_joint_selector.register_buffer(
            'extra_joints_idxs', torch.as_tensor(np.asarray(vertex_ids), dtype=torch.long))
        jm = constants.JOINT_MAP_49 if joint_map is None else joint_map
        self.joint_map = torch.tensor(list(jm), dtype=torch.long)   # plain attribute, as in the reference

        # smplx default parameters (state_dict compatibility); used when an argument is None
        if create_betas:
            self.betas = nn.Parameter(torch.zeros(batch_size, self.NUM_BETAS))
        if create_global_orient:
            self.global_orient = nn.Parameter(torch.zeros(batch_size, 3))
        if create_body_pose:
            self.body_pose = nn.Parameter(torch.zeros(batch_size, self.NUM_BODY_JOINTS * 3))
        if create_transl:
            self.transl = nn.Parameter(torch.zeros(batch_size, 3))
        self._dev = {}   # device -> (SmplHandle, Readout)
        assert self.v_template.shape == (V, 3)

    # -- device-side state ---------------------------------------------------------------------
    def __getstate__(self):
        """copy.deepcopy / pickle (EMA copies, spawn-start DataLoader workers: datasets/base_dataset.py:145 builds an SMPL
        per dataset) carry the module, not its native handles: those are rebuilt lazily on first use."""
        d = self.__dict__.copy()
        d['_dev'] = {}
        return d

    def _load_from_state_dict(self, *a, **k):
        super()._load_from_state_dict(*a, **k)
        self._dev = {}   # model buffers may have changed: re-arrange lazily

    def _state(self, device):
        key = str(device)
        st = self._dev.get(key)
        if st is None:
            c = lambda b: b.detach().cpu().numpy()  # noqa: E731
            V, J = self.v_template.shape[0], self.J_regressor.shape[0]
            h = ops.SmplHandle(c(self.v_template), c(self.shapedirs), c(self.posedirs), c(self.J_regressor),
                               c(self.lbs_weights), c(self.parents), device, self.gemm_mode)
            # joints = cat(24 chain joints, verts[:, extra_joints_idxs], J_regressor_extra @ verts)[:, joint_map]
            import scipy.sparse as sp
            vid = c(self.vertex_joint_selector.extra_joints_idxs)
            n45 = J + len(vid)
            src54 = sp.vstack([
                sp.csr_matrix((np.ones(J), (np.arange(J), V + np.arange(J))), shape=(J, V + J)),
                sp.csr_matrix((np.ones(len(vid)), (np.arange(len(vid)), vid)), shape=(len(vid), V + J)),
                sp.hstack([sp.csr_matrix(c(self.J_regressor_extra).astype(np.float64)),
                           sp.csr_matrix((self.J_regressor_extra.shape[0], J))]),
            ]).tocsr()
            jm = self.joint_map.numpy()
            ro = ops.Readout([('joints', src54[jm]), ('smpl_joints', src54[:n45])], V, J, device)
            st = (h, ro)
            self._dev[key] = st
        return st

    def set_gemm_mode(self, mode):
        self.gemm_mode = ops.GEMM_MODES[mode] if isinstance(mode, str) else int(mode)
        for h, _ in self._dev.values():
            h.set_gemm_mode(self.gemm_mode)

    # -- forward ---------------------------------------------------------------------------------
    def forward(self, betas=None, body_pose=None, global_orient=None, transl=None, return_verts=True,
                return_full_pose=False, pose2rot=True, return_transforms=False, **kwargs):
        """smplx.SMPL.forward + the PARE wrapper.  Rotation-matrix mode (models/whmr.py:132-137):
        body_pose [B,23,3,3], global_orient [B,1,3,3], pose2rot=False.  Axis-angle mode
        (core/trainer.py:415): body_pose [B,69], global_orient [B,3]."""
        global_orient = global_orient if global_orient is not None else self.global_orient
        body_pose = body_pose if body_pose is not None else self.body_pose
        betas = betas if betas is not None else self.betas
        if transl is None and hasattr(self, 'transl'):
            transl = self.transl
        B = max(betas.shape[0], global_orient.shape[0], body_pose.shape[0])
        if betas.shape[0] != B:
            betas = betas.expand(int(B / betas.shape[0]) * betas.shape[0], -1) if betas.shape[0] == 1 \
                else betas.repeat(int(B / betas.shape[0]), 1)
        J = self.J_regressor.shape[0]
        h, ro = self._state(betas.device)
        needs_grad = torch.is_grad_enabled() and any(t.requires_grad for t in (betas, body_pose, global_orient))
        if (not needs_grad and not return_full_pose and global_orient.shape[0] == B and body_pose.shape[0] == B
                and betas.is_cuda):
            # inference: the two pose arguments go to the kernel as they are (whmr_smpl_glue.root_pose) -- no concatenation
            verts, joints24, A, flat, _ = h.forward(betas, body_pose, not pose2rot, transl=transl,
                                                    want_transforms=return_transforms, readout=ro, root_pose=global_orient)
            r = ro.split(flat, B)
            return ModelOutput(vertices=verts if return_verts else None, joints=r['joints'], full_pose=None, betas=betas,
                               global_orient=global_orient, body_pose=body_pose, smpl_joints=r['smpl_joints'],
                               rel_transforms=A)
        if pose2rot:
            full_pose = torch.cat([global_orient.reshape(-1, 3).expand(B, -1) if global_orient.shape[0] != B
                                   else global_orient.reshape(B, 3),
                                   body_pose.reshape(body_pose.shape[0], -1).expand(B, -1)], dim=1)
        else:
            full_pose = torch.cat([global_orient.reshape(-1, 1, 3, 3).expand(B, -1, -1, -1),
                                   body_pose.reshape(body_pose.shape[0], J - 1, 3, 3).expand(B, -1, -1, -1)], dim=1)
        A = None
        if transl is None and not return_transforms:
            # torch.library op: differentiable w.r.t. betas / rotation matrices (core/trainer.py:380-636 back-propagates
            # through pred_vertices and pred_keypoints_3d)
            verts, joints24, flat = ops.smpl_lbs_readout(h.id, ro.id, betas, full_pose.contiguous(), not pose2rot)
        else:
            verts, joints24, A, flat, _ = h.forward(betas, full_pose, not pose2rot, transl=transl,
                                                    want_transforms=return_transforms, readout=ro)
        r = ro.split(flat, B)
        return ModelOutput(vertices=verts if return_verts else None, joints=r['joints'],
                           full_pose=full_pose if return_full_pose else None, betas=betas,
                           global_orient=global_orient, body_pose=body_pose, smpl_joints=r['smpl_joints'],
                           rel_transforms=A)


def get_smpl_faces(model_dir='data/smpl'):
    """models/smpl.py:86-89"""
    return SMPL(model_dir, batch_size=1, create_transl=False).faces
