"""Batched torch-CPU restatement of the reference's SMPL forward -- TEST INFRASTRUCTURE.

parity unpinned: the reference calls `pare.models.SMPL` (models/whmr.py:32,59) which
subclasses `smplx.SMPL` (smplx==0.1.28, environment.yml:167); neither is in the tree.
This file restates the *published* smplx 0.1.28 algorithm (lbs.py: blend_shapes,
vertices2joints, batch_rodrigues, batch_rigid_transform, lbs; body_models.py: SMPL.forward;
vertex_joint_selector.py) in the same operation order, plus the PARE/SPIN wrapper whose
line-for-line twin is commented in the reference at models/smpl.py:61-83.  The in-tree
statement of the same math is models/smpl_webuser/ (lbs.py:27-79, verts.py:42-50,
posemapper.py:36-43, serialization.py:102-108); `smpl_webuser_oracle.py` follows that one
independently and tests/test_oracle_cpu.py makes the two agree.

Works in float32 (the reference's dtype) or float64 (to apportion error).
"""
import numpy as np
import torch


def batch_rodrigues(rot_vecs):
    """smplx.lbs.batch_rodrigues (used by SMPL.forward when pose2rot=True; call sites
    core/trainer.py:415,420,781, evaluate/eval.py:159,200).  [N,3] -> [N,3,3].
    angle = ||v + 1e-8||;  R = I + sin*K + (1 - cos)*K.K"""
    n = rot_vecs.shape[0]
    dtype = rot_vecs.dtype
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    rot_dir = rot_vecs / angle
    cos = torch.unsqueeze(torch.cos(angle), dim=1)
    sin = torch.unsqueeze(torch.sin(angle), dim=1)
    rx, ry, rz = torch.split(rot_dir, 1, dim=1)
    zeros = torch.zeros((n, 1), dtype=dtype, device=rot_vecs.device)
    K = torch.cat([zeros, -rz, ry, rz, zeros, -rx, -ry, rx, zeros], dim=1).view(n, 3, 3)
    ident = torch.eye(3, dtype=dtype, device=rot_vecs.device).unsqueeze(0)
    return ident + sin * K + (1 - cos) * torch.bmm(K, K)


def blend_shapes(betas, shape_disps):
    """smplx.lbs.blend_shapes: einsum('bl,mkl->bmk').  Twin: models/smpl_webuser/verts.py:42-45."""
    return torch.einsum('bl,mkl->bmk', betas, shape_disps)


def vertices2joints(J_regressor, vertices):
    """smplx.lbs.vertices2joints: einsum('bik,ji->bjk').  Twin: serialization.py:104-107."""
    return torch.einsum('bik,ji->bjk', vertices, J_regressor)


def batch_rigid_transform(rot_mats, joints, parents):
    """smplx.lbs.batch_rigid_transform.  Twin: models/smpl_webuser/lbs.py:27-60
    (global_rigid_transformation): G_0 = [R_0|J_0], G_i = G_parent . [R_i | J_i - J_parent],
    A_i = G_i - [0 | G_i.(J_i,0)]."""
    B, nj = joints.shape[:2]
    joints = joints.unsqueeze(-1)
    rel = joints.clone()
    rel[:, 1:] = rel[:, 1:] - joints[:, parents[1:]]
    tm = torch.zeros(B, nj, 4, 4, dtype=joints.dtype, device=joints.device)
    tm[:, :, :3, :3] = rot_mats
    tm[:, :, :3, 3:] = rel
    tm[:, :, 3, 3] = 1.0
    chain = [tm[:, 0]]
    for i in range(1, nj):
        chain.append(torch.matmul(chain[int(parents[i])], tm[:, i]))
    transforms = torch.stack(chain, dim=1)
    posed_joints = transforms[:, :, :3, 3]
    joints_h = torch.nn.functional.pad(joints, [0, 0, 0, 1])
    rel_transforms = transforms - torch.nn.functional.pad(
        torch.matmul(transforms, joints_h), [3, 0, 0, 0, 0, 0, 0, 0])
    return posed_joints, rel_transforms


def lbs(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights,
        pose2rot=True):
    """smplx.lbs.lbs, operation order preserved: shape blend -> rest joints from
    v_shaped -> (rodrigues) -> pose feature (R[1:]-I row-major) -> pose offsets -> chain ->
    skinning.  Returns (verts [B,V,3], posed chain joints [B,24,3], and intermediates)."""
    B = max(betas.shape[0], pose.shape[0])
    dtype = betas.dtype
    v_shaped = v_template + blend_shapes(betas, shapedirs)
    J = vertices2joints(J_regressor, v_shaped)
    ident = torch.eye(3, dtype=dtype, device=betas.device)
    if pose2rot:
        rot_mats = batch_rodrigues(pose.reshape(-1, 3)).view(B, -1, 3, 3)
    else:
        rot_mats = pose.reshape(B, -1, 3, 3)
    pose_feature = (rot_mats[:, 1:, :, :] - ident).reshape(B, -1)
    pose_offsets = torch.matmul(pose_feature, posedirs).view(B, -1, 3)
    v_posed = pose_offsets + v_shaped
    J_transformed, A = batch_rigid_transform(rot_mats, J, parents)
    nj = J_regressor.shape[0]
    W = lbs_weights.unsqueeze(0).expand(B, -1, -1)
    T = torch.matmul(W, A.view(B, nj, 16)).view(B, -1, 4, 4)
    homo = torch.ones(B, v_posed.shape[1], 1, dtype=dtype, device=v_posed.device)
    v_posed_homo = torch.cat([v_posed, homo], dim=2)
    v_homo = torch.matmul(T, v_posed_homo.unsqueeze(-1))
    verts = v_homo[:, :, :3, 0]
    return verts, J_transformed, {'v_shaped': v_shaped, 'J': J, 'v_posed': v_posed, 'A': A,
                                  'rot_mats': rot_mats, 'pose_feature': pose_feature}


class SMPLOracle:
    """`pare.models.SMPL` semantics (commented twin: models/smpl.py:61-83) on top of smplx's
    SMPL.forward + VertexJointSelector(vertex_ids['smplh']):
       joints54 = [24 chain joints | verts[:, vertex_ids] (21) | J_regressor_extra.verts (9)]
       joints   = joints54[:, joint_map]  -> 49
    """

    def __init__(self, model, dtype=torch.float32, device='cpu'):
        """device: where the model tensors live ('cpu' for the oracle proper; bench.py's torch-GPU-eager leg passes
        'cuda' to time the same dense PyTorch path on the GPU)."""
        from_np = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dtype)  # noqa: E731
        self.dtype = dtype
        self.device = torch.device(device)
        self.v_template = from_np(model['v_template'])
        self.shapedirs = from_np(model['shapedirs'])
        self.posedirs = from_np(model['posedirs'])
        self.J_regressor = from_np(model['J_regressor'])
        self.lbs_weights = from_np(model['weights'])
        self.parents = torch.as_tensor(np.asarray(model['parents']), dtype=torch.long)   # host indices
        self.J_regressor_extra = from_np(model['J_regressor_extra'])
        self.vertex_ids = torch.as_tensor(np.asarray(model['vertex_ids']), dtype=torch.long).to(device)
        self.joint_map = torch.tensor(reference_tables()['joint_map_49'], dtype=torch.long).to(device)
        self.faces = model.get('f')

    def forward(self, betas, body_pose, global_orient, pose2rot=True, transl=None):
        betas = torch.as_tensor(betas).to(device=self.device, dtype=self.dtype)
        body_pose = torch.as_tensor(body_pose).to(device=self.device, dtype=self.dtype)
        global_orient = torch.as_tensor(global_orient).to(device=self.device, dtype=self.dtype)
        B = betas.shape[0]
        if pose2rot:
            full_pose = torch.cat([global_orient.reshape(B, 3), body_pose.reshape(B, -1)], dim=1)
        else:
            full_pose = torch.cat([global_orient.reshape(B, 1, 3, 3),
                                   body_pose.reshape(B, -1, 3, 3)], dim=1)
        verts, joints24, inter = lbs(betas, full_pose, self.v_template, self.shapedirs,
                                     self.posedirs, self.J_regressor, self.parents,
                                     self.lbs_weights, pose2rot=pose2rot)
        # smplx VertexJointSelector.forward: index_select + cat
        joints45 = torch.cat([joints24, torch.index_select(verts, 1, self.vertex_ids)], dim=1)
        if transl is not None:
            transl = torch.as_tensor(transl).to(device=self.device, dtype=self.dtype)
            joints45 = joints45 + transl.unsqueeze(1)
            verts = verts + transl.unsqueeze(1)
        # wrapper, models/smpl.py:71-83
        extra = vertices2joints(self.J_regressor_extra, verts)
        joints54 = torch.cat([joints45, extra], dim=1)
        joints = joints54[:, self.joint_map, :]
        out = {'vertices': verts, 'joints': joints, 'joints45': joints45, 'joints24': joints24,
               'global_orient': global_orient, 'body_pose': body_pose, 'betas': betas,
               'full_pose': full_pose}
        out.update(inter)
        return out

    __call__ = forward


_DEVICE_CACHE = {}
_TABLES = None


def reference_tables():
    """Index tables extracted from the reference's own module (models/smpl.py:14-58) by
    tests/golden/make_golden_tables.py: the 49-entry joint map, H36M_TO_J17 / J14.  The oracle reads THIS file, not the
    product package's constants, so a wrong entry in `whmr_b200.constants` cannot hide."""
    global _TABLES
    if _TABLES is None:
        import os
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'reference_tables.npz')
        _TABLES = dict(np.load(path))
    return _TABLES


def regressor_readouts(model, verts, dtype=None):
    """Everything Regressor.forward derives linearly from the posed vertices
    (models/whmr.py:176-187, 240-251): H36M joints (17 -> pelvis-centred 14), the dense
    Dmap0/Dmap1 downsample, the SSM markers and smpl_kp_3d (J_regressor on POSED verts +
    selected vertices).  Dense matmuls, exactly as the reference computes them."""
    h36m_to_j14 = [int(i) for i in reference_tables()['h36m_to_j14']]
    verts = torch.as_tensor(verts)
    dt = dtype or verts.dtype
    verts = verts.to(dt)
    dev = verts.device

    def t(k):   # on an accelerator the dense matrices are uploaded once (the reference keeps them as module buffers)
        if dev.type == 'cpu':
            return torch.from_numpy(np.ascontiguousarray(model[k])).to(dt)
        key = (id(model), k, str(dev), dt)
        if key not in _DEVICE_CACHE:
            _DEVICE_CACHE[key] = torch.from_numpy(np.ascontiguousarray(model[k])).to(device=dev, dtype=dt)
        return _DEVICE_CACHE[key]
    out = {}
    j17 = torch.matmul(t('J_regressor_h36m'), verts)               # whmr.py:177
    pelvis = j17[:, [0], :].clone()                                # :178
    out['h36m_j17'] = j17
    out['kp_3d_h36m'] = j17[:, h36m_to_j14, :] - pelvis    # :179-180
    out['sub_verts'] = torch.matmul(t('Dmap0'), verts)             # :182
    out['temp_verts'] = torch.matmul(t('Dmap1'), out['sub_verts'])  # :183
    out['markers'] = verts[:, torch.as_tensor(np.asarray(model['ssm']), dtype=torch.long).to(dev)]  # :184
    sj = vertices2joints(t('J_regressor'), verts)                  # :186
    vid = torch.as_tensor(np.asarray(model['vertex_ids']), dtype=torch.long).to(dev)
    out['smpl_kp_3d'] = torch.cat([sj, torch.index_select(verts, 1, vid)], dim=1)  # :187
    return out
