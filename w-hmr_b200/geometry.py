"""Drop-ins for the projection functions of the reference's utils/geometry.py, imported by name at
models/whmr.py:24-25, models/maf_extractor.py:10, core/trainer.py:27.  Same names, positional and
keyword use, argument meaning.  The projections are CUDA-only torch.library ops WITH autograd (the reference's
training graph back-propagates through them: kp_2d_w -> joints / Tz, core/trainer.py:518).

The three rotation helpers are also called by the reference at CONSTRUCTION time on CPU tensors
(`rot6d_to_rotmat(init_pose)` in Regressor.__init__ / Global_Orient_Regressor.__init__, models/whmr.py:65,285) and,
in training, under autograd.  Those two uses -- not the hot path -- take a small differentiable torch implementation
(`_torch_*` below); CUDA tensors that need no gradient take the sm_100a kernels."""
import torch

from . import constants, ops


def _use_torch(x):
    """init-time glue on CPU tensors, or a caller that differentiates through the helper"""
    return (not x.is_cuda) or (torch.is_grad_enabled() and x.requires_grad)


def _unit(v):
    return v / v.norm(dim=-1, keepdim=True).clamp_min(1e-12)


def _torch_rot6d_to_rotmat(x):
    a = x.reshape(-1, 3, 2)
    b1 = _unit(a[..., 0])
    b2 = _unit(a[..., 1] - (b1 * a[..., 1]).sum(-1, keepdim=True) * b1)
    return torch.stack((b1, b2, torch.cross(b1, b2, dim=-1)), dim=-1)


def _torch_unbiased_gram_schmidt(x):
    t1, t2, t3 = x[..., 0], x[..., 1], x[..., 2]          # columns
    r1 = _unit((torch.cross(t2, t3, dim=-1) + t1) / 2.0)
    q = (torch.cross(t3, r1, dim=-1) + t2) / 2.0
    r2 = _unit(q - (q * r1).sum(-1, keepdim=True) * r1)
    return torch.stack((r1, r2, torch.cross(r1, r2, dim=-1)), dim=-1)


def _torch_rotmat_to_angle_axis(R):
    m = R.transpose(-1, -2)                                 # the reference works on R^T
    m00, m01, m02 = m[:, 0, 0], m[:, 0, 1], m[:, 0, 2]
    m10, m11, m12 = m[:, 1, 0], m[:, 1, 1], m[:, 1, 2]
    m20, m21, m22 = m[:, 2, 0], m[:, 2, 1], m[:, 2, 2]
    t0, t1 = 1 + m00 - m11 - m22, 1 - m00 + m11 - m22
    t2, t3 = 1 - m00 - m11 + m22, 1 + m00 + m11 + m22
    q_a = torch.stack((m12 - m21, t0, m01 + m10, m20 + m02), -1)
    q_b = torch.stack((m20 - m02, m01 + m10, t1, m12 + m21), -1)
    q_c = torch.stack((m01 - m10, m20 + m02, m12 + m21, t2), -1)
    q_d = torch.stack((t3, m12 - m21, m20 - m02, m01 - m10), -1)
    d2, d01, d0n1 = m22 < 1e-6, m00 > m11, m00 < -m11
    ca, cb, cc = d2 & d01, d2 & ~d01, ~d2 & d0n1
    sel = lambda a, b, c, d: torch.where(ca.unsqueeze(-1) if a.dim() > 1 else ca, a,   # noqa: E731
                                         torch.where(cb.unsqueeze(-1) if a.dim() > 1 else cb, b,
                                                     torch.where(cc.unsqueeze(-1) if a.dim() > 1 else cc, c, d)))
    q = sel(q_a, q_b, q_c, q_d) * (0.5 / torch.sqrt(sel(t0, t1, t2, t3))).unsqueeze(-1)
    ss = (q[:, 1:] ** 2).sum(-1)
    st = torch.sqrt(ss)
    two_theta = 2.0 * torch.where(q[:, 0] < 0.0, torch.atan2(-st, -q[:, 0]), torch.atan2(st, q[:, 0]))
    k = torch.where(ss > 0.0, two_theta / st, torch.full_like(st, 2.0))
    aa = q[:, 1:] * k.unsqueeze(-1)
    return torch.where(torch.isnan(aa), torch.zeros_like(aa), aa)


def projection(pred_joints, pred_camera, retain_z=False):
    """utils/geometry.py:289-307.  [B,N,3], [B,3] (s,tx,ty) -> [B,N,2] in ~[-1,1].
    focal 1000 (core/constants.py:4), 256x256 crop (cfg.IMG_RES), principal point 0, R = I."""
    if retain_z:
        # the reference divides a [B,N,3] tensor by a 2-vector on this branch and raises
        # (utils/geometry.py:303-304); no call site uses it.  Same error behaviour here.
        raise RuntimeError("The size of tensor a (3) must match the size of tensor b (2) at non-singleton "
                           "dimension 2")
    return ops.project_weak_op(pred_joints, pred_camera, constants.FOCAL_LENGTH,
                               float(constants.IMG_RES_WIDTH), float(constants.IMG_RES_HEIGHT))


def perspective_projection(points, rotation, translation, focal_length, camera_center, retain_z=False):
    """utils/geometry.py:310-341.  rotation [B,3,3] or [1,3,3] (models/whmr.py:157-163 passes an
    expanded eye); focal_length scalar or [B]; camera_center [B,2]."""
    dev = points.device
    none = points.new_empty(0)
    if not torch.is_tensor(camera_center):
        camera_center = torch.as_tensor(camera_center, dtype=torch.float32)
    camera_center = camera_center.to(dev)
    focal = focal_length if torch.is_tensor(focal_length) else torch.tensor([float(focal_length)], device=dev)
    focal = focal.to(dev).reshape(-1) if focal.numel() > 1 else focal.to(dev).reshape(1)
    if rotation is not None and rotation.dim() == 2:
        rotation = rotation.unsqueeze(0)
    return ops.perspective_projection_op(points, none if rotation is None else rotation,
                                         none if translation is None else translation, focal, camera_center,
                                         bool(retain_z))


def convert_pare_to_full_img_cam(pare_cam, bbox_height, bbox_center, img_w, img_h, focal_length=None, Tz=None):
    """utils/geometry.py:139-157 -> cam_t [B,3].  Computed by the fused full-projection kernel."""
    B = pare_cam.shape[0]
    if focal_length is not None:
        Tz = 2 * focal_length / (bbox_height * pare_cam[:, 0])
    if not torch.is_tensor(Tz):
        Tz = torch.full((B,), float(Tz), device=pare_cam.device)
    as_vec = lambda x: x if torch.is_tensor(x) else torch.full((B,), float(x), device=pare_cam.device)  # noqa: E731
    orig_shape = torch.stack([as_vec(img_h), as_vec(img_w)], dim=-1)
    dummy = torch.zeros(B, 1, 3, device=pare_cam.device)
    # the weak + full projection op carries the autograd of cam_t w.r.t. pare_cam and Tz (models/whmr.py:154-155 feeds a
    # detached camera, so in the reference's graph the gradient reaches Tz only)
    _, _, _, cam_t = ops.project_weak_full_op(dummy, pare_cam, bbox_height, bbox_center, orig_shape, Tz,
                                              constants.FOCAL_LENGTH, float(constants.IMG_RES_WIDTH),
                                              float(constants.IMG_RES_HEIGHT))
    return cam_t


def full_image_projection(pred_joints, pred_cam, bbox_height, center, orig_shape, Tz, want_px=False):
    """The whole predicted-focal block of Regressor.forward (models/whmr.py:147-173) in one launch:
    -> kp_2d_w (normalised) [B,N,2], focal_length [B], pred_cam_t [B,3], (pixels [B,N,2])."""
    return ops.project_full(pred_joints, pred_cam, bbox_height, center, orig_shape, Tz, want_px=want_px)


def batch_rodrigues(theta):
    """utils/geometry.py:14-51 -- what `from utils.geometry import batch_rodrigues` gives the trainer (core/trainer.py:244):
    the QUATERNION variant (half-angle -> quaternion -> normalise -> matrix).  [N,3] -> [N,3,3]."""
    return ops.batch_rodrigues_quat(theta)


def batch_rodrigues_smplx(rot_vecs):
    """smplx.lbs.batch_rodrigues (imported at models/whmr.py:9), the variant inside SMPL.forward (differs from the
    quaternion one at ~1e-7)."""
    return ops.batch_rodrigues(rot_vecs)


def rot6d_to_rotmat(x):
    """utils/geometry.py:243-257."""
    return _torch_rot6d_to_rotmat(x) if _use_torch(x) else ops.rot6d_to_rotmat(x)


def unbiased_gram_schmidt(x):
    """utils/geometry.py:260-272 (models/whmr.py:129-130)."""
    return _torch_unbiased_gram_schmidt(x) if _use_torch(x) else ops.unbiased_gram_schmidt(x)


def rotation_matrix_to_angle_axis(rotation_matrix):
    """utils/geometry.py:54-83 for [N,3,3] input (the only form the reference's hot path uses, models/whmr.py:174)."""
    if rotation_matrix.shape[-2:] != (3, 3):
        raise ValueError("rotation_matrix_to_angle_axis: expected [N,3,3], got %s" % (tuple(rotation_matrix.shape),))
    if _use_torch(rotation_matrix):
        return _torch_rotmat_to_angle_axis(rotation_matrix)
    return ops.rotation_matrix_to_angle_axis(rotation_matrix)


def estimate_translation(S, joints_2d, focal_length=5000., img_size=[224., 224.]):
    """utils/geometry.py:386-408 (core/trainer.py:435): S [B,49,3], joints_2d [B,49,3] = (x, y, conf); joints 25:
    (the ground-truth set) take part.  The reference round-trips through NumPy on the host with one
    np.linalg.solve per sample; this stays on the device and on the current stream."""
    return ops.estimate_translation(S, joints_2d, focal_length, img_size, first_joint=25)
