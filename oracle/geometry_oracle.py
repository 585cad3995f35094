"""torch-CPU restatement of the reference's projection / rotation helpers -- TEST
INFRASTRUCTURE.  Pinned against the reference's own functions via tests/golden/
(make_golden.py imports /root/reference/utils/geometry.py and models/maf_extractor.py).

Each function names the reference lines it follows.  Operation order is kept where it
affects fp32 rounding (translate -> divide by z -> intrinsics -> normalise).
"""
import torch

FOCAL_LENGTH = 1000.0   # core/constants.py:4
IMG_W = 256.0           # configs/pymaf_config.yaml:83-85 (cfg.IMG_RES.WIDTH/HEIGHT)
IMG_H = 256.0


def perspective_projection(points, rotation, translation, focal_length, camera_center,
                           retain_z=False):
    """utils/geometry.py:310-341.  points [B,N,3]; rotation [B or 1,3,3]; translation [B,3];
    focal_length scalar or [B]; camera_center [B,2].  -> [B,N,2] (or [B,N,3])."""
    B = points.shape[0]
    p = torch.einsum('bij,bkj->bki', rotation.to(points.dtype), points)          # :329
    p = p + translation.unsqueeze(1)                            # :330
    q = p / p[:, :, -1].unsqueeze(-1)                           # :333
    dt = points.dtype        # fp32 like the reference; float64 when a test apportions rounding error / takes gradients
    f = torch.as_tensor(focal_length, dtype=dt, device=points.device)
    f = f.expand(B) if f.dim() == 0 else f
    cc = torch.as_tensor(camera_center, dtype=dt, device=points.device)
    # K = [[f,0,cx],[0,f,cy],[0,0,1]]  (:322-326) applied as einsum (:336)
    u = f.view(B, 1) * q[:, :, 0] + cc[:, 0].view(B, 1) * q[:, :, 2]
    v = f.view(B, 1) * q[:, :, 1] + cc[:, 1].view(B, 1) * q[:, :, 2]
    out = torch.stack([u, v, q[:, :, 2]], dim=-1)
    return out if retain_z else out[:, :, :-1]


def projection(pred_joints, pred_camera, retain_z=False):
    """utils/geometry.py:289-307: weak-perspective camera (s,tx,ty) -> t = [tx, ty,
    2*1000/(256*s + 1e-9)], focal 1000, centre 0, R = I, then divide by (W/2, H/2)."""
    B = pred_joints.shape[0]
    t = torch.stack([pred_camera[:, 1], pred_camera[:, 2],
                     2 * FOCAL_LENGTH / (IMG_H * pred_camera[:, 0] + 1e-9)], dim=-1)
    eye = torch.eye(3, device=pred_joints.device, dtype=pred_joints.dtype).unsqueeze(0).expand(B, -1, -1)
    kp = perspective_projection(pred_joints, eye, t, FOCAL_LENGTH,
                                torch.zeros(B, 2, device=pred_joints.device, dtype=pred_joints.dtype), retain_z=retain_z)
    if retain_z:
        _retain_z_div(kp)
    return kp / (torch.tensor([IMG_W, IMG_H], device=pred_joints.device, dtype=pred_joints.dtype) / 2.)


def _retain_z_div(kp):
    # utils/geometry.py:303-304 divides a [B,N,3] tensor by a 2-vector when retain_z=True,
    # which raises in torch (broadcast 3 vs 2).  The reference never calls it that way
    # (all call sites use retain_z=False: models/whmr.py:143,237, models/maf_extractor.py:138).
    raise RuntimeError("projection(retain_z=True) is ill-formed in the reference "
                       "(utils/geometry.py:303-304 broadcasts [B,N,3] / [2])")


def convert_pare_to_full_img_cam(pare_cam, bbox_height, bbox_center, img_w, img_h,
                                 focal_length=None, Tz=None):
    """utils/geometry.py:139-157."""
    s, tx, ty = pare_cam[:, 0], pare_cam[:, 1], pare_cam[:, 2]
    tz = Tz if focal_length is None else 2 * focal_length / (bbox_height * s)
    cx = 2 * (bbox_center[:, 0] - (img_w / 2.)) / (s * bbox_height)
    cy = 2 * (bbox_center[:, 1] - (img_h / 2.)) / (s * bbox_height)
    return torch.stack([tx + cx, ty + cy, tz], dim=-1)


def full_projection(pred_joints, pred_cam, bbox_height, center, orig_shape, Tz):
    """The predicted-focal full-image projection block of Regressor.forward,
    models/whmr.py:147-173.  Returns (kp_2d_w_norm [B,N,2], focal_length [B], pred_cam_t [B,3],
    kp_2d_w_px [B,N,2])."""
    s = pred_cam[:, 0]
    focal_length = s * bbox_height * Tz / 2.                      # :149
    img_shape = orig_shape[:, [1, 0]]                             # :152  (w, h)
    camera_center = img_shape / 2.                                # :153
    pred_cam_t = convert_pare_to_full_img_cam(pred_cam, bbox_height, center,
                                              orig_shape[:, 1], orig_shape[:, 0], Tz=Tz)  # :154
    eye = torch.eye(3, device=pred_joints.device, dtype=pred_joints.dtype).unsqueeze(0).expand(1, -1, -1)
    kp_px = perspective_projection(pred_joints, eye.expand(pred_joints.shape[0], -1, -1),
                                   pred_cam_t, focal_length, camera_center)  # :165-171
    kp_norm = kp_px / camera_center.unsqueeze(1) - 1              # :173
    return kp_norm, focal_length, pred_cam_t, kp_px


def quat_to_rotmat(quat):
    """utils/geometry.py:31-51."""
    q = quat / quat.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    w2, x2, y2, z2 = w.pow(2), x.pow(2), y.pow(2), z.pow(2)
    wx, wy, wz = w * x, w * y, w * z
    xy, xz, yz = x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2],
                       dim=1).view(-1, 3, 3)


def batch_rodrigues_quat(theta):
    """utils/geometry.py:14-28 -- the quaternion variant (core/trainer.py:244); differs from
    the smplx variant inside SMPL.forward at ~1e-7."""
    l1 = torch.norm(theta + 1e-8, p=2, dim=1)
    angle = l1.unsqueeze(-1)
    n = theta / angle
    half = angle * 0.5
    return quat_to_rotmat(torch.cat([torch.cos(half), torch.sin(half) * n], dim=1))


def rot6d_to_rotmat(x):
    """utils/geometry.py:243-257."""
    x = x.reshape(-1, 3, 2)
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = torch.nn.functional.normalize(a1)
    b2 = torch.nn.functional.normalize(a2 - torch.einsum('bi,bi->b', b1, a2).unsqueeze(-1) * b1)
    b3 = torch.linalg.cross(b1, b2, dim=1)
    return torch.stack((b1, b2, b3), dim=-1)


def unbiased_gram_schmidt(x):
    """utils/geometry.py:260-272."""
    k = x.shape[1]
    x = x.reshape(-1, 3, 3)
    t1, t2, t3 = x[:, :, 0], x[:, :, 1], x[:, :, 2]
    nrm = torch.nn.functional.normalize
    r1 = nrm((torch.linalg.cross(t2, t3, dim=1) + t1) / 2.)
    r2_ = (torch.linalg.cross(t3, r1, dim=1) + t2) / 2.
    r2 = nrm(r2_ - (torch.einsum('bi,bi->b', r2_, r1).unsqueeze(-1) * r1))
    r3 = torch.linalg.cross(r1, r2, dim=1)
    return torch.stack((r1, r2, r3), dim=-1).reshape(-1, k, 3, 3)


def rotation_matrix_to_angle_axis(R):
    """utils/geometry.py:54-83 + :160-240 + :86-136 (kornia path: 3x3 -> quaternion by the
    4-case trace test with eps=1e-6 on the TRANSPOSED matrix -> angle-axis; NaNs -> 0)."""
    R = R.reshape(-1, 3, 3)
    m = R.transpose(1, 2)            # rmat_t
    eps = 1e-6
    d2 = m[:, 2, 2] < eps
    d0_d1 = m[:, 0, 0] > m[:, 1, 1]
    d0_nd1 = m[:, 0, 0] < -m[:, 1, 1]
    t0 = 1 + m[:, 0, 0] - m[:, 1, 1] - m[:, 2, 2]
    q0 = torch.stack([m[:, 1, 2] - m[:, 2, 1], t0, m[:, 0, 1] + m[:, 1, 0], m[:, 2, 0] + m[:, 0, 2]], -1)
    t1 = 1 - m[:, 0, 0] + m[:, 1, 1] - m[:, 2, 2]
    q1 = torch.stack([m[:, 2, 0] - m[:, 0, 2], m[:, 0, 1] + m[:, 1, 0], t1, m[:, 1, 2] + m[:, 2, 1]], -1)
    t2 = 1 - m[:, 0, 0] - m[:, 1, 1] + m[:, 2, 2]
    q2 = torch.stack([m[:, 0, 1] - m[:, 1, 0], m[:, 2, 0] + m[:, 0, 2], m[:, 1, 2] + m[:, 2, 1], t2], -1)
    t3 = 1 + m[:, 0, 0] + m[:, 1, 1] + m[:, 2, 2]
    q3 = torch.stack([t3, m[:, 1, 2] - m[:, 2, 1], m[:, 2, 0] - m[:, 0, 2], m[:, 0, 1] - m[:, 1, 0]], -1)
    c0 = (d2 & d0_d1).view(-1, 1).to(R.dtype)
    c1 = (d2 & ~d0_d1).view(-1, 1).to(R.dtype)
    c2 = (~d2 & d0_nd1).view(-1, 1).to(R.dtype)
    c3 = (~d2 & ~d0_nd1).view(-1, 1).to(R.dtype)
    q = q0 * c0 + q1 * c1 + q2 * c2 + q3 * c3
    q = q / torch.sqrt(t0.view(-1, 1) * c0 + t1.view(-1, 1) * c1 + t2.view(-1, 1) * c2 + t3.view(-1, 1) * c3)
    q = q * 0.5
    # quaternion_to_angle_axis, :86-136
    q1_, q2_, q3_ = q[:, 1], q[:, 2], q[:, 3]
    sin_sq = q1_ * q1_ + q2_ * q2_ + q3_ * q3_
    sin_t = torch.sqrt(sin_sq)
    cos_t = q[:, 0]
    two_theta = 2.0 * torch.where(cos_t < 0.0, torch.atan2(-sin_t, -cos_t), torch.atan2(sin_t, cos_t))
    k = torch.where(sin_sq > 0.0, two_theta / sin_t, 2.0 * torch.ones_like(sin_t))
    aa = torch.stack([q1_ * k, q2_ * k, q3_ * k], dim=-1)
    aa[torch.isnan(aa)] = 0.0
    return aa


# ---- MAF_Extractor.project family (defined in the reference, not called in the live path) ----

def maf_get_trans(pred_cam, center, scale, img_focal, img_center):
    """models/maf_extractor.py:175-190."""
    b = scale * 200
    s, tx, ty = pred_cam.unbind(-1)
    bs = b * s
    return torch.stack([tx + 2 * (center[:, 0] - img_center[:, 0]) / bs,
                        ty + 2 * (center[:, 1] - img_center[:, 1]) / bs,
                        2 * img_focal / bs], dim=-1).unsqueeze(1)


def maf_perspective_projection(points, focal_length, camera_center, distortion=None):
    """models/maf_extractor.py:192-235 with rotation=None, translation=None."""
    B = points.shape[0]
    if distortion is not None:
        kc = distortion
        p = points[:, :, :2] / points[:, :, 2:]
        r2 = p[:, :, 0] ** 2 + p[:, :, 1] ** 2
        dx = 2 * kc[:, [2]] * p[:, :, 0] * p[:, :, 1] + kc[:, [3]] * (r2 + 2 * p[:, :, 0] ** 2)
        dy = 2 * kc[:, [3]] * p[:, :, 0] * p[:, :, 1] + kc[:, [2]] * (r2 + 2 * p[:, :, 1] ** 2)
        rad = 1 + kc[:, [0]] * r2 + kc[:, [1]] * r2.pow(2) + kc[:, [4]] * r2.pow(3)
        points = torch.stack([rad * p[:, :, 0] + dx, rad * p[:, :, 1] + dy, torch.ones_like(r2)], dim=-1)
    q = points / points[:, :, -1].unsqueeze(-1)
    f = torch.as_tensor(focal_length, dtype=torch.float32)
    f = f.expand(B) if f.dim() == 0 else f
    u = f.view(B, 1) * q[:, :, 0] + camera_center[:, 0].view(B, 1) * q[:, :, 2]
    v = f.view(B, 1) * q[:, :, 1] + camera_center[:, 1].view(B, 1) * q[:, :, 2]
    return torch.stack([u, v], dim=-1)


def maf_project(points, pred_cam, center, scale, img_focal, img_center, crop_size=256.0,
                distortion=None):
    """models/maf_extractor.py:145-173: full-frame projection, then map into the crop and
    normalise to [-1,1].  Returns (points2d_full, points2d_crop_norm)."""
    trans_full = maf_get_trans(pred_cam, center, scale, img_focal, img_center)
    full = maf_perspective_projection(points + trans_full, img_focal, img_center, distortion)
    b = scale * 200
    p2 = full - (center - b[:, None] / 2)[:, None, :]
    p2 = p2 * (crop_size / b)[:, None, None]
    half = torch.tensor([IMG_W, IMG_H]) / 2.
    return full, (p2 - half) / half


def estimate_translation_np(S, joints_2d, joints_conf, focal_length=5000, img_size=(224., 224.)):
    """utils/geometry.py:344-383: weighted least squares for one sample; S [n,3], joints_2d [n,2], conf [n]."""
    import numpy as np
    n = S.shape[0]
    f = np.array([focal_length, focal_length], dtype=np.float64)
    center = np.array(img_size, dtype=np.float64) / 2.
    Z = np.repeat(S[:, 2], 2)                                   # :360
    XY = S[:, 0:2].reshape(-1)                                  # :361
    O = np.tile(center, n)                                      # :362
    F = np.tile(f, n)                                           # :363
    w2 = np.repeat(np.sqrt(joints_conf), 2)                     # :364
    Q = np.stack([F * np.tile([1., 0.], n), F * np.tile([0., 1.], n), O - joints_2d.reshape(-1)], axis=1)  # :367-368
    c = (joints_2d.reshape(-1) - O) * Z - F * XY                # :369
    Q = w2[:, None] * Q                                         # :372-374 (diagflat product)
    c = w2 * c
    return np.linalg.solve(Q.T @ Q, Q.T @ c)                    # :377-381


def estimate_translation(S, joints_2d, focal_length=5000., img_size=(224., 224.)):
    """utils/geometry.py:386-408: joints 25: only, one solve per sample, float32 result."""
    import numpy as np
    S = np.asarray(S)[:, 25:, :]
    j2 = np.asarray(joints_2d)[:, 25:, :]
    out = np.zeros((S.shape[0], 3), dtype=np.float32)
    for i in range(S.shape[0]):
        out[i] = estimate_translation_np(S[i], j2[i, :, :2], j2[i, :, 2], focal_length, img_size)
    return out
