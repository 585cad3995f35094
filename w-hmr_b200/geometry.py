"""Drop-ins for the projection functions of the reference's utils/geometry.py, imported by name at
models/whmr.py:24-25, models/maf_extractor.py:10, core/trainer.py:27.  Same names, positional and
keyword use, argument meaning; CUDA tensors only (no fallback)."""
import torch

from . import constants, ops


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
    return ops.perspective_projection(points, rotation, translation, focal_length, camera_center, retain_z)


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
    _, _, cam_t, _ = ops.project_full(dummy, pare_cam, bbox_height, bbox_center, orig_shape, Tz)
    return cam_t


def full_image_projection(pred_joints, pred_cam, bbox_height, center, orig_shape, Tz, want_px=False):
    """The whole predicted-focal block of Regressor.forward (models/whmr.py:147-173) in one launch:
    -> kp_2d_w (normalised) [B,N,2], focal_length [B], pred_cam_t [B,3], (pixels [B,N,2])."""
    return ops.project_full(pred_joints, pred_cam, bbox_height, center, orig_shape, Tz, want_px=want_px)


def batch_rodrigues(rot_vecs):
    """smplx.lbs.batch_rodrigues (imported at models/whmr.py:9), the variant inside SMPL.forward."""
    return ops.batch_rodrigues(rot_vecs)


def rot6d_to_rotmat(x):
    """utils/geometry.py:243-257."""
    return ops.rot6d_to_rotmat(x)


def unbiased_gram_schmidt(x):
    """utils/geometry.py:260-272 (models/whmr.py:129-130)."""
    return ops.unbiased_gram_schmidt(x)


def rotation_matrix_to_angle_axis(rotation_matrix):
    """utils/geometry.py:54-83 for [N,3,3] input (the only form the reference's hot path uses, models/whmr.py:174)."""
    if rotation_matrix.shape[-2:] != (3, 3):
        raise ValueError("rotation_matrix_to_angle_axis: expected [N,3,3], got %s" % (tuple(rotation_matrix.shape),))
    return ops.rotation_matrix_to_angle_axis(rotation_matrix)


def estimate_translation(S, joints_2d, focal_length=5000., img_size=[224., 224.]):
    """utils/geometry.py:386-408 (core/trainer.py:435): S [B,49,3], joints_2d [B,49,3] = (x, y, conf); joints 25:
    (the ground-truth set) take part.  The reference round-trips through NumPy on the host with one
    np.linalg.solve per sample; this stays on the device and on the current stream."""
    return ops.estimate_translation(S, joints_2d, focal_length, img_size, first_joint=25)
