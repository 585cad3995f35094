"""Evaluation-metric oracle (BASELINE config 5) -- TEST INFRASTRUCTURE.

Follows utils/pose_utils.py:10-75 (compute_similarity_transform[_batch], reconstruction_error)
and the MPJPE / PVE lines of evaluate/eval.py:208-223.  Pinned by the reference's vendored
known-answer test models/ViTPose/tests/test_evaluation/test_mesh_eval.py:8-14
(target = 0.5*source + t  =>  aligned == target to 6 decimals).
"""
import numpy as np


def compute_similarity_transform(S1, S2):
    """utils/pose_utils.py:10-58.  S1,S2 [N,3] -> S1 aligned onto S2 (float64 on host)."""
    S1 = np.asarray(S1, dtype=np.float64).T
    S2 = np.asarray(S2, dtype=np.float64).T
    mu1 = S1.mean(axis=1, keepdims=True)
    mu2 = S2.mean(axis=1, keepdims=True)
    X1, X2 = S1 - mu1, S2 - mu2
    var1 = np.sum(X1 ** 2)
    K = X1.dot(X2.T)
    U, s, Vh = np.linalg.svd(K)
    V = Vh.T
    Z = np.eye(3)
    Z[-1, -1] *= np.sign(np.linalg.det(U.dot(V.T)))
    R = V.dot(Z.dot(U.T))
    scale = np.trace(R.dot(K)) / var1
    t = mu2 - scale * (R.dot(mu1))
    return (scale * R.dot(S1) + t).T


def compute_similarity_transform_batch(S1, S2):
    return np.stack([compute_similarity_transform(a, b) for a, b in zip(S1, S2)])


def pa_mpjpe(pred, gt):
    """utils/pose_utils.py:67-75 with reduction=None: per-sample mean joint error after alignment."""
    hat = compute_similarity_transform_batch(pred, gt)
    return np.sqrt(((hat - np.asarray(gt, dtype=np.float64)) ** 2).sum(-1)).mean(-1)


def mpjpe(pred, gt):
    """evaluate/eval.py:222: sqrt(sum((pred-gt)^2, -1)).mean(-1) per sample."""
    d = np.asarray(pred, dtype=np.float64) - np.asarray(gt, dtype=np.float64)
    return np.sqrt((d ** 2).sum(-1)).mean(-1)


def pve(pred_verts, gt_verts):
    """evaluate/eval.py:208-209: per-frame mean vertex distance."""
    d = np.asarray(pred_verts, dtype=np.float64) - np.asarray(gt_verts, dtype=np.float64)
    return np.sqrt((d ** 2).sum(-1)).mean(-1)


def eval_pass(model, gt_pose, gt_betas, pred_rotmat, pred_betas, joint_mapper=None):
    """evaluate/eval.py:157-223 restated on the CPU oracle: GT SMPL from axis-angle, predicted SMPL from rotation
    matrices, J_regressor_h36m . vertices -> joint_mapper -> minus the regressed pelvis (:198-219), then MPJPE
    (:222), PA-MPJPE (utils/pose_utils.py:67-75) and PVE (:208).  Returns per-frame errors in metres."""
    import torch
    from .smpl_oracle import SMPLOracle
    H36M_TO_J17 = [6, 5, 4, 1, 2, 3, 16, 15, 14, 11, 12, 13, 8, 10, 0, 7, 9]   # models/smpl.py:57
    mapper = H36M_TO_J17[:14] if joint_mapper is None else list(joint_mapper)   # H36M_TO_J14, models/smpl.py:58
    o = SMPLOracle(model, torch.float32)
    gt = o(gt_betas, gt_pose[:, 3:], gt_pose[:, :3], pose2rot=True)
    pr = o(pred_betas, pred_rotmat[:, 1:], pred_rotmat[:, :1], pose2rot=False)
    Jr = torch.from_numpy(np.asarray(model['J_regressor_h36m'], dtype=np.float32))

    def kp(v):
        j = torch.matmul(Jr[None].expand(v.shape[0], -1, -1), v)
        return (j[:, mapper] - j[:, [0]]).numpy()

    gk, pk = kp(gt['vertices']), kp(pr['vertices'])
    return {'mpjpe': mpjpe(pk, gk), 'pa_mpjpe': pa_mpjpe(pk, gk),
            'pve': pve(pr['vertices'].numpy(), gt['vertices'].numpy())}
