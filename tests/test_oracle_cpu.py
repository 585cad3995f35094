"""The oracle against (a) outputs of the reference's own code (tests/golden/reference_outputs.npz,
made by tests/golden/make_golden.py), (b) the reference's vendored known-answer tests, and
(c) itself: the two independent SMPL restatements + analytic invariants (SMPL is parity
unpinned -- see oracle/__init__.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import geometry_oracle as G
from oracle import metrics_oracle as M
from oracle import sampling_oracle as S
from oracle import smpl_oracle, smpl_webuser_oracle
import whmr_b200.synthetic as syn

T = lambda a: torch.from_numpy(np.asarray(a))  # noqa: E731


def close(a, b, atol, rtol=0.0):
    a = a.numpy() if torch.is_tensor(a) else np.asarray(a)
    np.testing.assert_allclose(a, np.asarray(b), atol=atol, rtol=rtol)


# ------------------------------------------------------------------ pinned by the reference's code
def test_projection_matches_reference(golden):
    close(G.projection(T(golden['proj_points']), T(golden['proj_cam'])), golden['proj_out'], 1e-6, 1e-6)


def test_full_projection_matches_reference(golden):
    g = golden
    kp_norm, focal, cam_t, kp_px = G.full_projection(
        T(g['proj_points']), T(g['proj_cam']), T(g['full_bbox_h']), T(g['full_center']),
        T(g['full_orig_shape']), T(g['full_Tz']))
    close(focal, g['full_focal'], 0, 1e-6)
    close(cam_t, g['full_cam_t'], 1e-6, 1e-6)
    close(kp_px, g['full_kp_px'], 2e-3, 1e-6)
    close(kp_norm, g['full_kp_norm'], 2e-6, 1e-6)
    close(G.convert_pare_to_full_img_cam(T(g['proj_cam']), T(g['full_bbox_h']), T(g['full_center']),
                                         T(g['full_orig_shape'])[:, 1], T(g['full_orig_shape'])[:, 0],
                                         focal_length=5000.), g['full_cam_t_f5000'], 1e-6, 1e-6)
    close(G.perspective_projection(T(g['proj_points']), T(g['pp_rot']), T(g['full_cam_t']),
                                   T(g['full_focal']), T(g['full_orig_shape'])[:, [1, 0]] / 2., retain_z=True),
          g['pp_retain_z'], 2e-3, 1e-6)


def test_rotation_helpers_match_reference(golden):
    g = golden
    close(G.batch_rodrigues_quat(T(g['rodq_in'])), g['rodq_out'], 1e-6)
    close(G.rot6d_to_rotmat(T(g['rot6d_in'])), g['rot6d_out'], 1e-6)
    close(G.unbiased_gram_schmidt(T(g['ugs_in'])), g['ugs_out'], 1e-6)
    close(G.rotation_matrix_to_angle_axis(T(g['r2aa_in'])), g['r2aa_out'], 1e-5)


def test_rodrigues_identity_kat(golden):
    # models/ViTPose/tests/test_losses/test_mesh_losses.py:24-25: theta = 0 -> identity
    R = G.batch_rodrigues_quat(torch.zeros(24, 3))
    close(R, np.broadcast_to(np.eye(3, dtype=np.float32), (24, 3, 3)), 1e-7)
    R2 = smpl_oracle.batch_rodrigues(torch.zeros(24, 3))
    close(R2, np.broadcast_to(np.eye(3, dtype=np.float32), (24, 3, 3)), 1e-7)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_sampling_matches_reference(golden, tag):
    g = golden
    pf = S.grid_sample_points(T(g['samp_%s_feat' % tag]), T(g['samp_%s_points' % tag]))
    close(pf, g['samp_%s_point_feat' % tag], 1e-6)
    hand = S.bilinear_points_np(g['samp_%s_feat' % tag], g['samp_%s_points' % tag])
    close(hand, g['samp_%s_point_feat' % tag], 2e-6)
    convs = [(T(g['maf_conv%d_weight' % i]), T(g['maf_conv%d_bias' % i])) for i in range(3)]
    close(S.reduce_dim(pf, convs), g['samp_%s_mesh_align' % tag], 1e-5)


def test_maf_forward_and_project_match_reference(golden):
    g = golden
    p2d = G.projection(T(g['fwd_p']), T(g['fwd_cam']))
    pf = S.grid_sample_points(T(g['fwd_feat']), p2d)
    close(pf, g['fwd_point_feat'], 1e-5)
    full, crop = G.maf_project(T(g['fwd_p']), T(g['fwd_cam']), T(g['mproj_center']), T(g['mproj_scale']),
                               T(g['mproj_focal']), T(g['mproj_img_center']))
    close(full, g['mproj_full'], 2e-3, 1e-6)
    close(crop, g['mproj_crop'], 1e-5, 1e-6)
    tr = G.maf_get_trans(T(g['fwd_cam']), T(g['mproj_center']), T(g['mproj_scale']), T(g['mproj_focal']),
                         T(g['mproj_img_center']))
    close(G.maf_perspective_projection(T(g['fwd_p']) + tr, T(g['mproj_focal']), T(g['mproj_img_center']),
                                       distortion=T(g['mproj_kc'])), g['mproj_distorted'], 2e-3, 1e-6)


def test_procrustes_matches_reference_and_kat(golden):
    g = golden
    hat = M.compute_similarity_transform_batch(g['pa_S1'], g['pa_S2'])
    close(hat, g['pa_S1_hat'], 1e-5)
    close(M.pa_mpjpe(g['pa_S1'], g['pa_S2']), g['pa_err'], 1e-5)
    # models/ViTPose/tests/test_evaluation/test_mesh_eval.py:8-14
    rng = np.random.default_rng(5)
    src = rng.random((14, 3)); tgt = src * 0.5 + rng.random((1, 3))
    np.testing.assert_array_almost_equal(M.compute_similarity_transform(src, tgt), tgt)


# ------------------------------------------------------------------ SMPL: pinned to the reference's in-tree code
def _smpl_golden():
    import hashlib
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "smpl_webuser_outputs.npz"))
    model = syn.make_smpl_model(seed=int(g['model_seed']), weights="random")
    h = hashlib.sha256()
    for k in ('v_template', 'shapedirs', 'posedirs', 'weights', 'J_regressor', 'parents'):
        h.update(np.ascontiguousarray(model[k]).tobytes())
    assert h.hexdigest() == str(g['model_sha256']), "synthetic model generator drifted from the one the goldens used"
    return g, model


def test_estimate_translation_matches_reference(golden):
    """utils/geometry.py:344-408 restated vs outputs of the reference's own function."""
    from oracle import geometry_oracle as G
    S, j2 = golden['et_S'], golden['et_joints_2d']
    np.testing.assert_allclose(G.estimate_translation(S, j2, 5000., (224., 224.)), golden['et_out'], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(G.estimate_translation(S, j2, 1000., (256., 192.)), golden['et_out_f1000'], rtol=2e-6,
                               atol=2e-6)


def test_smpl_oracle_matches_reference_smpl_webuser_outputs():
    """Vertices and posed joints produced by EXECUTING the reference's models/smpl_webuser code
    (tests/golden/make_golden_smpl.py) pin both restatements: the batched smplx-style oracle (fp64 and fp32)
    and the per-body NumPy one."""
    g, model = _smpl_golden()
    pose, betas = g['pose'], g['betas']
    n = pose.shape[0]
    o64 = smpl_oracle.SMPLOracle(model, torch.float64)
    r = o64(betas.astype(np.float64), pose[:, 3:].astype(np.float64), pose[:, :3].astype(np.float64), pose2rot=True)
    # 2e-7: float32 storage of the goldens (6e-8 at 1 m) + smplx's angle = ||theta + 1e-8|| vs exact cv2.Rodrigues
    close(r['vertices'], g['verts'], 2e-7)
    close(r['joints24'], g['Jtr'], 2e-7)
    o32 = smpl_oracle.SMPLOracle(model, torch.float32)
    r32 = o32(betas, pose[:, 3:], pose[:, :3], pose2rot=True)
    close(r32['vertices'], g['verts'], 2e-6)
    close(r32['joints24'], g['Jtr'], 2e-6)
    for i in (0, 8, n - 1):
        v, jtr = smpl_webuser_oracle.smpl_body(model, pose[i], betas[i])
        close(v, g['verts'][i], 1e-7)
        close(jtr, g['Jtr'][i], 1e-7)


# ------------------------------------------------------------------ SMPL: the two restatements against each other
def _bodies(n):
    b = syn.make_bodies(n, seed=1)
    return b


@pytest.mark.parametrize("weights", ["random", "skeleton"])
def test_smpl_two_restatements_agree(weights):
    model = syn.make_smpl_model(seed=0 if weights == "random" else 3, weights=weights)
    b = _bodies(12)
    o64 = smpl_oracle.SMPLOracle(model, torch.float64)
    r = o64(b['betas'], b['pose_aa'][:, 3:], b['pose_aa'][:, :3], pose2rot=True)
    for i in range(12):
        v, jtr = smpl_webuser_oracle.smpl_body(model, b['pose_aa'][i], b['betas'][i])
        # not 1e-12: smplx's angle = ||theta + 1e-8|| perturbs every rotation by O(1e-8) rad
        # relative to the exact cv2.Rodrigues the smpl_webuser code uses
        close(r['vertices'][i], v, 5e-8)
        close(r['joints24'][i], jtr, 5e-8)
    # float32 path of the same oracle stays within 1e-6 m of float64 (budget for the 1e-5 m gate)
    o32 = smpl_oracle.SMPLOracle(model, torch.float32)
    r32 = o32(b['betas'], b['pose_aa'][:, 3:], b['pose_aa'][:, :3], pose2rot=True)
    assert (r32['vertices'].double() - r['vertices']).abs().max() < 2e-6
    assert (r32['joints'].double() - r['joints']).abs().max() < 2e-6
    # rotmat mode == axis-angle mode
    rr = o64(b['betas'], b['rotmat'][:, 1:].astype(np.float64), b['rotmat'][:, :1].astype(np.float64),
             pose2rot=False)
    assert (rr['vertices'] - r['vertices']).abs().max() < 1e-5   # rotmat inputs are fp32-rounded


def test_smpl_invariants(smpl_model):
    o = smpl_oracle.SMPLOracle(smpl_model, torch.float64)
    b = _bodies(6)
    B = 6
    zg = np.zeros((B, 3))
    eye = np.broadcast_to(np.eye(3), (B, 24, 3, 3)).copy()
    r0 = o(b['betas'], eye[:, 1:], eye[:, :1], pose2rot=False)
    # identity pose => verts = v_shaped ; chain joints = J_regressor . v_shaped
    # (1e-7, not 1e-12: the fp32 skinning weights sum to 1 only to ~6e-8)
    assert (r0['vertices'] - r0['v_shaped']).abs().max() < 1e-7
    assert (r0['joints24'] - r0['J']).abs().max() < 1e-12
    # axis-angle zero: the 1e-8 only enters the norm, rot_dir = 0/angle = 0 => exactly identity
    rz = o(b['betas'], np.zeros((B, 69)), zg, pose2rot=True)
    assert (rz['vertices'] - r0['vertices']).abs().max() == 0
    # global rotation equivariance about the root joint
    r = o(b['betas'], b['pose_aa'][:, 3:], zg, pose2rot=True)
    rg = o(b['betas'], b['pose_aa'][:, 3:], b['pose_aa'][:, :3], pose2rot=True)
    R = torch.from_numpy(syn.rodrigues_np(b['pose_aa'][:, :3]))
    root = r['J'][:, :1]
    expect = torch.einsum('bij,bvj->bvi', R, r['vertices'] - root) + root
    assert (rg['vertices'] - expect).abs().max() < 1e-7   # weights sum to 1 only to ~6e-8
    # joints output: 49 entries assembled per models/smpl.py:66-76
    assert rg['joints'].shape == (B, 49, 3)
    assert torch.equal(rg['joints'][:, 8], rg['joints24'][:, 0])          # 'OP MidHip' -> joint 0
    vid = smpl_model['vertex_ids']
    assert torch.equal(rg['joints'][:, 0], rg['vertices'][:, vid[0]])      # 'OP Nose' -> 24 -> nose vertex


def test_rigid_vertex_moves_with_its_joint():
    model = syn.make_smpl_model(seed=0, weights="random")
    model['weights'][:] = 0
    model['weights'][:, 5] = 1.0          # every vertex bound to joint 5
    model['posedirs'][:] = 0
    o = smpl_oracle.SMPLOracle(model, torch.float64)
    b = _bodies(3)
    r = o(b['betas'], b['pose_aa'][:, 3:], b['pose_aa'][:, :3], pose2rot=True)
    A5 = r['A'][:, 5]
    vh = torch.cat([r['v_shaped'], torch.ones(3, 6890, 1, dtype=torch.float64)], -1)
    expect = torch.einsum('bij,bvj->bvi', A5, vh)[..., :3]
    assert (r['vertices'] - expect).abs().max() < 1e-12


def test_pkl_schema_roundtrip(tmp_path, smpl_model):
    # schema of models/ViTPose/tests/utils/mesh_utils.py:19-27
    p = syn.write_smpl_pkl(smpl_model, str(tmp_path))
    import pickle
    d = pickle.load(open(p, 'rb'))
    assert d['posedirs'].shape == (6890, 3, 207) and d['shapedirs'].shape == (6890, 3, 10)
    assert d['J_regressor'].shape == (24, 6890) and d['kintree_table'].shape == (2, 24)
    m2 = syn.load_smpl_pkl(p)
    for k in ('v_template', 'shapedirs', 'posedirs', 'J_regressor', 'weights'):
        np.testing.assert_array_equal(m2[k], smpl_model[k])
    assert list(m2['parents']) == list(smpl_model['parents'])


def test_index_tables_equal_reference_module():
    """whmr_b200.constants against the tables extracted by importing the reference's models/smpl.py
    (tests/golden/make_golden_tables.py): every entry of the 49-joint map, the joint names, H36M_TO_J17 / J14, the focal."""
    import numpy as np
    from oracle.smpl_oracle import reference_tables
    from whmr_b200 import constants as c
    t = reference_tables()
    assert list(c.JOINT_MAP_49) == [int(i) for i in t['joint_map_49']]
    assert list(c.JOINT_NAMES) == [str(n) for n in t['joint_names']]
    assert dict(c.JOINT_MAP) == {str(k): int(v) for k, v in zip(t['joint_map_keys'], t['joint_map_vals'])}
    assert list(c.H36M_TO_J17) == [int(i) for i in t['h36m_to_j17']]
    assert list(c.H36M_TO_J14) == [int(i) for i in t['h36m_to_j14']]
    assert float(c.FOCAL_LENGTH) == float(t['focal_length'])
    assert len(set(np.asarray(t['joint_map_49']).tolist())) <= 49 and int(np.max(t['joint_map_49'])) == 53


def test_geometry_glue_on_cpu_matches_reference_golden(golden):
    """The reference calls rot6d_to_rotmat on CPU tensors while it builds the model (models/whmr.py:65,285), so the drop-in
    module must serve CPU (and autograd) callers: its torch path against the reference's own outputs."""
    import torch
    from whmr_b200 import geometry as geo
    T = lambda k: torch.from_numpy(golden[k])  # noqa: E731
    assert float((geo.rot6d_to_rotmat(T('rot6d_in')) - T('rot6d_out')).abs().max()) <= 1e-6
    assert float((geo.unbiased_gram_schmidt(T('ugs_in')) - T('ugs_out')).abs().max()) <= 1e-6
    assert float((geo.rotation_matrix_to_angle_axis(T('r2aa_in')) - T('r2aa_out')).abs().max()) <= 2e-6
    x = T('ugs_in').clone().requires_grad_(True)
    geo.rotation_matrix_to_angle_axis(geo.unbiased_gram_schmidt(x).reshape(-1, 3, 3)).sum().backward()
    assert x.grad is not None and bool(torch.isfinite(x.grad).all())
