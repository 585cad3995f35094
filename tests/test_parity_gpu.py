"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the reference-generated
golden fixtures.  Tolerances are the north star's: vertices/joints <= 1e-5 m max-abs, projected
keypoints <= 1e-3 px, sampled features <= 1e-4 relative."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

VERT_TOL = 1e-5      # metres
PX_TOL = 1e-3        # pixels
FEAT_RTOL = 1e-4     # relative


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch.device("cuda:0")


def _oracle(model, dtype=torch.float32):
    from oracle.smpl_oracle import SMPLOracle
    return SMPLOracle(model, dtype)


def _smpl(model, dev, gemm_mode):
    from whmr_b200.smpl import SMPL
    return SMPL(model=model, gemm_mode=gemm_mode).to(dev)


def _bodies(B, seed=1):
    import whmr_b200.synthetic as syn
    return syn.make_bodies(B, seed=seed)


def _maxabs(a, b):
    return float((a.detach().double().cpu() - torch.as_tensor(b).double()).abs().max())


GEMM_MODES = ["fp32_simt", "bf16x3", "3xtf32"]


# ------------------------------------------------------------------------------------------ SMPL
@pytest.mark.parametrize("gemm_mode", GEMM_MODES)
@pytest.mark.parametrize("weights", ["random", "skeleton", "dense"])
def test_smpl_forward_config1(dev, weights, gemm_mode):
    """BASELINE config 1: B=64, axis-angle and rotmat modes, vs the fp32 and fp64 oracle."""
    import whmr_b200.synthetic as syn
    model = syn.make_smpl_model(seed={"random": 0, "skeleton": 3, "dense": 5}[weights], weights=weights)
    b = _bodies(64)
    smpl = _smpl(model, dev, gemm_mode)
    o32, o64 = _oracle(model), _oracle(model, torch.float64)
    T = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    # axis-angle mode (core/trainer.py:415)
    out = smpl(betas=T(b['betas']), body_pose=T(b['pose_aa'][:, 3:]), global_orient=T(b['pose_aa'][:, :3]),
               pose2rot=True)
    ref = o32(b['betas'], b['pose_aa'][:, 3:], b['pose_aa'][:, :3], pose2rot=True)
    ref64 = o64(b['betas'], b['pose_aa'][:, 3:], b['pose_aa'][:, :3], pose2rot=True)
    assert out.vertices.shape == (64, 6890, 3) and out.joints.shape == (64, 49, 3)
    assert _maxabs(out.vertices, ref['vertices']) <= VERT_TOL
    assert _maxabs(out.joints, ref['joints']) <= VERT_TOL
    assert _maxabs(out.smpl_joints, ref['joints45']) <= VERT_TOL
    assert _maxabs(out.vertices, ref64['vertices']) <= VERT_TOL
    # rotation-matrix mode (models/whmr.py:132-137)
    out = smpl(betas=T(b['betas']), body_pose=T(b['rotmat'][:, 1:]), global_orient=T(b['rotmat'][:, :1]),
               pose2rot=False, return_transforms=True)
    ref = o32(b['betas'], b['rotmat'][:, 1:], b['rotmat'][:, :1], pose2rot=False)
    assert _maxabs(out.vertices, ref['vertices']) <= VERT_TOL
    assert _maxabs(out.joints, ref['joints']) <= VERT_TOL
    assert _maxabs(out.rel_transforms.view(64, 24, 3, 4), ref['A'][:, :, :3, :]) <= VERT_TOL


@pytest.mark.parametrize("gemm_mode", GEMM_MODES)
def test_smpl_matches_reference_smpl_webuser_outputs(dev, gemm_mode):
    """The CUDA path against vertices / posed joints produced by executing the REFERENCE's own in-tree SMPL code
    (models/smpl_webuser/{serialization,posemapper,verts,lbs}.py, see tests/golden/make_golden_smpl.py)."""
    import hashlib
    import whmr_b200.synthetic as syn
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "smpl_webuser_outputs.npz"))
    model = syn.make_smpl_model(seed=int(g['model_seed']), weights="random")
    h = hashlib.sha256()
    for k in ('v_template', 'shapedirs', 'posedirs', 'weights', 'J_regressor', 'parents'):
        h.update(np.ascontiguousarray(model[k]).tobytes())
    assert h.hexdigest() == str(g['model_sha256'])
    smpl = _smpl(model, dev, gemm_mode)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    out = smpl(betas=T(g['betas']), body_pose=T(g['pose'][:, 3:]), global_orient=T(g['pose'][:, :3]), pose2rot=True)
    assert _maxabs(out.vertices, g['verts']) <= VERT_TOL
    assert _maxabs(out.smpl_joints[:, :24], g["Jtr"]) <= VERT_TOL


def test_smpl_simt_error_budget(dev, smpl_model):
    """The exact-fp32 GEMM path should sit ~1e-6 from the fp64 oracle (summation order only)."""
    b = _bodies(32)
    smpl = _smpl(smpl_model, dev, "fp32_simt")
    T = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    out = smpl(betas=T(b['betas']), body_pose=T(b['rotmat'][:, 1:]), global_orient=T(b['rotmat'][:, :1]), pose2rot=False)
    ref64 = _oracle(smpl_model, torch.float64)(b['betas'], b['rotmat'][:, 1:].astype(np.float64),
                                               b['rotmat'][:, :1].astype(np.float64), pose2rot=False)
    assert _maxabs(out.vertices, ref64['vertices']) <= 3e-6


@pytest.mark.parametrize("gemm_mode", GEMM_MODES)
@pytest.mark.parametrize("B", [0, 1, 3, 65, 257])
def test_smpl_ragged_batches(dev, smpl_model, B, gemm_mode):
    """empty, single and tile-straddling batches (the vendored smplx test also drives batch 0:
    models/ViTPose/tests/test_external/test_smpl.py:60-78)."""
    smpl = _smpl(smpl_model, dev, gemm_mode)
    b = _bodies(max(B, 1), seed=7)
    T = lambda a: torch.from_numpy(a[:B]).to(dev)  # noqa: E731
    out = smpl(betas=T(b['betas']), body_pose=T(b['rotmat'][:, 1:]), global_orient=T(b['rotmat'][:, :1]), pose2rot=False)
    assert out.vertices.shape == (B, 6890, 3) and out.joints.shape == (B, 49, 3)
    if B:
        ref = _oracle(smpl_model)(b['betas'][:B], b['rotmat'][:B, 1:], b['rotmat'][:B, :1], pose2rot=False)
        assert _maxabs(out.vertices, ref['vertices']) <= VERT_TOL
        assert _maxabs(out.joints, ref['joints']) <= VERT_TOL


@pytest.mark.parametrize("gemm_mode", GEMM_MODES)
def test_smpl_chunked_large_batch_properties(dev, smpl_model, gemm_mode):
    """Full-size property checks (no oracle at this size): chunk boundaries must be invisible, the
    identity pose returns v_shaped, bodies are independent of their batch position."""
    from whmr_b200 import ops
    B = 2048 + 37
    b = _bodies(B, seed=11)
    smpl = _smpl(smpl_model, dev, gemm_mode)
    T = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    betas, rot = T(b['betas']), T(b['rotmat'])
    big = smpl(betas=betas, body_pose=rot[:, 1:], global_orient=rot[:, :1], pose2rot=False)
    # same bodies in a different order / batch size -> bitwise identical per body
    perm = torch.randperm(B, device=dev)[:300]
    small = smpl(betas=betas[perm], body_pose=rot[perm, 1:], global_orient=rot[perm, :1], pose2rot=False)
    assert torch.equal(big.vertices[perm], small.vertices)
    assert torch.equal(big.joints[perm], small.joints)
    # spot-check 16 bodies straddling chunk boundaries against the oracle
    idx = [0, 1, 255, 256, 511, 512, 767, 768, 769, 1023, 1535, 1536, 2047, 2048, B - 2, B - 1]
    ref = _oracle(smpl_model)(b['betas'][idx], b['rotmat'][idx, 1:], b['rotmat'][idx, :1], pose2rot=False)
    assert _maxabs(big.vertices[idx], ref['vertices']) <= VERT_TOL
    # identity pose => v_shaped
    eye = torch.eye(3, device=dev).expand(B, 24, 3, 3).contiguous()
    ident = smpl(betas=betas, body_pose=eye[:, 1:], global_orient=eye[:, :1], pose2rot=False)
    vs = torch.from_numpy(smpl_model['v_template']).to(dev)[None] + torch.einsum(
        'bl,mkl->bmk', betas, torch.from_numpy(smpl_model['shapedirs']).to(dev))
    # the shape blend rides in the pose-blend contraction: bf16x3 carries 2^-16 relative error on
    # |S.beta| <= ~0.1 m (measured 2.1e-6), the fp32 / 3xtf32 paths stay below 1e-6
    assert float((ident.vertices - vs).abs().max()) <= (4e-6 if gemm_mode == "bf16x3" else 1.5e-6)
    assert ops is not None


@pytest.mark.parametrize("gemm_mode", ["bf16x3", "3xtf32"])
def test_smpl_fused_chunk_boundary_and_readouts(dev, smpl_model, gemm_mode):
    """(both arithmetics of the fused kernel: bf16 hi/lo and the kind::tf32 instantiation = 3xTF32)
    The fused kernel runs in 4096-body chunks (the chunk bounds the read-out partial buffer): bodies on either side of
    the boundary, with the full BodyModelHead read-out table, are bitwise independent of their batch position and match
    the oracle; the flat group-major read-out buffer is laid out for the WHOLE batch, not per chunk."""
    from oracle.smpl_oracle import regressor_readouts
    from whmr_b200.regressor import BodyModelHead
    B = 4096 + 53
    b = _bodies(B, seed=13)
    smpl = _smpl(smpl_model, dev, gemm_mode)
    head = BodyModelHead(smpl, smpl_model['Dmap0'], smpl_model['Dmap1'], smpl_model['ssm'], smpl_model['J_regressor_h36m'])
    assert smpl._state(dev)[0].is_fused()
    T = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    rot, betas, cam = T(b['rotmat']), T(b['betas']), T(b['cam'])
    big = head(rot, betas, cam, J_regressor=True)
    idx = [0, 1, 4094, 4095, 4096, 4097, B - 1]
    sel = torch.tensor(idx, device=dev)
    small = head(rot[sel], betas[sel], cam[sel], J_regressor=True)
    for k in ('verts', 'kp_3d', 'smpl_kp_3d', 'sub_verts', 'temp_verts', 'markers', 'joints49', 'kp_2d'):
        assert torch.equal(big[k][sel], small[k]), k
    ref = _oracle(smpl_model)(b['betas'][idx], b['rotmat'][idx, 1:], b['rotmat'][idx, :1], pose2rot=False)
    rr = regressor_readouts(smpl_model, ref['vertices'])
    assert _maxabs(big['verts'][sel], ref['vertices']) <= VERT_TOL
    assert _maxabs(big['kp_3d'][sel], rr['kp_3d_h36m']) <= VERT_TOL
    assert _maxabs(big['temp_verts'][sel], rr['temp_verts']) <= VERT_TOL


def test_smpl_two_kernel_paths_still_match(dev):
    """WHMR_FUSED=0: pose blend (pose_blend_tc_kernel, bf16x3 and 3xTF32) and skinning (skin_tc_kernel) as two kernels with
    the pose-offset intermediate in HBM -- the round-1 path the fused kernel replaced; kept as a switch, so kept tested.
    The switch is read once per process, hence the subprocess."""
    import subprocess
    import sys
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
import whmr_b200.synthetic as syn
from whmr_b200.smpl import SMPL
from oracle.smpl_oracle import SMPLOracle
dev = torch.device("cuda:0")
model = syn.make_smpl_model(seed=0)
orc = SMPLOracle(model)
for mode in ("bf16x3", "3xtf32"):
    smpl = SMPL(model=model, gemm_mode=mode).to(dev)
    assert not smpl._state(dev)[0].is_fused()
    for B in (1, 37, 300):
        b = syn.make_bodies(B, seed=B)
        out = smpl(betas=torch.from_numpy(b["betas"]).to(dev), body_pose=torch.from_numpy(b["rotmat"][:, 1:]).to(dev),
                   global_orient=torch.from_numpy(b["rotmat"][:, :1]).to(dev), pose2rot=False)
        ref = orc(b["betas"], b["rotmat"][:, 1:], b["rotmat"][:, :1], pose2rot=False)
        ev = float((out.vertices.cpu() - ref["vertices"]).abs().max())
        ej = float((out.joints.cpu() - ref["joints"]).abs().max())
        assert ev <= 1e-5 and ej <= 1e-5, (mode, B, ev, ej)
print("two-kernel ok")
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, WHMR_FUSED="0"), capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0 and "two-kernel ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_smpl_transl_and_default_params(dev, smpl_model):
    from whmr_b200.smpl import SMPL
    smpl = SMPL(model=smpl_model, batch_size=4, create_transl=True, gemm_mode="fp32_simt").to(dev)
    b = _bodies(4)
    T = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    tr = torch.tensor([[0.1, -0.2, 0.3]] * 4, device=dev)
    out = smpl(betas=T(b['betas']), body_pose=T(b['pose_aa'][:, 3:]), global_orient=T(b['pose_aa'][:, :3]), transl=tr)
    ref = _oracle(smpl_model)(b['betas'], b['pose_aa'][:, 3:], b['pose_aa'][:, :3], pose2rot=True, transl=tr.cpu())
    assert _maxabs(out.vertices, ref['vertices']) <= VERT_TOL
    assert _maxabs(out.joints, ref['joints']) <= VERT_TOL
    out0 = smpl()      # all-default parameters: zero pose, zero betas, zero transl
    assert out0.vertices.shape == (4, 6890, 3)
    assert _maxabs(out0.vertices[0], smpl_model['v_template']) <= 2e-6


def test_smpl_host_buffer_entry(dev, smpl_model):
    """whmr_smpl_forward_host: pinned host buffers in/out (the bench's e2e leg)."""
    smpl = _smpl(smpl_model, dev, "fp32_simt")
    h, _ = smpl._state(dev)
    b = _bodies(33)
    betas = torch.from_numpy(b['betas']).pin_memory()
    rot = torch.from_numpy(b['rotmat']).pin_memory()
    verts = torch.empty(33, 6890, 3).pin_memory()
    joints = torch.empty(33, 24, 3).pin_memory()
    h.forward_host(betas, rot, True, verts, joints)
    ref = _oracle(smpl_model)(b['betas'], b['rotmat'][:, 1:], b['rotmat'][:, :1], pose2rot=False)
    assert _maxabs(verts, ref['vertices']) <= VERT_TOL
    assert _maxabs(joints, ref['joints24']) <= VERT_TOL


def test_smpl_forward_into_4_byte_aligned_output(dev, smpl_model):
    """The C ABI takes ANY float-aligned output pointer: a `verts` buffer that starts 4 bytes off an 8-byte boundary must not go
    through the epilogue's float2 stores (smpl_fused_tc.cuh, store64)."""
    from whmr_b200 import _lib
    from whmr_b200._lib import check
    smpl = _smpl(smpl_model, dev, "bf16x3")
    h, _ = smpl._state(dev)
    B = 37
    b = _bodies(B, seed=19)
    betas, rot = torch.from_numpy(b['betas']).to(dev), torch.from_numpy(b['rotmat']).to(dev).contiguous()
    buf = torch.zeros(B * 6890 * 3 + 1, device=dev)
    verts = buf[1:]
    assert verts.data_ptr() % 8 == 4
    joints = torch.empty(B, 24, 3, device=dev)
    ws, n = h.workspace(B)
    with torch.cuda.device(dev):
        check(_lib.lib().whmr_smpl_forward(h._h, betas.data_ptr(), rot.data_ptr(), 1, None, B, verts.data_ptr(),
                                           joints.data_ptr(), None, ws.data_ptr(), n,
                                           torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = _oracle(smpl_model)(b['betas'], b['rotmat'][:, 1:], b['rotmat'][:, :1], pose2rot=False)
    assert _maxabs(verts.view(B, 6890, 3), ref['vertices']) <= VERT_TOL
    assert float(buf[0]) == 0.0


def test_batch_rodrigues(dev):
    from oracle.smpl_oracle import batch_rodrigues
    from whmr_b200.geometry import batch_rodrigues_smplx as gpu_rod
    g = torch.Generator().manual_seed(0)
    aa = torch.randn(500, 3, generator=g) * 1.5
    aa[0] = 0
    aa[1] = torch.tensor([np.pi, 0, 0])
    assert _maxabs(gpu_rod(aa.to(dev)), batch_rodrigues(aa)) <= 1e-6


@pytest.mark.parametrize("gemm_mode", GEMM_MODES)
def test_smpl_real_magnitude_stress(dev, gemm_mode):
    """Blend-shape magnitudes of a real SMPL model: posedirs x10 (sigma 0.02) and shapedirs x3 (sigma 0.03) give pose
    offsets up to ~0.7 m.  The split-operand tensor-core modes keep their margin: <= 1e-5 m against the fp32 oracle and
    against the fp64 oracle."""
    import whmr_b200.synthetic as syn
    model = dict(syn.make_smpl_model(seed=0, weights="random"))
    model['posedirs'] = (model['posedirs'] * 10.0).astype(np.float32)
    model['shapedirs'] = (model['shapedirs'] * 3.0).astype(np.float32)
    b = _bodies(48, seed=21)
    smpl = _smpl(model, dev, gemm_mode)
    T = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    out = smpl(betas=T(b['betas']), body_pose=T(b['rotmat'][:, 1:]), global_orient=T(b['rotmat'][:, :1]), pose2rot=False)
    ref = _oracle(model)(b['betas'], b['rotmat'][:, 1:], b['rotmat'][:, :1], pose2rot=False)
    ref64 = _oracle(model, torch.float64)(b['betas'], b['rotmat'][:, 1:].astype(np.float64),
                                          b['rotmat'][:, :1].astype(np.float64), pose2rot=False)
    off = float((ref64['vertices'] - ref64['vertices'].mean(1, keepdim=True)).abs().max())
    e32, e64 = _maxabs(out.vertices, ref['vertices']), _maxabs(out.vertices, ref64['vertices'])
    print("stress %s: |v| up to %.2f m, max err vs fp32 oracle %.2e m, vs fp64 %.2e m" % (gemm_mode, off, e32, e64))
    assert e32 <= VERT_TOL and e64 <= VERT_TOL
    assert _maxabs(out.joints, ref['joints']) <= VERT_TOL


def test_smpl_65536_bodies_spot_check(dev, smpl_model):
    """BASELINE configs[2]'s largest size in ONE call (16 chunks of 4,096 bodies, 5.4 GB of vertices): 48 bodies spread
    over the call (incl. both sides of chunk boundaries and the last body) against the oracle, and position
    independence (the same bodies as a 48-body batch are bitwise identical)."""
    B = 65536
    free, _ = torch.cuda.mem_get_info()
    if free < 12 * 2**30:
        pytest.skip("needs ~8 GB of free device memory")
    b = _bodies(B, seed=17)
    smpl = _smpl(smpl_model, dev, "bf16x3")
    T = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    betas, rot = T(b['betas']), T(b['rotmat'])
    big = smpl(betas=betas, body_pose=rot[:, 1:], global_orient=rot[:, :1], pose2rot=False)
    rng = np.random.default_rng(0)
    idx = sorted(set([0, 1, 4095, 4096, 8191, 8192, 32767, 32768, 61439, 61440, B - 2, B - 1] +
                     rng.integers(0, B, size=36).tolist()))
    ref = _oracle(smpl_model)(b['betas'][idx], b['rotmat'][idx, 1:], b['rotmat'][idx, :1], pose2rot=False)
    sel = torch.tensor(idx, device=dev)
    assert _maxabs(big.vertices[sel], ref['vertices']) <= VERT_TOL
    assert _maxabs(big.joints[sel], ref['joints']) <= VERT_TOL
    small = smpl(betas=betas[sel], body_pose=rot[sel, 1:], global_orient=rot[sel, :1], pose2rot=False)
    assert torch.equal(big.vertices[sel], small.vertices) and torch.equal(big.joints[sel], small.joints)
    del big, small
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------ read-outs
@pytest.mark.parametrize("gemm_mode", ["fp32_simt", "bf16x3"])
def test_body_model_head_matches_regressor_forward(dev, smpl_model, gemm_mode):
    """All tensors of Regressor.forward's dict that come from the body model (models/whmr.py:189-208)."""
    from oracle import geometry_oracle as G
    from oracle.smpl_oracle import regressor_readouts
    from whmr_b200.regressor import BodyModelHead
    B = 40
    b = _bodies(B, seed=3)
    smpl = _smpl(smpl_model, dev, gemm_mode)
    head = BodyModelHead(smpl, smpl_model['Dmap0'], smpl_model['Dmap1'], smpl_model['ssm'],
                         smpl_model['J_regressor_h36m'])
    T = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    out = head(T(b['rotmat']), T(b['betas']), T(b['cam']), T(b['bbox_height']), T(b['center']), T(b['orig_shape']),
               T(b['Tz']), J_regressor=True)
    ref = _oracle(smpl_model)(b['betas'], b['rotmat'][:, 1:], b['rotmat'][:, :1], pose2rot=False)
    rr = regressor_readouts(smpl_model, ref['vertices'])
    assert _maxabs(out['verts'], ref['vertices']) <= VERT_TOL
    assert _maxabs(out['joints49'], ref['joints']) <= VERT_TOL
    for k_gpu, k_ref in (('sub_verts', 'sub_verts'), ('temp_verts', 'temp_verts'), ('markers', 'markers'),
                         ('smpl_kp_3d', 'smpl_kp_3d'), ('kp_3d', 'kp_3d_h36m')):
        assert out[k_gpu].is_contiguous()
        assert _maxabs(out[k_gpu], rr[k_ref]) <= VERT_TOL, k_gpu
    assert _maxabs(out['pelvis'], rr['smpl_kp_3d'][:, :1]) <= VERT_TOL
    # projections of the GPU joints, checked against the oracle applied to the same joints
    j = out['joints49'].cpu()
    kp = G.projection(j, T(b['cam']).cpu())
    assert _maxabs(out['kp_2d'], kp) * 128.0 <= PX_TOL            # normalised by 256/2 -> pixels
    kpn, focal, cam_t, kp_px = G.full_projection(j, *[torch.from_numpy(b[k]) for k in
                                                      ('cam', 'bbox_height', 'center', 'orig_shape', 'Tz')])
    np.testing.assert_allclose(out['focal_length'].cpu().numpy(), focal.numpy(), rtol=1e-6)
    assert _maxabs(out['pred_cam_t'], cam_t) <= 1e-5
    half = torch.from_numpy(b['orig_shape'][:, ::-1].copy()).unsqueeze(1) / 2
    assert float(((out['kp_2d_w'].cpu() - kpn).abs() * half).max()) <= PX_TOL
    # without H36M regressor 'kp_3d' is the 49 joints (models/whmr.py:140)
    out2 = head(T(b['rotmat']), T(b['betas']), T(b['cam']))
    assert out2['kp_3d'].shape == (B, 49, 3) and 'kp_2d_w' not in out2


def test_readout_generic_csr(dev):
    """long rows, short rows, sub_row, chain-joint sources, group-major output."""
    import scipy.sparse as sp
    from whmr_b200 import ops
    rng = np.random.default_rng(0)
    V, J, B = 500, 24, 9
    dense = rng.random((7, V + J)) * (rng.random((7, V + J)) < 0.2)
    picks = rng.integers(0, V + J, size=11)
    groups = [('dense', sp.csr_matrix(dense)), ('picks', picks), ('empty', sp.csr_matrix((2, V + J)))]
    sub = np.full(7 + 11 + 2, -1, dtype=np.int32)
    sub[3] = 0
    sub[8] = 7
    ro = ops.Readout(groups, V, J, dev, sub_rows=sub)
    verts = torch.randn(B, V, 3, device=dev)
    joints = torch.randn(B, J, 3, device=dev)
    r = ro.apply(verts, joints)
    src = torch.cat([verts, joints], 1).double().cpu()
    d = torch.einsum('rs,bsc->brc', torch.from_numpy(dense), src)
    d[:, 3] -= d[:, 0].clone()
    p = src[:, picks]
    p[:, 1] -= p[:, 0].clone()
    assert _maxabs(r['dense'], d) <= 1e-5
    assert _maxabs(r['picks'], p) <= 1e-6
    assert float(r['empty'].abs().max()) == 0.0


# ------------------------------------------------------------------------------------------ projection
def test_projection_matches_reference_golden(dev, golden):
    from whmr_b200 import geometry as geo
    g = golden
    T = lambda k: torch.from_numpy(g[k]).to(dev)  # noqa: E731
    assert _maxabs(geo.projection(T('proj_points'), T('proj_cam')), g['proj_out']) * 128 <= PX_TOL
    kpn, focal, cam_t, px = geo.full_image_projection(T('proj_points'), T('proj_cam'), T('full_bbox_h'), T('full_center'),
                                                      T('full_orig_shape'), T('full_Tz'), want_px=True)
    assert _maxabs(px, g['full_kp_px']) <= PX_TOL
    assert _maxabs(cam_t, g['full_cam_t']) <= 1e-5
    np.testing.assert_allclose(focal.cpu().numpy(), g['full_focal'], rtol=1e-6)
    assert _maxabs(kpn, g['full_kp_norm']) <= 1e-5
    cc = T('full_orig_shape')[:, [1, 0]] / 2.
    pp = geo.perspective_projection(T('proj_points'), T('pp_rot'), T('full_cam_t'), T('full_focal'), cc, retain_z=True)
    assert _maxabs(pp, g['pp_retain_z']) <= PX_TOL
    # keyword use + broadcast eye, exactly as models/whmr.py:157-163
    pp2 = geo.perspective_projection(T('proj_points'), rotation=torch.eye(3, device=dev).unsqueeze(0).expand(1, -1, -1),
                                     translation=T('full_cam_t'), focal_length=T('full_focal'), camera_center=cc)
    assert _maxabs(pp2, g['full_kp_px']) <= PX_TOL
    ct = geo.convert_pare_to_full_img_cam(T('proj_cam'), T('full_bbox_h'), T('full_center'), T('full_orig_shape')[:, 1],
                                          T('full_orig_shape')[:, 0], focal_length=5000.)
    assert _maxabs(ct, g['full_cam_t_f5000']) <= 1e-5
    with pytest.raises(RuntimeError):
        geo.projection(T('proj_points'), T('proj_cam'), retain_z=True)


def test_rotation_glue_matches_reference_golden(dev, golden):
    """rot6d_to_rotmat / unbiased_gram_schmidt / rotation_matrix_to_angle_axis vs the reference's outputs."""
    from oracle import geometry_oracle as G
    from whmr_b200 import geometry as geo
    g = golden
    T = lambda k: torch.from_numpy(g[k]).to(dev)  # noqa: E731
    assert _maxabs(geo.rot6d_to_rotmat(T('rot6d_in')), g['rot6d_out']) <= 1e-6
    assert _maxabs(geo.unbiased_gram_schmidt(T('ugs_in')), g['ugs_out']) <= 1e-6
    assert _maxabs(geo.rotation_matrix_to_angle_axis(T('r2aa_in')), g['r2aa_out']) <= 2e-5
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(4096, 24, 3, 3, generator=gen) * 0.3 + torch.eye(3)
    R = geo.unbiased_gram_schmidt(x.to(dev))
    assert _maxabs(R, G.unbiased_gram_schmidt(x)) <= 2e-6
    aa = geo.rotation_matrix_to_angle_axis(R.reshape(-1, 3, 3))
    ref = G.rotation_matrix_to_angle_axis(R.reshape(-1, 3, 3).cpu())
    assert _maxabs(aa, ref) <= 5e-5          # atan2 near pi amplifies fp32 rounding of the quaternion
    assert torch.isfinite(aa).all()


def test_projection_large_random(dev):
    from oracle import geometry_oracle as G
    from whmr_b200 import geometry as geo
    b = _bodies(1000, seed=5)
    g = torch.Generator().manual_seed(1)
    pts = torch.randn(1000, 49, 3, generator=g) * 0.4
    kp = geo.projection(pts.to(dev), torch.from_numpy(b['cam']).to(dev))
    assert _maxabs(kp, G.projection(pts, torch.from_numpy(b['cam']))) * 128 <= PX_TOL


def test_geometry_dropins_carry_gradients(dev):
    """The names INTEGRATION.md swaps into models/whmr.py:24-25 / core/trainer.py:27 sit inside the training graph
    (kp_2d_w -> joints / Tz, core/trainer.py:518): perspective_projection, convert_pare_to_full_img_cam and projection
    against autograd through the float64 oracle, in the exact composition of models/whmr.py:147-173."""
    from oracle import geometry_oracle as G
    from whmr_b200 import geometry as geo
    b = _bodies(24, seed=9)
    g = torch.Generator().manual_seed(3)
    joints = (torch.randn(24, 49, 3, generator=g) * 0.4)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731
    cam, Tz, bh, ctr, osh = T(b['cam']), T(b['Tz']), T(b['bbox_height']), T(b['center']), T(b['orig_shape'])
    w1, w2 = torch.randn(24, 49, 2, generator=g), torch.randn(24, 49, 2, generator=g)

    def graph(mod, j, c, tz, dt, dv):
        cv = lambda x: x.to(device=dv, dtype=dt)  # noqa: E731
        s = c[:, 0].detach()
        focal = s * cv(bh) * tz / 2.
        cc = cv(osh)[:, [1, 0]] / 2.
        cam_t = mod.convert_pare_to_full_img_cam(c.detach(), cv(bh), cv(ctr), cv(osh)[:, 1], cv(osh)[:, 0], Tz=tz)
        eye = torch.eye(3, device=dv, dtype=dt).unsqueeze(0).expand(1, -1, -1)
        kpw = mod.perspective_projection(j, rotation=eye, translation=cam_t, focal_length=focal, camera_center=cc)
        kpw = kpw / cc.unsqueeze(1) - 1
        kp = mod.projection(j.detach(), c)
        return (kpw * cv(w1)).sum() + (kp * cv(w2)).sum()

    leaf = lambda x, dt, dv: x.to(device=dv, dtype=dt).clone().requires_grad_(True)  # noqa: E731
    j64, c64, t64 = leaf(joints, torch.float64, 'cpu'), leaf(cam, torch.float64, 'cpu'), leaf(Tz, torch.float64, 'cpu')
    graph(G, j64, c64, t64, torch.float64, 'cpu').backward()
    jg, cg, tg = leaf(joints, torch.float32, dev), leaf(cam, torch.float32, dev), leaf(Tz, torch.float32, dev)
    graph(geo, jg, cg, tg, torch.float32, dev).backward()
    rel = lambda a, r: float((a.detach().cpu().double() - r).abs().max() / (r.abs().max() + 1e-30))  # noqa: E731
    assert jg.grad is not None and cg.grad is not None and tg.grad is not None
    assert rel(jg.grad, j64.grad) <= 1e-4
    assert rel(cg.grad, c64.grad) <= 1e-4
    assert rel(tg.grad, t64.grad) <= 1e-4
    # scalar focal + no rotation + retain_z, gradient to points and translation
    pts = leaf(joints, torch.float32, dev)
    tr = leaf(torch.tensor([[0.1, -0.2, 5.0]]).expand(24, 3).contiguous(), torch.float32, dev)
    out = geo.perspective_projection(pts, None, tr, 1000., torch.zeros(24, 2, device=dev), retain_z=True)
    out[..., :2].sum().backward()
    p64 = leaf(joints, torch.float64, 'cpu')
    t64b = leaf(torch.tensor([[0.1, -0.2, 5.0]]).expand(24, 3).contiguous(), torch.float64, 'cpu')
    x = p64 + t64b.unsqueeze(1)
    (1000. * x[..., :2] / x[..., 2:3]).sum().backward()
    assert rel(pts.grad, p64.grad) <= 1e-4 and rel(tr.grad, t64b.grad) <= 1e-4


def test_geometry_glue_cpu_and_grad_dispatch(dev, golden):
    """init-time CPU use (models/whmr.py:65) and autograd use take the torch path; plain CUDA tensors the kernels"""
    from whmr_b200 import geometry as geo
    x = torch.from_numpy(golden['rot6d_in'])
    assert _maxabs(geo.rot6d_to_rotmat(x), golden['rot6d_out']) <= 1e-6             # CPU tensor: no raise
    assert _maxabs(geo.rot6d_to_rotmat(x.to(dev)), golden['rot6d_out']) <= 1e-6     # kernel
    y = torch.from_numpy(golden['ugs_in']).to(dev).requires_grad_(True)
    geo.unbiased_gram_schmidt(y).sum().backward()
    assert y.grad is not None


def test_handles_are_released_with_their_module(dev, smpl_model):
    """ops._HANDLES / _READOUTS hold weak references: dropping the SMPL module frees the ~90 MB of device constants."""
    import gc
    from whmr_b200 import ops
    smpl = _smpl(smpl_model, dev, "bf16x3")
    b = _bodies(4)
    T = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    smpl(betas=T(b['betas']), body_pose=T(b['rotmat'][:, 1:]), global_orient=T(b['rotmat'][:, :1]), pose2rot=False)
    torch.cuda.synchronize()
    n_h, n_r = len(ops._HANDLES), len(ops._READOUTS)
    free0 = torch.cuda.mem_get_info()[0]
    del smpl
    gc.collect()
    torch.cuda.synchronize()
    assert len(ops._HANDLES) == n_h - 1 and len(ops._READOUTS) == n_r - 1
    assert torch.cuda.mem_get_info()[0] - free0 >= 50 * 2**20      # cudaFree'd constants are back


# ------------------------------------------------------------------------------------------ sampling
@pytest.mark.parametrize("tag", ["a", "b"])
def test_sampling_matches_reference_golden(dev, golden, tag):
    from whmr_b200.maf_extractor import MAF_Extractor
    g = golden
    ext = MAF_Extractor(mesh_downsampling=None).to(dev)
    sd = {k: torch.from_numpy(g['maf_' + k.replace('.', '_')]) for k in
          ('conv0.weight', 'conv0.bias', 'conv1.weight', 'conv1.bias', 'conv2.weight', 'conv2.bias')}
    ext.load_state_dict(sd, strict=False)
    feat = torch.from_numpy(g['samp_%s_feat' % tag]).to(dev)
    pts = torch.from_numpy(g['samp_%s_points' % tag]).to(dev)
    maf, pf = ext.sampling(pts, im_feat=feat)
    ref = g['samp_%s_point_feat' % tag]
    assert _maxabs(pf, ref) <= FEAT_RTOL * np.abs(ref).max()
    assert _maxabs(maf, g['samp_%s_mesh_align' % tag]) <= 1e-3    # PyTorch Conv1d MLP (may use TF32 convs)
    # state-driven forward (self.im_feat / self.cam), models/maf_extractor.py:126-143
    ext.im_feat = torch.from_numpy(g['fwd_feat']).to(dev)
    ext.cam = torch.from_numpy(g['fwd_cam']).to(dev)
    _, pf2 = ext(torch.from_numpy(g['fwd_p']).to(dev), None, None, None, None)
    assert _maxabs(pf2, g['fwd_point_feat']) <= FEAT_RTOL * np.abs(g['fwd_point_feat']).max()


@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
@pytest.mark.parametrize("H,W,N,C", [(32, 24, 63, 256), (64, 48, 67, 256), (128, 96, 67, 256), (14, 14, 431, 256),
                                      (28, 28, 431, 256), (56, 56, 431, 256), (7, 5, 1, 3), (16, 16, 6890, 8)])
def test_sampling_vs_grid_sample(dev, layout, H, W, N, C):
    from oracle.sampling_oracle import grid_sample_points
    from whmr_b200 import ops
    import whmr_b200.synthetic as syn
    B = 3
    g = torch.Generator().manual_seed(H * 1000 + N)
    feat = torch.randn(B, C, H, W, generator=g)
    pts = torch.from_numpy(syn.make_sample_points(B, N, seed=H))
    pts[0, 0] = torch.tensor([-1.0, -1.0])
    pts[1, 0] = torch.tensor([1.0, 1.0])
    pts[2, 0] = torch.tensor([float('inf'), 0.0]) if N > 1 else pts[2, 0]
    ref = grid_sample_points(feat, pts.clone().nan_to_num(posinf=5.0))
    if layout == "nchw":
        out = ops.sample_bilinear(feat.to(dev), pts.to(dev), ops.LAYOUT_NCHW)
    else:
        out = ops.sample_bilinear(feat.permute(0, 2, 3, 1).contiguous().to(dev), pts.to(dev), ops.LAYOUT_NHWC)
    assert out.shape == (B, C, N)
    assert _maxabs(out, ref) <= FEAT_RTOL * float(ref.abs().max())


@pytest.mark.parametrize("layout", ["nchw", "channels_last"])
@pytest.mark.parametrize("H,W", [(14, 14), (28, 28), (56, 56)])
def test_sampling_config3_full_batch(dev, layout, H, W):
    """BASELINE configs[3] at its full size (B = 1024, N = 431, C = 256; the 14x14 / 28x28 levels take the shared-memory
    staged NCHW kernel): 24 bodies spread over the batch against grid_sample on the CPU."""
    from oracle.sampling_oracle import grid_sample_points
    from whmr_b200 import ops
    import whmr_b200.synthetic as syn
    B, N, C = 1024, 431, 256
    g = torch.Generator(device=dev).manual_seed(H)
    feat = torch.randn(B, C, H, W, generator=g, device=dev)
    pts = torch.from_numpy(syn.make_sample_points(B, N, seed=3)).to(dev)
    f_in = feat.contiguous(memory_format=torch.channels_last) if layout == "channels_last" else feat
    out = ops.sample_bilinear(f_in, pts, ops.LAYOUT_NCHW)
    assert out.shape == (B, C, N) and out.is_contiguous()
    idx = sorted(set([0, 1, 511, 512, B - 1] + np.random.default_rng(H).integers(0, B, size=19).tolist()))
    ref = grid_sample_points(feat[idx].cpu(), pts[idx].cpu())
    assert _maxabs(out[idx], ref) <= FEAT_RTOL * float(ref.abs().max())


@pytest.mark.parametrize("H,W", [(1, 1), (1, 9), (9, 1), (2, 2)])
def test_sampling_degenerate_maps(dev, H, W):
    """align_corners=True with a size-1 axis maps every coordinate to pixel 0 (grid_sample semantics); empty point sets
    and empty batches return empty tensors."""
    from oracle.sampling_oracle import grid_sample_points
    from whmr_b200 import ops
    B, C, N = 2, 5, 33
    g = torch.Generator().manual_seed(H * 10 + W)
    feat = torch.randn(B, C, H, W, generator=g)
    pts = torch.rand(B, N, 2, generator=g) * 2.4 - 1.2
    ref = grid_sample_points(feat, pts)
    for layout, f in ((ops.LAYOUT_NCHW, feat), (ops.LAYOUT_NHWC, feat.permute(0, 2, 3, 1).contiguous())):
        out = ops.sample_bilinear(f.to(dev), pts.to(dev), layout)
        assert _maxabs(out, ref) <= FEAT_RTOL * max(float(ref.abs().max()), 1e-6)
    assert ops.sample_bilinear(feat.to(dev), pts[:, :0].to(dev), ops.LAYOUT_NCHW).shape == (B, C, 0)
    assert ops.sample_bilinear(feat[:0].to(dev), pts[:0].to(dev), ops.LAYOUT_NCHW).shape == (0, C, N)


def test_project_sample_channels_last_matches_nchw(dev):
    """MAF_Extractor.forward (weak projection + sampling in one launch) on channels_last maps goes through the NHWC kernel
    with the projection fused in: same 2-D points, same features as the NCHW kernel and as the oracle."""
    from oracle import geometry_oracle as G
    from oracle.sampling_oracle import grid_sample_points
    from whmr_b200 import constants, ops
    import whmr_b200.synthetic as syn
    B, C, H, W, N = 4, 70, 24, 20, 67
    g = torch.Generator().manual_seed(77)
    feat = torch.randn(B, C, H, W, generator=g)
    p3 = torch.randn(B, N, 3, generator=g) * 0.35
    cam = torch.from_numpy(syn.make_bodies(B, seed=9)['cam'])
    a_feat, a_pts = ops.project_sample_op(feat.to(dev), p3.to(dev), cam.to(dev), constants.FOCAL_LENGTH, 256., 256.,
                                          ops.LAYOUT_NCHW)
    fcl = feat.to(dev).contiguous(memory_format=torch.channels_last)
    b_feat, b_pts = ops.project_sample_op(fcl, p3.to(dev), cam.to(dev), constants.FOCAL_LENGTH, 256., 256., ops.LAYOUT_NCHW)
    assert torch.equal(a_pts, b_pts)
    ref_pts = G.projection(p3, cam)
    assert _maxabs(b_pts, ref_pts) * 128 <= PX_TOL
    ref = grid_sample_points(feat, b_pts.cpu())
    assert _maxabs(b_feat, ref) <= FEAT_RTOL * float(ref.abs().max())
    assert _maxabs(a_feat, ref) <= FEAT_RTOL * float(ref.abs().max())


def test_sampling_from_pinned_host_maps_in_place(dev, smpl_model):
    """A feature map in page-locked HOST memory is read in place by the sampling kernels (unified addressing; only the taps'
    sectors cross PCIe): same bits as the device-resident map, for the plain sampler, the projection + sampling launch, the
    drop-in MAF_Extractor and a whole loop pass (eager and as a CUDA graph) with the finest level host-resident.  A pageable
    CPU tensor is still refused: there is no CPU path."""
    from whmr_b200 import _lib, constants, ops
    from whmr_b200.loop import RegressorLoop, make_loop_inputs
    from whmr_b200.maf_extractor import MAF_Extractor
    import whmr_b200.synthetic as syn
    g = torch.Generator().manual_seed(31)
    for (B, C, H, W, N) in ((3, 70, 24, 20, 67), (2, 64, 14, 14, 431), (2, 33, 128, 96, 67)):
        feat = torch.randn(B, C, H, W, generator=g)
        hfeat = feat.pin_memory()
        assert ops.is_host_map(hfeat) and not ops.is_host_map(feat)
        pts = (torch.rand(B, N, 2, generator=g) * 2.2 - 1.1).to(dev)
        assert torch.equal(ops.sample_bilinear(hfeat, pts), ops.sample_bilinear(feat.to(dev), pts))
        nhwc = feat.permute(0, 2, 3, 1).contiguous()
        assert torch.equal(ops.sample_bilinear(nhwc.pin_memory(), pts, ops.LAYOUT_NHWC),
                           ops.sample_bilinear(nhwc.to(dev), pts, ops.LAYOUT_NHWC))
        hcl = nhwc.pin_memory().permute(0, 3, 1, 2)      # [B,C,H,W] view in channels_last strides of pinned memory
        assert ops.is_host_map(hcl) and not hcl.is_contiguous()
        assert torch.equal(ops.sample_bilinear(hcl, pts), ops.sample_bilinear(nhwc.to(dev), pts, ops.LAYOUT_NHWC))
        p3 = (torch.randn(B, N, 3, generator=g) * 0.35).to(dev)
        cam = torch.from_numpy(syn.make_bodies(B, seed=9)['cam']).to(dev)
        a = ops.project_sample(hfeat, p3, cam, constants.FOCAL_LENGTH, 256., 256.)
        b = ops.project_sample(feat.to(dev), p3, cam, constants.FOCAL_LENGTH, 256., 256.)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and a[0].device == p3.device
        with pytest.raises(_lib.WhmrError):
            ops.sample_bilinear(feat, pts)
        with pytest.raises(NotImplementedError):     # no silent loss of a gradient: a host map is inference-only
            ops.sample_bilinear(feat.pin_memory().requires_grad_(True), pts)
    # drop-in extractor (sampling op + the module's PyTorch MLP) on a host-resident map
    ext = MAF_Extractor(mesh_downsampling=None).to(dev).eval()
    feat = torch.randn(2, 256, 32, 24, generator=g)
    p3 = (torch.randn(2, 67, 3, generator=g) * 0.35).to(dev)
    cam = torch.from_numpy(syn.make_bodies(2, seed=3)['cam']).to(dev)
    ext.fused = False
    with torch.no_grad():
        ext.im_feat, ext.cam = feat.pin_memory(), cam
        y_h, pf_h = ext(p3, None, None, None, None)
        ext.im_feat = feat.to(dev)
        y_d, pf_d = ext(p3, None, None, None, None)
    assert torch.equal(pf_h, pf_d) and torch.equal(y_h, y_d)
    # loop pass with the finest level left on the host
    B = 6
    loop = RegressorLoop(smpl_model, dev)
    feats, params, bbox = make_loop_inputs(B, dev, seed=8)
    ref = loop.step(feats, params, bbox)
    ref = {k: [t.clone() for t in v] if isinstance(v, list) else v.clone() for k, v in ref.items() if k in ('verts', 'point_feats', 'kp_2d_w')}
    hf = [feats[0], feats[1], feats[2].cpu().pin_memory()]
    got = loop.step(hf, params, bbox)
    assert torch.equal(got['verts'], ref['verts']) and torch.equal(got['kp_2d_w'], ref['kp_2d_w'])
    assert all(torch.equal(a, b) for a, b in zip(got['point_feats'], ref['point_feats']))
    gr, outs = loop.capture(hf, params, bbox)
    for t in outs['point_feats']:
        t.zero_()
    gr.replay()
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(outs['point_feats'], ref['point_feats']))


def test_maf_project_matches_reference_golden(dev, golden):
    from whmr_b200.maf_extractor import MAF_Extractor
    g = golden
    T = lambda k: torch.from_numpy(g[k]).to(dev)  # noqa: E731
    ext = MAF_Extractor(mesh_downsampling=None).to(dev)
    full, crop = ext.project(T('fwd_p'), T('fwd_cam'), T('mproj_center'), T('mproj_scale'), T('mproj_focal'),
                             T('mproj_img_center'), return_full=True)
    assert _maxabs(full, g['mproj_full']) <= PX_TOL
    assert _maxabs(crop, g['mproj_crop']) * 128 <= PX_TOL
    tr = ext.get_trans(T('fwd_cam'), T('mproj_center'), T('mproj_scale'), T('mproj_focal'), T('mproj_img_center'))
    d = ext.perspective_projection(T('fwd_p') + tr, None, None, T('mproj_focal'), T('mproj_img_center'),
                                   distortion=T('mproj_kc'))
    assert _maxabs(d, g['mproj_distorted']) <= PX_TOL


# ------------------------------------------------------------------------------------------ metrics
def test_joint_errors_match_reference_golden(dev, golden):
    from oracle import metrics_oracle as M
    from whmr_b200 import ops
    g = golden
    mp, pa = ops.joint_errors(torch.from_numpy(g['pa_S1']).to(dev), torch.from_numpy(g['pa_S2']).to(dev))
    assert _maxabs(pa, g['pa_err']) <= 1e-6
    assert _maxabs(mp, M.mpjpe(g['pa_S1'], g['pa_S2'])) <= 1e-6
    rng = np.random.default_rng(3)
    S1 = rng.normal(0, 0.3, size=(2000, 14, 3)).astype(np.float32)
    S2 = (S1 + rng.normal(0, 0.05, size=S1.shape)).astype(np.float32)
    S2[:5] = S1[:5] * 0.5 + 0.3          # vendored KAT: exact similarity => zero error
    S1[5] = 0; S1[5, :, 0] = np.linspace(0, 1, 14); S2[5] = S1[5] * 2     # rank-1 (collinear) input
    mp, pa = ops.joint_errors(torch.from_numpy(S1).to(dev), torch.from_numpy(S2).to(dev))
    assert _maxabs(pa, M.pa_mpjpe(S1, S2)) <= 2e-6
    assert float(pa[:5].max()) <= 1e-6


def test_eval_pass_matches_oracle(dev, smpl_model):
    """BASELINE config 5 (evaluate/eval.py:157-223): GT SMPL (axis-angle) + predicted SMPL (rotmat) + H36M 17 -> 14
    pelvis-centred + MPJPE / PA-MPJPE / PVE, against the CPU oracle; then size-independent properties on a
    3DPW-sized shard walked in chunks."""
    from oracle import metrics_oracle as M
    from whmr_b200.evaluate import EvalPass
    import whmr_b200.synthetic as syn
    smpl = _smpl(smpl_model, dev, None)
    ev = EvalPass(smpl, smpl_model['J_regressor_h36m'])
    n = 24
    gt, pr = syn.make_bodies(n, seed=21), syn.make_bodies(n, seed=22)
    pr_rot = pr['rotmat'].copy(); pr_betas = pr['betas'].copy()
    pr_rot[:4] = gt['rotmat'][:4]; pr_betas[:4] = gt['betas'][:4]        # exact predictions -> (near) zero errors
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    got = ev(T(gt['pose_aa']), T(gt['betas']), T(pr_rot), T(pr_betas))
    ref = M.eval_pass(smpl_model, gt['pose_aa'], gt['betas'], pr_rot, pr_betas)
    for k in ('mpjpe', 'pa_mpjpe', 'pve'):
        assert _maxabs(got[k], ref[k]) <= 2e-5, k          # joints/vertices carry <= 1e-5 m each side
    assert float(got['mpjpe'][:4].max()) <= 2e-5 and float(got['pve'][:4].max()) <= 2e-5
    # predicted vertices handed in directly (the model's global_verts, eval.py:181)
    pv, _ = ev.joints(T(pr_betas), T(pr_rot), True)
    got2 = ev(T(gt['pose_aa']), T(gt['betas']), pred_vertices=pv)
    for k in ('mpjpe', 'pa_mpjpe', 'pve'):
        assert _maxabs(got2[k], got[k].cpu()) <= 2e-6, k
    # chunked pass over a larger shard: chunk size must be invisible, frames independent of position
    N = 3000
    g2, p2 = syn.make_bodies(N, seed=23), syn.make_bodies(N, seed=24)
    a = ev.run_sharded(g2['pose_aa'], g2['betas'], p2['rotmat'], p2['betas'], chunk=1024)
    b = ev.run_sharded(g2['pose_aa'], g2['betas'], p2['rotmat'], p2['betas'], chunk=777)
    for k in a:
        assert a[k].shape == (N,) and torch.equal(a[k], b[k]), k
        assert bool(torch.isfinite(a[k]).all()) and float(a[k].min()) >= 0
    assert bool((a['pa_mpjpe'] <= a['mpjpe'] + 1e-6).all())            # alignment never increases the error


# ------------------------------------------------------------------------------------------ backward
def test_projection_backward_matches_autograd_of_oracle(dev):
    """d(weak projection)/d(points, cam) and d(weak + predicted-focal block)/d(points, cam, Tz) against torch autograd
    through the float64 oracle (utils/geometry.py:289-307, models/whmr.py:147-173 restated)."""
    from oracle import geometry_oracle as G
    from whmr_b200 import constants, ops
    import whmr_b200.synthetic as syn
    B, N = 7, 49
    b = syn.make_bodies(B, seed=41)
    rng = np.random.default_rng(5)
    pts = rng.normal(0, 0.4, size=(B, N, 3)).astype(np.float32)
    T = lambda a, g=False: torch.from_numpy(np.ascontiguousarray(a)).to(dev).requires_grad_(g)  # noqa: E731
    # (the geometry oracle builds float32 constants, as the reference does: autograd reference in float32)
    D = lambda a, g=False: torch.from_numpy(np.ascontiguousarray(a)).float().requires_grad_(g)  # noqa: E731
    gk = rng.normal(size=(B, N, 2)).astype(np.float32)
    gw = rng.normal(size=(B, N, 2)).astype(np.float32)
    gf = rng.normal(size=(B,)).astype(np.float32)
    gt = rng.normal(size=(B, 3)).astype(np.float32)
    # weak projection
    p, c = T(pts, True), T(b['cam'], True)
    out = ops.project_weak_op(p, c, constants.FOCAL_LENGTH, 256., 256.)
    out.backward(T(gk))
    p64, c64 = D(pts, True), D(b['cam'], True)
    G.projection(p64, c64).backward(D(gk))
    rel = lambda a, r: float((a.detach().cpu().double() - r).abs().max() / (r.abs().max() + 1e-30))  # noqa: E731
    assert rel(p.grad, p64.grad) <= 1e-4 and rel(c.grad, c64.grad) <= 1e-4
    # weak + full block, all four outputs feeding the loss
    p, c, tz = T(pts, True), T(b['cam'], True), T(b['Tz'], True)
    kp, kpw, fl, cam_t = ops.project_weak_full_op(p, c, T(b['bbox_height']), T(b['center']), T(b['orig_shape']), tz,
                                                  constants.FOCAL_LENGTH, 256., 256.)
    ((kp * T(gk)).sum() + (kpw * T(gw)).sum() + (fl * T(gf)).sum() + (cam_t * T(gt)).sum()).backward()
    p64, c64, tz64 = D(pts, True), D(b['cam'], True), D(b['Tz'], True)
    kn, f64, ct64, _ = G.full_projection(p64, c64, D(b['bbox_height']), D(b['center']), D(b['orig_shape']), tz64)
    ((G.projection(p64, c64) * D(gk)).sum() + (kn * D(gw)).sum() + (f64 * D(gf)).sum() + (ct64 * D(gt)).sum()).backward()
    assert rel(p.grad, p64.grad) <= 1e-4
    assert rel(c.grad, c64.grad) <= 1e-4
    assert rel(tz.grad, tz64.grad) <= 1e-4


@pytest.mark.parametrize("layout", ["nchw", "channels_last", "shared_grid"])
def test_sampling_backward_matches_grid_sample_autograd(dev, layout):
    """d(sample_bilinear)/d(feature map) against autograd through F.grid_sample (the reference's call,
    models/maf_extractor.py:119); points are detached in the reference (models/whmr.py:586-591)."""
    from oracle.sampling_oracle import grid_sample_points
    from whmr_b200 import constants, ops
    import whmr_b200.synthetic as syn
    B, C, H, W, N = 3, 32, 16, 12, 67
    g = torch.Generator().manual_seed(9)
    feat = torch.randn(B, C, H, W, generator=g)
    pts = torch.from_numpy(syn.make_sample_points(B, N, seed=7))
    if layout == "shared_grid":
        pts = pts[:1].expand(B, -1, -1).contiguous()
    go = torch.randn(B, C, N, generator=g)
    f64 = feat.double().requires_grad_(True)
    grid_sample_points(f64, pts.double()).backward(go.double())
    fd = feat.to(dev)
    if layout == "channels_last":
        fd = fd.contiguous(memory_format=torch.channels_last)
    fd.requires_grad_(True)
    pd = pts[0].to(dev) if layout == "shared_grid" else pts.to(dev)
    out = ops.sample_bilinear_op(fd, pd, ops.LAYOUT_NCHW)
    out.backward(go.to(dev))
    assert fd.grad.shape == feat.shape
    assert _maxabs(fd.grad, f64.grad) <= 1e-5 * float(f64.grad.abs().max())
    # MAF_Extractor.forward path: projection fused into the sampling launch
    b = syn.make_bodies(B, seed=43)
    p3 = torch.from_numpy(np.random.default_rng(2).normal(0, 0.35, size=(B, N, 3)).astype(np.float32))
    fd2 = feat.to(dev).requires_grad_(True)
    pf, p2d = ops.project_sample_op(fd2, p3.to(dev), torch.from_numpy(b['cam']).to(dev), constants.FOCAL_LENGTH, 256., 256.,
                                    ops.LAYOUT_NCHW)
    pf.backward(go.to(dev))
    f64b = feat.double().requires_grad_(True)
    grid_sample_points(f64b, p2d.detach().cpu().double()).backward(go.double())
    assert _maxabs(fd2.grad, f64b.grad) <= 1e-5 * float(f64b.grad.abs().max())


class _default_f64:
    """the geometry oracle creates its constants (eye, camera centre) in the default dtype, as the reference does"""

    def __enter__(self):
        self.old = torch.get_default_dtype()
        torch.set_default_dtype(torch.float64)

    def __exit__(self, *a):
        torch.set_default_dtype(self.old)


def _relmax(a, r):
    r = r.detach().double().cpu()
    return float((a.detach().double().cpu() - r).abs().max() / (r.abs().max() + 1e-30))


@pytest.mark.parametrize("gemm_mode", GEMM_MODES)
@pytest.mark.parametrize("weights", ["random", "dense"])
def test_smpl_backward_matches_autograd_of_oracle(dev, weights, gemm_mode):
    """d(vertices, 49 joints, 45 smpl joints)/d(betas, rotation matrices) of SMPL.forward in the model's mode
    (pose2rot=False, models/whmr.py:132-137) against torch autograd through the float64 oracle (smplx lbs restated)."""
    import whmr_b200.synthetic as syn
    model = syn.make_smpl_model(seed={"random": 0, "dense": 5}[weights], weights=weights)
    B = 11
    b = _bodies(B, seed=51)
    rng = np.random.default_rng(8)
    gv = rng.normal(size=(B, 6890, 3)).astype(np.float32)
    gj = rng.normal(size=(B, 49, 3)).astype(np.float32) * 30
    gs = rng.normal(size=(B, 45, 3)).astype(np.float32) * 30
    smpl = _smpl(model, dev, gemm_mode)
    T = lambda a, g=False: torch.from_numpy(np.ascontiguousarray(a)).to(dev).requires_grad_(g)  # noqa: E731
    D = lambda a, g=False: torch.from_numpy(np.ascontiguousarray(a)).double().requires_grad_(g)  # noqa: E731
    o64 = _oracle(model, torch.float64)

    def ref_grads(use_v, use_j):
        be, rm = D(b['betas'], True), D(b['rotmat'], True)
        ref = o64(be, rm[:, 1:], rm[:, :1], pose2rot=False)
        loss = 0
        if use_v:
            loss = loss + (ref['vertices'] * D(gv)).sum()
        if use_j:
            loss = loss + (ref['joints'] * D(gj)).sum() + (ref['joints45'] * D(gs)).sum()
        loss.backward()
        return be.grad, rm.grad

    for use_v, use_j in ((True, True), (True, False), (False, True)):
        be, rm = T(b['betas'], True), T(b['rotmat'], True)
        out = smpl(betas=be, body_pose=rm[:, 1:], global_orient=rm[:, :1], pose2rot=False)
        loss = 0
        if use_v:
            loss = loss + (out.vertices * T(gv)).sum()
        if use_j:
            loss = loss + (out.joints * T(gj)).sum() + (out.smpl_joints * T(gs)).sum()
        loss.backward()
        gb_ref, gr_ref = ref_grads(use_v, use_j)
        assert be.grad.shape == (B, 10) and rm.grad.shape == (B, 24, 3, 3)
        assert _relmax(be.grad, gb_ref) <= 1e-4, (use_v, use_j)
        assert _relmax(rm.grad, gr_ref) <= 1e-4, (use_v, use_j)


def test_body_model_head_backward(dev, smpl_model):
    """Training losses of core/trainer.py:380-636 sit on verts, kp_3d (H36M 17->14, pelvis-centred), smpl_kp_3d,
    the down-sampled meshes and the projected keypoints: gradients through every read-out of BodyModelHead and both
    projections down to (pred_rotmat, pred_shape, pred_cam, Tz) against autograd through the float64 oracle."""
    from oracle import geometry_oracle as G
    from oracle.smpl_oracle import regressor_readouts
    from whmr_b200.regressor import BodyModelHead
    B = 9
    b = _bodies(B, seed=53)
    rng = np.random.default_rng(9)
    smpl = _smpl(smpl_model, dev, "bf16x3")
    head = BodyModelHead(smpl, smpl_model['Dmap0'], smpl_model['Dmap1'], smpl_model['ssm'], smpl_model['J_regressor_h36m'])
    T = lambda a, g=False: torch.from_numpy(np.ascontiguousarray(a)).to(dev).requires_grad_(g)  # noqa: E731
    D = lambda a, g=False: torch.from_numpy(np.ascontiguousarray(a)).double().requires_grad_(g)  # noqa: E731
    keys = {'verts': (6890, 3), 'kp_3d': (14, 3), 'smpl_kp_3d': (45, 3), 'sub_verts': (1723, 3), 'temp_verts': (431, 3),
            'markers': (67, 3), 'kp_2d': (49, 2), 'kp_2d_w': (49, 2)}
    gs = {k: rng.normal(size=(B,) + s).astype(np.float32) * (1.0 if s[0] > 1000 else 20.0) for k, s in keys.items()}
    for stage in (None, 1, 2):
        head.train_stage = stage
        rm, be, cam, tz = T(b['rotmat'], True), T(b['betas'], True), T(b['cam'], True), T(b['Tz'], True)
        out = head(rm, be, cam, T(b['bbox_height']), T(b['center']), T(b['orig_shape']), tz, J_regressor=True, is_train=True)
        sum((out[k] * T(g)).sum() for k, g in gs.items()).backward()
        rm64, be64, cam64, tz64 = D(b['rotmat'], True), D(b['betas'], True), D(b['cam'], True), D(b['Tz'], True)
        ref = _oracle(smpl_model, torch.float64)(be64, rm64[:, 1:], rm64[:, :1], pose2rot=False)
        rr = regressor_readouts(smpl_model, ref['vertices'], torch.float64)
        j = ref['joints']
        # models/whmr.py:142-165: which projection sees the joints depends on cfg.TRAIN.STAGE; pred_cam is detached
        # inside the predicted-focal block.  stage None = the reference's configured default (TRAIN.STAGE: 2)
        eff = 2 if stage is None else stage
        jw = j if eff == 1 else j.detach()
        jf = j if eff == 2 else j.detach()
        camf = cam64.detach()
        with _default_f64():
            kp = G.projection(jw, cam64)
            kpn = G.full_projection(jf, camf, D(b['bbox_height']), D(b['center']), D(b['orig_shape']), tz64)[0]
        refs = {'verts': ref['vertices'], 'kp_3d': rr['kp_3d_h36m'], 'smpl_kp_3d': rr['smpl_kp_3d'],
                'sub_verts': rr['sub_verts'], 'temp_verts': rr['temp_verts'], 'markers': rr['markers'], 'kp_2d': kp,
                'kp_2d_w': kpn}
        sum((refs[k] * D(g)).sum() for k, g in gs.items()).backward()
        assert _relmax(rm.grad, rm64.grad) <= 2e-4, stage
        assert _relmax(be.grad, be64.grad) <= 2e-4, stage
        assert _relmax(cam.grad, cam64.grad) <= 2e-4, stage
        assert _relmax(tz.grad, tz64.grad) <= 2e-4, stage


def test_estimate_translation_matches_reference_golden(dev, golden):
    """utils/geometry.py:386-408 (host NumPy + np.linalg.solve per sample in the reference) on the device."""
    from whmr_b200 import geometry as geo
    from oracle import geometry_oracle as G
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    S, j2 = golden['et_S'], golden['et_joints_2d']
    out = geo.estimate_translation(T(S), T(j2), focal_length=5000., img_size=[224., 224.])
    np.testing.assert_allclose(out.cpu().numpy(), golden['et_out'], rtol=2e-6, atol=2e-6)
    out = geo.estimate_translation(T(S), T(j2), focal_length=1000., img_size=[256., 192.])
    np.testing.assert_allclose(out.cpu().numpy(), golden['et_out_f1000'], rtol=2e-6, atol=2e-6)
    # a training-sized batch against the oracle
    rng = np.random.default_rng(3)
    B = 300
    S = rng.normal(0, 0.35, size=(B, 49, 3)).astype(np.float32)
    t = np.stack([rng.uniform(-.4, .4, B), rng.uniform(-.4, .4, B), rng.uniform(2.5, 12, B)], 1).astype(np.float32)
    pe = S + t[:, None]
    kp = (5000. * pe[..., :2] / pe[..., 2:] + 112.).astype(np.float32)
    j2 = np.concatenate([kp, rng.uniform(0.1, 1, size=(B, 49, 1)).astype(np.float32)], -1)
    out = geo.estimate_translation(T(S), T(j2)).cpu().numpy()
    np.testing.assert_allclose(out, G.estimate_translation(S, j2), rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(out, t, rtol=2e-3, atol=2e-3)      # noise-free key points: recovers the translation
    assert geo.estimate_translation(T(S[:0]), T(j2[:0])).shape == (0, 3)


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors instead of routing anywhere else."""
    from whmr_b200 import ops
    from whmr_b200._lib import WhmrError
    with pytest.raises(WhmrError):
        ops.project_weak(torch.zeros(1, 2, 3), torch.ones(1, 3), 1000., 256., 256.)
    assert os.path.exists(os.path.join(os.path.dirname(ops.__file__), "libwhmr_b200.so"))


def test_regressor_loop_matches_oracle_and_graph_replay(dev, smpl_model):
    """The whole hot-path loop (BASELINE configs[1] shape, small batch): eager, side-stream overlap and CUDA-graph
    replay all match the CPU oracle loop."""
    from oracle.loop_oracle import LoopOracle, to_cpu_inputs
    from whmr_b200.loop import RegressorLoop, make_loop_inputs
    B = 12
    loop = RegressorLoop(smpl_model, dev)
    feats, params, bbox = make_loop_inputs(B, dev, seed=4)
    ref = LoopOracle(smpl_model).step(*to_cpu_inputs(feats, params, bbox))

    def check(got):
        assert _maxabs(got['verts'], ref['verts']) <= VERT_TOL
        assert _maxabs(got['global_verts'], ref['global_verts']) <= VERT_TOL
        assert _maxabs(got['kp_3d'], ref['kp_3d']) <= VERT_TOL
        assert _maxabs(got['global_kp_3d'], ref['global_kp_3d']) <= VERT_TOL
        assert _maxabs(got['markers'], ref['markers']) <= VERT_TOL
        assert _maxabs(got['kp_2d'], ref['kp_2d']) * 128 <= PX_TOL
        half = (bbox['orig_shape'].cpu()[:, [1, 0]] / 2).unsqueeze(1)
        assert float(((got['kp_2d_w'].cpu() - ref['kp_2d_w']).abs() * half).max()) <= PX_TOL
        for a, b in zip(got['point_feats'], ref['point_feats']):
            assert _maxabs(a, b) <= FEAT_RTOL * float(b.abs().max())
        # rotation glue folded into the chain kernel (models/whmr.py:129-130, 174, 190, 632-633)
        assert _maxabs(got['rotmat'], ref['rotmat']) <= 2e-6
        assert _maxabs(got['pose'], ref['pose']) <= 2e-5 and got['pose'].shape == (B, 72)
        assert _maxabs(got['theta'], ref['theta']) <= 2e-5 and got['theta'].shape == (B, 85)
        assert _maxabs(got['global_pose'], ref['global_pose']) <= 2e-5

    check(loop.step(feats, params, bbox))
    loop.overlap = True
    check(loop.step(feats, params, bbox))
    g, outs = loop.capture(feats, params, bbox)
    for k in ('verts', 'global_verts', 'kp_3d', 'markers', 'kp_2d', 'kp_2d_w', 'pose', 'theta', 'rotmat', 'global_pose'):   # (not the aliased inputs)
        outs[k].zero_()
    g.replay()
    torch.cuda.synchronize()
    check(outs)
    loop.overlap = False
    g2, outs2 = loop.capture(feats, params, bbox)
    g2.replay()
    torch.cuda.synchronize()
    check(outs2)


# ------------------------------------------------------------------ sampling + reduce_dim in one kernel
def _golden_extractor(g, dev):
    from whmr_b200.maf_extractor import MAF_Extractor
    ext = MAF_Extractor(mesh_downsampling=None).to(dev).eval()
    sd = {k: torch.from_numpy(g['maf_' + k.replace('.', '_')]) for k in
          ('conv0.weight', 'conv0.bias', 'conv1.weight', 'conv1.bias', 'conv2.weight', 'conv2.bias')}
    ext.load_state_dict(sd, strict=False)
    return ext


def _launches():
    from whmr_b200 import _lib
    return _lib.launch_count()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_fused_sampling_mlp_matches_reference_golden(dev, golden, tag):
    """MAF_Extractor.sampling / .forward under no_grad = one maf_fused_kernel launch; mesh_align_feat against the output
    of the reference's own module (tests/golden/make_golden.py: models/maf_extractor.py:75-143) at <= 1e-4 relative."""
    g = golden
    ext = _golden_extractor(g, dev)
    feat = torch.from_numpy(g['samp_%s_feat' % tag]).to(dev)
    pts = torch.from_numpy(g['samp_%s_points' % tag]).to(dev)
    with torch.no_grad():
        ext.sampling(pts, im_feat=feat)           # first call also splits the weights (one more launch)
        n0 = _launches()
        maf, pf = ext.sampling(pts, im_feat=feat)
        assert _launches() - n0 == 1
    ref_m, ref_p = g['samp_%s_mesh_align' % tag], g['samp_%s_point_feat' % tag]
    assert maf.shape == ref_m.shape and pf.shape == ref_p.shape
    assert _maxabs(pf, ref_p) <= FEAT_RTOL * np.abs(ref_p).max()
    assert _maxabs(maf, ref_m) <= FEAT_RTOL * np.abs(ref_m).max()
    # channels_last maps take the NHWC instantiation: same numbers
    with torch.no_grad():
        maf_cl, pf_cl = ext.sampling(pts, im_feat=feat.contiguous(memory_format=torch.channels_last))
    assert _maxabs(maf_cl, ref_m) <= FEAT_RTOL * np.abs(ref_m).max()
    assert torch.equal(pf_cl, pf)
    # without the [B,256,N] output
    ext.return_point_feat = False
    with torch.no_grad():
        maf2, none = ext.sampling(pts, im_feat=feat)
    assert none is None and torch.equal(maf2, maf)
    ext.return_point_feat = True
    # state-driven forward (projection fused in), models/maf_extractor.py:126-143
    ext.im_feat = torch.from_numpy(g['fwd_feat']).to(dev)
    ext.cam = torch.from_numpy(g['fwd_cam']).to(dev)
    with torch.no_grad():
        n0 = _launches()
        maf3, pf3 = ext(torch.from_numpy(g['fwd_p']).to(dev), None, None, None, None)
        assert _launches() - n0 == 1
    assert _maxabs(pf3, g['fwd_point_feat']) <= FEAT_RTOL * np.abs(g['fwd_point_feat']).max()
    assert _maxabs(maf3, g['fwd_mesh_align']) <= FEAT_RTOL * np.abs(g['fwd_mesh_align']).max()
    # a parameter update is picked up (version counter), and the autograd path still gives the PyTorch MLP
    with torch.no_grad():
        ext.conv2.bias.add_(0.25)
        maf4, _ = ext.sampling(pts, im_feat=feat)
    maf5, _ = ext.sampling(pts, im_feat=feat)      # grad enabled + parameters require grad -> unfused path
    assert maf5.requires_grad
    assert _maxabs(maf4, maf5.detach().cpu()) <= 2e-3 * float(maf5.detach().abs().max())   # cuDNN runs the convs in TF32
    assert _maxabs(maf4, ref_m) > 0.1


@pytest.mark.parametrize("layout", ["nchw", "channels_last"])
@pytest.mark.parametrize("B,N,H,W,mode", [(5, 67, 64, 48, "points"), (3, 63, 32, 24, "grid"), (4, 67, 128, 96, "project"),
                                           (96, 431, 14, 14, "points"), (2, 431, 56, 56, "project"), (1, 1, 7, 5, "points")])
def test_fused_sampling_mlp_vs_oracle(dev, layout, B, N, H, W, mode):
    """maf_fused_kernel against grid_sample + the fp64 reduce_dim oracle: tile tails (B*N not a multiple of 128), tiles
    spanning bodies, more tiles than CTAs (96 x 431 points = 324 tiles on 296 CTAs), the shared iteration-0 grid, the
    projection-fused form, points outside the map, both memory layouts."""
    from oracle import geometry_oracle as G
    from oracle.sampling_oracle import grid_sample_points, reduce_dim
    from whmr_b200 import constants
    from whmr_b200.maf_extractor import MAF_Extractor
    import whmr_b200.synthetic as syn
    gen = torch.Generator().manual_seed(B * 1000 + N + H)
    ext = MAF_Extractor(mesh_downsampling=None)
    with torch.no_grad():
        for p in ext.parameters():
            p.copy_(torch.randn(p.shape, generator=gen) * (0.5 if p.dim() == 1 else 2.0 / np.sqrt(p.shape[1])))
    convs = [(c.weight.detach().double(), c.bias.detach().double()) for c in ext.filters]
    ext = ext.to(dev).eval()
    feat = torch.randn(B, 256, H, W, generator=gen)
    f_in = feat.to(dev)
    if layout == "channels_last":
        f_in = f_in.contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        if mode == "project":
            b = syn.make_bodies(B, seed=7)
            cam = torch.from_numpy(b['cam'])
            p3 = torch.randn(B, N, 3, generator=gen) * torch.tensor([0.45, 0.45, 0.2])
            pts = G.projection(p3, cam)
            ext.im_feat, ext.cam = f_in, cam.to(dev)
            maf, pf = ext(p3.to(dev), None, None, None, None)
        elif mode == "grid":
            pts1 = torch.from_numpy(syn.grid_points('vitpose'))[:N]
            pts = pts1[None].expand(B, -1, -1)
            maf, pf = ext.sampling(pts1.to(dev), im_feat=f_in)
        else:
            pts = torch.from_numpy(syn.make_sample_points(B, N, seed=H))
            maf, pf = ext.sampling(pts.to(dev), im_feat=f_in)
    ref_p = grid_sample_points(feat.double(), pts.double())
    ref_m = reduce_dim(ref_p, convs)
    assert pf.shape == (B, 256, N) and maf.shape == (B, 32 * N)
    assert _maxabs(pf, ref_p) <= FEAT_RTOL * float(ref_p.abs().max())
    assert _maxabs(maf, ref_m) <= FEAT_RTOL * float(ref_m.abs().max())
    assert float(ref_m.abs().max()) > 0.5 and float((ref_m > 0).double().mean()) > 0.1     # the check is not vacuous


def test_batch_rodrigues_quaternion_variant_matches_reference_golden(dev, golden):
    """utils/geometry.py:14-51 (the variant `utils.geometry.batch_rodrigues` names; core/trainer.py:244) against the output
    of the reference's own function, plus identity at theta = 0 (vendored KAT tests/test_losses/test_mesh_losses.py:24-25)."""
    from whmr_b200 import geometry
    R = geometry.batch_rodrigues(torch.from_numpy(golden['rodq_in']).to(dev))
    assert R.shape == golden['rodq_out'].shape
    assert _maxabs(R, golden['rodq_out']) <= 1e-6
    I = geometry.batch_rodrigues(torch.zeros(5, 3, device=dev))
    assert _maxabs(I, torch.eye(3).expand(5, 3, 3)) <= 1e-6


def test_loop_projections_folded_into_the_finishing_launch(dev, smpl_model):
    """Deferred schedule: the four joint projections (models/whmr.py:142-173, 237) ride in the read-out finishing launch
    (whmr_readout_finish_project_multi) -- 14 launches per pass instead of 18 -- and give the results of the stand-alone
    projection kernels (same device functions) and of the oracle."""
    from oracle.loop_oracle import LoopOracle, to_cpu_inputs
    from whmr_b200.loop import RegressorLoop, make_loop_inputs
    B = 9
    loop = RegressorLoop(smpl_model, dev)
    feats, params, bbox = make_loop_inputs(B, dev, seed=6)
    loop.step(feats, params, bbox)
    n0 = _launches()
    a = loop.step(feats, params, bbox)
    n_fused = _launches() - n0
    loop.head.fuse_projection = False
    n0 = _launches()
    b = loop.step(feats, params, bbox)
    n_separate = _launches() - n0
    assert (n_fused, n_separate) == (14, 18)
    for k in ('kp_2d', 'kp_2d_w', 'focal_length', 'pred_cam_t'):
        assert a[k].shape == b[k].shape
        assert _maxabs(a[k], b[k].cpu()) <= 1e-6 * max(1.0, float(b[k].abs().max())), k
    ref = LoopOracle(smpl_model).step(*to_cpu_inputs(feats, params, bbox))
    half = (bbox['orig_shape'].cpu()[:, [1, 0]] / 2).unsqueeze(1)
    assert float(((a['kp_2d_w'].cpu() - ref['kp_2d_w']).abs() * half).max()) <= PX_TOL
    assert _maxabs(a['kp_2d'], ref['kp_2d']) * 128 <= PX_TOL
    assert _maxabs(a['focal_length'], ref['focal_length']) <= 1e-6 * float(ref['focal_length'].abs().max())
    assert _maxabs(a['pred_cam_t'], ref['pred_cam_t']) <= 1e-5


def test_fused_sampling_mlp_random_shapes(dev):
    """Randomised sweep of the fused sampling + reduce_dim kernel and the samplers behind it: tile boundaries (B*N below,
    at and just above multiples of 128), one-pixel maps and map edges, more tiles than CTAs with a ragged tail, both memory
    layouts, the [B,256,N] output on and off, against grid_sample + the fp64 MLP oracle."""
    from oracle.sampling_oracle import grid_sample_points, reduce_dim
    from whmr_b200.maf_extractor import MAF_Extractor
    rng = np.random.default_rng(11)
    gen = torch.Generator().manual_seed(11)
    ext = MAF_Extractor(mesh_downsampling=None)
    with torch.no_grad():
        for p in ext.parameters():
            p.copy_(torch.randn(p.shape, generator=gen) * (0.5 if p.dim() == 1 else 2.0 / np.sqrt(p.shape[1])))
    convs = [(c.weight.detach().double(), c.bias.detach().double()) for c in ext.filters]
    ext = ext.to(dev).eval()
    cases = [(1, 128, 3, 3), (2, 64, 1, 9), (1, 129, 9, 1), (3, 43, 1, 1), (7, 55, 6, 4), (150, 130, 5, 5), (1, 127, 2, 2)]
    for _ in range(9):
        cases.append((int(rng.integers(1, 9)), int(rng.integers(1, 200)), int(rng.integers(1, 20)), int(rng.integers(1, 20))))
    for ci, (B, N, H, W) in enumerate(cases):
        feat = torch.randn(B, 256, H, W, generator=gen)
        pts = (torch.rand(B, N, 2, generator=gen) * 2.6 - 1.3)
        pts[0, 0] = torch.tensor([1.0, 1.0])
        ref_p = grid_sample_points(feat.double(), pts.double())
        ref_m = reduce_dim(ref_p, convs)
        for lay in ("nchw", "channels_last"):
            f_in = feat.to(dev)
            if lay == "channels_last":
                f_in = f_in.contiguous(memory_format=torch.channels_last)
            ext.return_point_feat = bool(ci % 2)
            with torch.no_grad():
                maf, pf = ext.sampling(pts.to(dev), im_feat=f_in)
            assert maf.shape == (B, 32 * N), (B, N, H, W, lay)
            assert _maxabs(maf, ref_m) <= FEAT_RTOL * max(float(ref_m.abs().max()), 1e-3), (B, N, H, W, lay)
            if pf is not None:
                assert _maxabs(pf, ref_p) <= FEAT_RTOL * max(float(ref_p.abs().max()), 1e-3), (B, N, H, W, lay)
            else:
                assert not ext.return_point_feat
    ext.return_point_feat = True


@pytest.mark.parametrize("maxm", ["3", "4"])
def test_smpl_fused_two_issuers_option(dev, maxm):
    """WHMR_FUSED_ISSUERS=2 (two pose-blend issuing threads with per-issuer full barriers on the shared operand rings, the
    skinning issuer loading its own A^T tiles; profiles/r02_notes.md section 4): same vertices, joints and read-outs as the
    oracle at ragged batch sizes incl. a multi-chunk one.  The switch is read once per process, hence the subprocess."""
    import subprocess
    import sys
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
import whmr_b200.synthetic as syn
from whmr_b200 import ops
from whmr_b200.loop import RegressorLoop
from oracle.smpl_oracle import SMPLOracle, regressor_readouts
dev = torch.device("cuda:0")
model = syn.make_smpl_model(seed=0)
loop = RegressorLoop(model, dev)
h, _ = loop.smpl._state(dev)
ro = loop.head._readout(dev, True)
orc = SMPLOracle(model)
for B in (1, 37, 300, 4096 + 130):
    b = syn.make_bodies(B, seed=B)
    betas, rot = torch.from_numpy(b["betas"]).to(dev), torch.from_numpy(b["rotmat"]).to(dev)
    v, j, flat = ops.smpl_lbs_readout(h.id, ro.id, betas, rot, True)
    torch.cuda.synchronize()
    sel = sorted(set(list(range(min(B, 6))) + list(range(max(0, B - 6), B)) + ([4090, 4095, 4096, 4097] if B > 4100 else [])))
    ref = orc(b["betas"][sel], b["rotmat"][sel][:, 1:], b["rotmat"][sel][:, :1], pose2rot=False)
    ev = float((v[sel].cpu() - ref["vertices"]).abs().max())
    rr = regressor_readouts(model, ref["vertices"])
    got = ro.split(flat, B)
    em = float((got["markers"][sel].cpu() - rr["markers"]).abs().max())
    ek = float((got["kp_3d_h36m"][sel].cpu() - rr["kp_3d_h36m"]).abs().max())
    assert ev <= 1e-5 and em <= 1e-5 and ek <= 1e-5, (B, ev, em, ek)
print("two-issuer ok")
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),)
    env = dict(os.environ, WHMR_FUSED_ISSUERS="2", WHMR_FUSED_MAXM=maxm)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "two-issuer ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
