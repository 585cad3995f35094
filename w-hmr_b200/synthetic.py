"""Seeded synthetic SMPL-shaped model and inputs.

The licensed SMPL weights (data/smpl/SMPL_NEUTRAL.pkl, models/whmr.py:72) and the
auxiliary regressors are not shipped with the reference, so every parity test and
benchmark in this repo runs on a synthetic model with the real *schema* -- the one the
reference's vendored tests emit (models/ViTPose/tests/utils/mesh_utils.py:19-27) --
but with non-trivial values.  NumPy only; no torch import here.
"""
import os
import pickle

import numpy as np

from .constants import (NUM_BETAS, NUM_JOINTS, NUM_POSE_BASIS, NUM_VERTS, SMPL_PARENTS,
                        vertex_joint_selector_ids)

N_SUB1 = 1723   # Dmap0 rows, models/whmr.py:92-96
N_SUB2 = 431    # Dmap1 rows
N_MARKERS = 67  # data/smpl/smpl_ssm.npy, models/whmr.py:100,336


def _sparse_rows(rng, n_rows, n_cols, nnz):
    """Dense [n_rows,n_cols] with `nnz` random non-negative entries per row, rows sum to 1."""
    m = np.zeros((n_rows, n_cols), dtype=np.float64)
    for r in range(n_rows):
        cols = rng.choice(n_cols, size=nnz, replace=False)
        w = rng.uniform(0.05, 1.0, size=nnz)
        m[r, cols] = w / w.sum()
    return m.astype(np.float32)


def _skeleton_rest_joints():
    """A human-ish rest skeleton (metres) inside the v_template box; only used to give the
    'skeleton' weight mode spatial locality."""
    j = np.zeros((NUM_JOINTS, 3))
    j[0] = (0, 0.0, 0)          # pelvis
    j[1] = (0.09, -0.08, 0); j[2] = (-0.09, -0.08, 0); j[3] = (0, 0.12, 0)
    j[4] = (0.10, -0.45, 0); j[5] = (-0.10, -0.45, 0); j[6] = (0, 0.25, 0)
    j[7] = (0.10, -0.85, 0); j[8] = (-0.10, -0.85, 0); j[9] = (0, 0.32, 0)
    j[10] = (0.11, -0.89, 0.10); j[11] = (-0.11, -0.89, 0.10); j[12] = (0, 0.52, 0)
    j[13] = (0.07, 0.43, 0); j[14] = (-0.07, 0.43, 0); j[15] = (0, 0.62, 0.03)
    j[16] = (0.17, 0.45, 0); j[17] = (-0.17, 0.45, 0)
    j[18] = (0.25, 0.2, 0) ; j[19] = (-0.25, 0.2, 0)
    j[20] = (0.28, -0.05, 0); j[21] = (-0.28, -0.05, 0)
    j[22] = (0.29, -0.13, 0); j[23] = (-0.29, -0.13, 0)
    return j


def make_smpl_model(seed=0, weights="random", n_verts=NUM_VERTS, regressor_nnz=32):
    """Synthetic SMPL-shaped model.

    weights="random"   : SURVEY 8d literal spec -- 4 distinct random joints per vertex.
    weights="skeleton" : 4 nearest joints of a rest skeleton (spatially coherent, like
                         the real model); used to test locality-sensitive code paths.
    weights="dense"    : every joint has a non-zero weight (stresses the general path).
    Returns a dict of float32/int arrays in smplx's in-memory layout:
      v_template [V,3], shapedirs [V,3,10], posedirs [207, V*3] (smplx: reshape(-1,207).T),
      J_regressor [24,V], weights [V,24], parents [24] int64, kintree_table [2,24] uint32,
      f [13776,3] int32, J_regressor_extra [9,V], J_regressor_h36m [17,V],
      vertex_ids [21] int64, ssm [67] int64, Dmap0 [1723,V], Dmap1 [431,1723] (dense, one-hot rows).
    """
    rng = np.random.default_rng(seed)
    V = n_verts
    m = {}
    box = np.array([0.3, 0.9, 0.15])
    if weights == "skeleton":
        # sample vertices around the skeleton's bones so nearest-joint structure is meaningful
        J0 = _skeleton_rest_joints()
        par = np.array(SMPL_PARENTS)
        bone = rng.integers(1, NUM_JOINTS, size=V)
        t = rng.uniform(0, 1, size=(V, 1))
        p = J0[par[bone]] * (1 - t) + J0[bone] * t
        m['v_template'] = (p + rng.normal(0, 0.04, size=(V, 3))).astype(np.float32)
    else:
        m['v_template'] = (rng.uniform(-1, 1, size=(V, 3)) * box).astype(np.float32)
    m['shapedirs'] = rng.normal(0, 0.01, size=(V, 3, NUM_BETAS)).astype(np.float32)
    m['posedirs'] = rng.normal(0, 0.002, size=(NUM_POSE_BASIS, V * 3)).astype(np.float32)

    W = np.zeros((V, NUM_JOINTS), dtype=np.float64)
    if weights == "random":
        for v in range(V):
            js = rng.choice(NUM_JOINTS, size=4, replace=False)
            w = rng.uniform(0, 1, size=4) + 1e-3
            W[v, js] = w / w.sum()
    elif weights == "skeleton":
        J0 = _skeleton_rest_joints()
        d = np.linalg.norm(m['v_template'][:, None, :].astype(np.float64) - J0[None], axis=-1)
        order = np.argsort(d, axis=1)[:, :4]
        for v in range(V):
            w = np.exp(-(d[v, order[v]] / 0.08) ** 2) + 1e-4
            W[v, order[v]] = w / w.sum()
    elif weights == "dense":
        W = rng.uniform(0, 1, size=(V, NUM_JOINTS)) + 1e-3
        W /= W.sum(axis=1, keepdims=True)
    else:
        raise ValueError(weights)
    m['weights'] = W.astype(np.float32)

    m['J_regressor'] = _sparse_rows(rng, NUM_JOINTS, V, regressor_nnz)
    m['J_regressor_extra'] = _sparse_rows(rng, 9, V, regressor_nnz)
    m['J_regressor_h36m'] = _sparse_rows(rng, 17, V, regressor_nnz)
    parents = np.array(SMPL_PARENTS, dtype=np.int64)
    m['parents'] = parents
    kt = np.zeros((2, NUM_JOINTS), dtype=np.uint32)
    kt[0] = np.where(parents < 0, np.uint32(4294967295), parents).astype(np.uint32)
    kt[1] = np.arange(NUM_JOINTS, dtype=np.uint32)
    m['kintree_table'] = kt
    m['f'] = rng.integers(0, V, size=(13776, 3)).astype(np.int32)
    if V == NUM_VERTS:
        m['vertex_ids'] = np.array(vertex_joint_selector_ids(), dtype=np.int64)
    else:
        m['vertex_ids'] = np.sort(rng.choice(V, size=21, replace=False)).astype(np.int64)
    m['ssm'] = np.sort(rng.choice(V, size=min(N_MARKERS, V), replace=False)).astype(np.int64)
    n1 = min(N_SUB1, V)
    n2 = min(N_SUB2, n1)
    keep0 = np.sort(rng.choice(V, size=n1, replace=False))
    keep1 = np.sort(rng.choice(n1, size=n2, replace=False))
    D0 = np.zeros((n1, V), dtype=np.float32); D0[np.arange(n1), keep0] = 1.0
    D1 = np.zeros((n2, n1), dtype=np.float32); D1[np.arange(n2), keep1] = 1.0
    m['Dmap0'] = D0
    m['Dmap1'] = D1
    return m


def write_smpl_pkl(model, model_dir, gender='neutral'):
    """Emit SMPL_<GENDER>.pkl with the schema of models/ViTPose/tests/utils/mesh_utils.py:19-27
    (and the real SMPL pickles): posedirs [V,3,207], J_regressor scipy csc [24,V]."""
    from scipy.sparse import csc_matrix
    os.makedirs(model_dir, exist_ok=True)
    V = model['v_template'].shape[0]
    d = {
        'f': model['f'],
        'J_regressor': csc_matrix(model['J_regressor'].astype(np.float64)),
        'kintree_table': model['kintree_table'],
        'J': (model['J_regressor'].astype(np.float64) @ model['v_template'].astype(np.float64)),
        'weights': model['weights'].astype(np.float64),
        'posedirs': model['posedirs'].T.reshape(V, 3, NUM_POSE_BASIS).astype(np.float64),
        'v_template': model['v_template'].astype(np.float64),
        'shapedirs': model['shapedirs'].astype(np.float64),
        'bs_type': 'lrotmin', 'bs_style': 'lbs',
    }
    path = os.path.join(model_dir, 'SMPL_%s.pkl' % gender.upper())
    with open(path, 'wb') as fh:
        pickle.dump(d, fh)
    return path


def load_smpl_pkl(path):
    """Read an SMPL .pkl (real or synthetic) into the in-memory layout of make_smpl_model.
    Follows models/smpl_webuser/serialization.py:78-108 and smplx's loader."""
    with open(path, 'rb') as fh:
        d = pickle.load(fh, encoding='latin1')
    V = np.asarray(d['v_template']).shape[0]
    Jr = d['J_regressor']
    Jr = np.asarray(Jr.todense()) if hasattr(Jr, 'todense') else np.asarray(Jr)
    kt = np.asarray(d['kintree_table']).astype(np.int64)
    parents = kt[0].copy()
    parents[0] = -1
    m = {
        'v_template': np.asarray(d['v_template'], dtype=np.float32),
        'shapedirs': np.asarray(d['shapedirs'], dtype=np.float32)[:, :, :NUM_BETAS],
        'posedirs': np.asarray(d['posedirs'], dtype=np.float32).reshape(V * 3, -1).T.copy(),
        'J_regressor': Jr.astype(np.float32),
        'weights': np.asarray(d['weights'], dtype=np.float32),
        'parents': parents,
        'kintree_table': np.asarray(d['kintree_table']),
        'f': np.asarray(d['f']).astype(np.int32),
    }
    return m


# --------------------------------------------------------------------------------------
# inputs
# --------------------------------------------------------------------------------------

def rodrigues_np(aa, dtype=np.float64):
    """smplx==0.1.28 batch_rodrigues restated in NumPy: angle = ||theta + 1e-8||,
    R = I + sin*K + (1-cos)*K@K.  aa [...,3] -> [...,3,3]."""
    aa = np.asarray(aa, dtype=dtype)
    shp = aa.shape[:-1]
    a = aa.reshape(-1, 3)
    angle = np.linalg.norm(a + dtype(1e-8), axis=1, keepdims=True)
    d = a / angle
    c = np.cos(angle)[:, :, None]
    s = np.sin(angle)[:, :, None]
    K = np.zeros((a.shape[0], 3, 3), dtype=dtype)
    K[:, 0, 1] = -d[:, 2]; K[:, 0, 2] = d[:, 1]
    K[:, 1, 0] = d[:, 2]; K[:, 1, 2] = -d[:, 0]
    K[:, 2, 0] = -d[:, 1]; K[:, 2, 1] = d[:, 0]
    R = np.eye(3, dtype=dtype)[None] + s * K + (1 - c) * (K @ K)
    return R.reshape(shp + (3, 3))


def rot6d_to_rotmat_np(x):
    """utils/geometry.py:243-257 in NumPy fp64. x [...,6] viewed as (3,2) -> [...,3,3]."""
    x = np.asarray(x, dtype=np.float64).reshape(-1, 3, 2)
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = a1 / np.maximum(np.linalg.norm(a1, axis=1, keepdims=True), 1e-12)
    a2p = a2 - (b1 * a2).sum(1, keepdims=True) * b1
    b2 = a2p / np.maximum(np.linalg.norm(a2p, axis=1, keepdims=True), 1e-12)
    b3 = np.cross(b1, b2)
    return np.stack([b1, b2, b3], axis=-1)


_REAL_ROWS = None


def real_pose_shape_rows():
    """The real pose/shape rows of the reference's vendored fixtures:
    4 rows of tests/data/mosh/test_mosh.npz + 4 rows of tests/data/h36m/test_h36m.npz +
    the mean parameters (rot6d) + theta=0.  Committed as tests/golden/real_pose_shape.npz by
    tests/golden/make_golden.py.  Returns (pose_aa [10,72] f32, betas [10,10] f32) or None."""
    global _REAL_ROWS
    if _REAL_ROWS is None:
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                         'tests', 'golden', 'real_pose_shape.npz')
        if os.path.exists(p):
            d = np.load(p)
            _REAL_ROWS = (d['pose_aa'].astype(np.float32), d['betas'].astype(np.float32))
        else:
            _REAL_ROWS = False
    return _REAL_ROWS or None


def make_bodies(B, seed=1, rank=0, with_real_rows=True):
    """Synthetic per-body inputs: betas, axis-angle pose, the matching rotmats
    and the camera / bbox quantities projection needs.  Reproducible per (seed, rank)."""
    rng = np.random.default_rng([seed, rank])
    betas = np.clip(rng.normal(0, 1, size=(B, NUM_BETAS)), -3, 3).astype(np.float32)
    pose = rng.normal(0, 0.3, size=(B, NUM_JOINTS, 3))
    pose[:, 0] = rng.normal(0, 1.0, size=(B, 3))
    pose = pose.reshape(B, 72).astype(np.float32)
    rows = real_pose_shape_rows() if with_real_rows else None
    if rows is not None:
        n = min(B, rows[0].shape[0])
        pose[:n] = rows[0][:n]
        betas[:n] = rows[1][:n]
    rotmat = rodrigues_np(pose.reshape(B, NUM_JOINTS, 3).astype(np.float64)).astype(np.float32)
    cam = np.stack([rng.uniform(0.6, 1.2, B), rng.uniform(-0.2, 0.2, B),
                    rng.uniform(-0.2, 0.2, B)], axis=1).astype(np.float32)
    Tz = rng.uniform(1, 10, B).astype(np.float32)
    bbox_height = rng.uniform(100, 600, B).astype(np.float32)
    shapes = np.array([[1080, 1920], [720, 1280]], dtype=np.float32)  # (h, w)
    orig_shape = shapes[rng.integers(0, 2, B)]
    center = (rng.uniform(0.1, 0.9, size=(B, 2)) * orig_shape[:, ::-1]).astype(np.float32)
    scale = (bbox_height / 200.0).astype(np.float32)
    return {
        'betas': betas, 'pose_aa': pose, 'rotmat': rotmat, 'cam': cam, 'Tz': Tz,
        'bbox_height': bbox_height, 'orig_shape': orig_shape, 'center': center, 'scale': scale,
    }


def make_sample_points(B, N, seed=2, rank=0, frac_outside=0.05):
    """Sampling grid points in [-1,1] with a fraction pushed outside (zero-padding branch)."""
    rng = np.random.default_rng([seed, rank, N])
    pts = rng.uniform(-1, 1, size=(B, N, 2))
    out = rng.uniform(0, 1, size=(B, N)) < frac_outside
    pts[out] *= rng.uniform(1.0, 1.6, size=(int(out.sum()), 2))
    return pts.astype(np.float32)


def grid_points(backbone='vitpose'):
    """models/whmr.py:338-347 -- the iteration-0 sampling grid ([1,2,N] buffer, transposed to
    [B,N,2] at :596).  vitpose: 7 wide x 9 high = 63 points; res50: 8x8 = 64.  Returns [N,2]."""
    w, h = (7, 9) if backbone == 'vitpose' else (8, 8)
    xs = np.linspace(-1, 1, w)
    ys = np.linspace(-1, 1, h)
    gx, gy = np.meshgrid(xs, ys, indexing='ij')  # torch.meshgrid default == 'ij'
    return np.stack([gx.reshape(-1), gy.reshape(-1)], axis=1).astype(np.float32)
