"""Tensor-level entry points over the C ABI + their `torch.library` custom-op registrations.

Everything here takes/returns CUDA float32 torch tensors, launches on torch's current stream and
allocates outputs / scratch through torch's caching allocator (so the C side never allocates).
No function in this file has a CPU or PyTorch-eager fallback.
"""
import ctypes as C
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import check

GEMM_FP32_SIMT, GEMM_TC_BF16X3, GEMM_TC_3XTF32 = 0, 1, 2
GEMM_MODES = {"fp32_simt": 0, "bf16x3": 1, "3xtf32": 2}
LAYOUT_NCHW, LAYOUT_NHWC = 0, 1


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _req(t, name, dtype=torch.float32, align=4):
    """contiguous CUDA tensor of `dtype`; copies only when the caller's view is not already so."""
    if t is None:
        return None
    if not torch.is_tensor(t):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise _lib.WhmrError("%s is on %s: whmr_b200 ops run on CUDA only (no CPU fallback)" % (name, t.device))
    if t.dtype != dtype:
        t = t.to(dtype)
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % align:
        t = t.clone()
    return t


def _p(t):
    return None if t is None else t.data_ptr()


# ----------------------------------------------------------------------------------------------
# SMPL handle
# ----------------------------------------------------------------------------------------------
# id -> handle for the torch.library ops (which take plain ints).  WEAK: the owning SMPL / BodyModelHead module holds the
# only strong reference, so dropping the module (or reloading its state dict, which rebuilds the handle) runs __del__ ->
# whmr_smpl_destroy and frees the ~90 MB of re-arranged model constants on the device.
_HANDLES = weakref.WeakValueDictionary()


class SmplHandle:
    """Device-resident, pre-arranged SMPL model (whmr_smpl_create)."""

    def __init__(self, v_template, shapedirs, posedirs, J_regressor, lbs_weights, parents,
                 device, gemm_mode=GEMM_TC_BF16X3):
        f = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float32))  # noqa: E731
        self._arrs = [f(v_template), f(shapedirs), f(posedirs), f(J_regressor), f(lbs_weights),
                      np.ascontiguousarray(np.asarray(parents, dtype=np.int64))]
        vt, sd, pd, jr, w, par = self._arrs
        V, J = vt.shape[0], jr.shape[0]
        NB = sd.shape[-1]
        assert sd.shape == (V, 3, NB) and pd.shape == ((J - 1) * 9, V * 3) and w.shape == (V, J)
        desc = _lib.SmplModelDesc(V, J, NB, vt.ctypes.data, sd.ctypes.data, pd.ctypes.data,
                                  jr.ctypes.data, w.ctypes.data, par.ctypes.data)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.WhmrError("SMPL handle needs a CUDA device, got %s (no CPU fallback)" % (self.device,))
        if not torch.cuda.is_available():
            raise _lib.WhmrError("no CUDA device is available: whmr_b200 has no CPU fallback")
        self.V, self.J, self.NB = V, J, NB
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(_lib.lib().whmr_smpl_create(C.byref(desc), int(gemm_mode), C.byref(self._h)))
        self.gemm_mode = int(gemm_mode)
        self.id = id(self)
        _HANDLES[self.id] = self

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().whmr_smpl_destroy(self._h)
            self._h = C.c_void_p()
        try:
            _HANDLES.pop(getattr(self, "id", None), None)
        except Exception:   # interpreter shutdown
            pass

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_gemm_mode(self, mode):
        check(_lib.lib().whmr_smpl_set_gemm_mode(self._h, int(mode)))
        self.gemm_mode = int(mode)

    def set_probe_events(self, after_chain=None, after_pose_blend=None, after_skin=None):
        """torch.cuda.Event(enable_timing=True, external=True) triple recorded inside the next forward calls."""
        evs = (after_chain, after_pose_blend, after_skin)
        for e in evs:
            if e is not None:
                e.record()           # materialise the lazily created cudaEvent_t
        check(_lib.lib().whmr_smpl_set_probe_events(self._h, *[None if e is None else e.cuda_event for e in evs]))

    def is_fused(self):
        """True when pose blend + skinning run as one kernel (smpl_fused_tc)."""
        return bool(_lib.lib().whmr_smpl_is_fused(self._h))

    def info(self):
        v = [C.c_int32() for _ in range(5)]
        check(_lib.lib().whmr_smpl_get_info(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("n_verts", "n_joints", "n_betas", "ell_width", "gemm_mode"), [x.value for x in v]))

    def workspace(self, B):
        n = int(_lib.lib().whmr_smpl_workspace_bytes(self._h, int(B)))
        return torch.empty(n, dtype=torch.uint8, device=self.device), n

    def forward(self, betas, pose, pose_is_rotmat, transl=None, want_transforms=False, readout=None,
                defer_finish=False, glue=None, root_pose=None):
        """-> (verts [B,V,3], chain joints [B,J,3], A [B,J,12] or None[, flat read-out buffer, read-out scratch])
        With defer_finish the finishing pass of the read-outs may be left to the caller (`readout_finish`, any
        stream); it was deferred iff the returned scratch tensor is non-empty.
        glue: dict(gram_schmidt=bool, cam=[B,3] or None) -> Regressor.forward's rotation glue runs inside the chain kernel
        and three more tensors are appended to the result: rotmat [B,J,3,3] (the rotations used), pose [B,3J], theta.
        root_pose: the root rotation [B,9|3] (`global_orient`) as its own tensor; `pose` then holds the other J-1 joints
        (`body_pose`) -- SMPL.forward's two arguments without the concatenation."""
        betas = _req(betas, "betas", align=16)
        pose = _req(pose, "pose", align=16)
        root_pose = _req(root_pose, "root_pose")
        transl = _req(transl, "transl")
        B = betas.shape[0]
        exp = (self.J - (0 if root_pose is None else 1)) * (9 if pose_is_rotmat else 3)
        if root_pose is not None and root_pose.numel() != B * (9 if pose_is_rotmat else 3):
            raise ValueError("smpl forward: root_pose %s does not match B=%d" % (tuple(root_pose.shape), B))
        if pose.numel() != B * exp or betas.numel() != B * self.NB:
            raise ValueError("smpl forward: betas %s / pose %s do not match B=%d, J=%d, n_betas=%d, rotmat=%s"
                             % (tuple(betas.shape), tuple(pose.shape), B, self.J, self.NB, pose_is_rotmat))
        verts = torch.empty(B, self.V, 3, dtype=torch.float32, device=self.device)
        joints = torch.empty(B, self.J, 3, dtype=torch.float32, device=self.device)
        A = torch.empty(B, self.J, 12, dtype=torch.float32, device=self.device) if want_transforms else None
        ws, n = self.workspace(B)
        g, extra = None, ()
        if glue is None and root_pose is not None:
            g = _lib.SmplGlue(0, None, None, None, None, _p(root_pose))
        if glue is not None:
            cam = _req(glue.get("cam"), "cam")
            rotmat = torch.empty(B, self.J, 3, 3, dtype=torch.float32, device=self.device)
            pose_aa = torch.empty(B, self.J * 3, dtype=torch.float32, device=self.device)
            theta = torch.empty(B, 3 + self.NB + self.J * 3, dtype=torch.float32, device=self.device)
            g = _lib.SmplGlue(int(bool(glue.get("gram_schmidt"))), _p(rotmat), _p(pose_aa), _p(theta), _p(cam), _p(root_pose))
            extra = (rotmat, pose_aa, theta)
        if readout is None and g is None:
            with torch.cuda.device(self.device):
                check(_lib.lib().whmr_smpl_forward(self._h, _p(betas), _p(pose), int(bool(pose_is_rotmat)), _p(transl),
                                                   B, _p(verts), _p(joints), _p(A), _p(ws), n, _stream()))
            return verts, joints, A
        L = _lib.lib()
        ro_flat = ro_ws = None
        nro = 0
        if readout is not None:
            ro_flat = torch.empty(B * readout.R * 3, dtype=torch.float32, device=self.device)
            nro = int(L.whmr_readout_workspace_bytes(readout._h, min(B, int(L.whmr_smpl_chunk_bodies(self._h)))))
            ro_ws = torch.empty(max(nro, 1), dtype=torch.uint8, device=self.device)
        deferred = C.c_int(0)
        with torch.cuda.device(self.device):
            check(L.whmr_smpl_forward_regressor(self._h, _p(betas), _p(pose), int(bool(pose_is_rotmat)), _p(transl), B,
                                                _p(verts), _p(joints), _p(A), None if readout is None else readout._h,
                                                _p(ro_flat), _p(ro_ws), nro, int(bool(defer_finish)), C.byref(deferred),
                                                None if g is None else C.byref(g), _p(ws), n, _stream()))
        if readout is None:
            return (verts, joints, A) + extra
        return (verts, joints, A, ro_flat, (ro_ws if deferred.value else ro_ws[:0])) + extra

    # per-stage launches (bench.py per-kernel timing, tests)
    def backward(self, betas, pose, g_verts=None, g_joints=None):
        """Rotation-matrix mode: (g_verts [B,V,3] | None, g_joints [B,J,3] | None) -> (g_betas [B,NB], g_pose [B,J,9])."""
        betas = _req(betas, "betas", align=16)
        pose = _req(pose, "pose", align=16)
        g_verts, g_joints = _req(g_verts, "g_verts"), _req(g_joints, "g_joints")
        B = betas.shape[0]
        if pose.numel() != B * self.J * 9:
            raise ValueError("smpl backward needs rotation matrices [B,%d,3,3], got %s" % (self.J, tuple(pose.shape)))
        g_betas = torch.empty(B, self.NB, dtype=torch.float32, device=self.device)
        g_pose = torch.empty(B, self.J, 3, 3, dtype=torch.float32, device=self.device)
        L = _lib.lib()
        n = int(L.whmr_smpl_backward_workspace_bytes(self._h, B))
        ws = torch.empty(n, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            check(L.whmr_smpl_backward(self._h, _p(betas), _p(pose), B, _p(g_verts), _p(g_joints), _p(g_betas), _p(g_pose),
                                       _p(ws), n, _stream()))
        return g_betas, g_pose

    def stage_chain(self, betas, pose, pose_is_rotmat, ws, n, joints=None, A=None, transl=None):
        check(_lib.lib().whmr_smpl_stage_chain(self._h, _p(betas), _p(pose), int(bool(pose_is_rotmat)), _p(transl),
                                               betas.shape[0], _p(joints), _p(A), _p(ws), n, _stream()))

    def stage_pose_blend(self, B, ws, n):
        check(_lib.lib().whmr_smpl_stage_pose_blend(self._h, int(B), _p(ws), n, _stream()))

    def stage_skin(self, betas, verts, ws, n):
        check(_lib.lib().whmr_smpl_stage_skin(self._h, _p(betas), betas.shape[0], _p(verts), _p(ws), n, _stream()))

    def forward_host(self, betas_h, pose_h, pose_is_rotmat, verts_h, joints_h=None):
        """HOST (ideally pinned) tensors in and out; synchronous (bench.py e2e leg)."""
        B = betas_h.shape[0]
        with torch.cuda.device(self.device):
            check(_lib.lib().whmr_smpl_forward_host(self._h, betas_h.data_ptr(), pose_h.data_ptr(),
                                                    int(bool(pose_is_rotmat)), B, verts_h.data_ptr(),
                                                    None if joints_h is None else joints_h.data_ptr(), _stream()))


def batch_rodrigues(aa):
    """smplx-variant Rodrigues (the one inside SMPL.forward). aa [n,3] -> [n,3,3]."""
    aa = _req(aa, "rot_vecs")
    n = aa.shape[0]
    out = torch.empty(n, 3, 3, dtype=torch.float32, device=aa.device)
    with torch.cuda.device(aa.device):
        check(_lib.lib().whmr_batch_rodrigues(_p(aa), n, _p(out), _stream()))
    return out


def _rot_op(fn_name, x, in_elems, out_shape):
    x = _req(x, "x")
    n = x.numel() // in_elems
    out = torch.empty((n,) + out_shape, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(getattr(_lib.lib(), fn_name)(_p(x), n, _p(out), _stream()))
    return out


def rot6d_to_rotmat(x):
    """utils/geometry.py:243-257: [..,6] -> [n,3,3]"""
    return _rot_op("whmr_rot6d_to_rotmat", x, 6, (3, 3))


def unbiased_gram_schmidt(x):
    """utils/geometry.py:260-272: [B,k,3,3] -> [B,k,3,3]"""
    k = x.shape[1]
    return _rot_op("whmr_unbiased_gram_schmidt", x, 9, (3, 3)).reshape(-1, k, 3, 3)


def rotation_matrix_to_angle_axis(R):
    """utils/geometry.py:54-83: [N,3,3] -> [N,3]"""
    return _rot_op("whmr_rotmat_to_axis_angle", R, 9, (3,))


def batch_rodrigues_quat(theta):
    """utils/geometry.py:14-51 (the quaternion variant, core/trainer.py:244): [N,3] -> [N,3,3]"""
    return _rot_op("whmr_batch_rodrigues_quat", theta, 3, (3, 3))


# ----------------------------------------------------------------------------------------------
# read-out
# ----------------------------------------------------------------------------------------------
_READOUTS = weakref.WeakValueDictionary()   # weak for the same reason as _HANDLES


class Readout:
    """Sparse linear read-out table (whmr_readout_create).  Built from named groups; each group
    is a dense [R_g, n_src] matrix, a list of source indices (picks) or a scipy CSR matrix, over
    the source set [V vertices ; J chain joints]."""

    def __init__(self, groups, n_verts, n_joints, device, sub_rows=None):
        import scipy.sparse as sp
        self.names, self.sizes = [], []
        mats = []
        n_src = n_verts + n_joints
        for name, g in groups:
            if sp.issparse(g):
                m = sp.csr_matrix(g, dtype=np.float64)
                if m.shape[1] < n_src:
                    m = sp.hstack([m, sp.csr_matrix((m.shape[0], n_src - m.shape[1]))]).tocsr()
            else:
                g = np.asarray(g)
                if g.ndim == 1:   # index picks
                    m = sp.csr_matrix((np.ones(len(g)), (np.arange(len(g)), g.astype(np.int64))),
                                      shape=(len(g), n_src))
                else:
                    d = np.zeros((g.shape[0], n_src), dtype=np.float64)
                    d[:, :g.shape[1]] = g
                    m = sp.csr_matrix(d)
            m.sort_indices()
            mats.append(m)
            self.names.append(name)
            self.sizes.append(m.shape[0])
        full = sp.vstack(mats).tocsr() if mats else sp.csr_matrix((0, n_src))
        full.sort_indices()
        self.R = full.shape[0]
        self.csr = full
        row_ptr = np.ascontiguousarray(full.indptr.astype(np.int32))
        col = np.ascontiguousarray(full.indices.astype(np.int32))
        vals = np.ascontiguousarray(full.data.astype(np.float32))
        sub = None
        if sub_rows is not None:
            sub = np.ascontiguousarray(np.asarray(sub_rows, dtype=np.int32))
            assert sub.shape[0] == self.R
        gs = np.ascontiguousarray(np.asarray(self.sizes, dtype=np.int32))
        self.device = torch.device(device)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(_lib.lib().whmr_readout_create(self.R, n_verts, n_joints, row_ptr.ctypes.data,
                                                 col.ctypes.data if len(col) else None,
                                                 vals.ctypes.data if len(vals) else None,
                                                 None if sub is None else sub.ctypes.data,
                                                 len(gs), gs.ctypes.data if len(gs) else None, C.byref(self._h)))
        self.V, self.J = n_verts, n_joints
        self.id = id(self)
        _READOUTS[self.id] = self

    def __del__(self):
        try:
            _READOUTS.pop(self.id, None)
            if self._h.value:
                _lib.lib().whmr_readout_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def split(self, flat, B):
        """group-major flat buffer -> dict name -> contiguous [B, R_g, 3] view"""
        res, off = {}, 0
        for name, r in zip(self.names, self.sizes):
            res[name] = flat[off:off + B * r * 3].view(B, r, 3)
            off += B * r * 3
        return res

    def backward(self, g_flat, g_verts, g_joints=None):
        """Accumulate the gradient of every read-out row (flat group-major g_flat) into g_verts [B,V,3] / g_joints."""
        g_flat = _req(g_flat, "g_flat")
        B = g_verts.shape[0]
        with torch.cuda.device(g_verts.device):
            check(_lib.lib().whmr_readout_backward(self._h, _p(g_flat), B, _p(g_verts), _p(g_joints), _stream()))

    def apply(self, verts, joints=None):
        """-> dict name -> contiguous [B, R_g, 3] tensor (all from one launch pair)."""
        verts = _req(verts, "verts")
        joints = _req(joints, "joints")
        B = verts.shape[0]
        out = torch.empty(B * self.R * 3, dtype=torch.float32, device=verts.device)
        with torch.cuda.device(verts.device):
            check(_lib.lib().whmr_readout_apply(self._h, _p(verts), _p(joints), B, _p(out), _stream()))
        return self.split(out, B)


def gather_vertices(verts, idx):
    verts = _req(verts, "verts")
    idx = _req(idx, "idx", dtype=torch.int32)
    B, V = verts.shape[0], verts.shape[1]
    out = torch.empty(B, idx.numel(), 3, dtype=torch.float32, device=verts.device)
    with torch.cuda.device(verts.device):
        check(_lib.lib().whmr_gather_vertices(_p(verts), _p(idx), B, V, idx.numel(), _p(out), _stream()))
    return out


# ----------------------------------------------------------------------------------------------
# projection
# ----------------------------------------------------------------------------------------------
def project_weak(points, cam, focal, img_w, img_h):
    points = _req(points, "points")
    cam = _req(cam, "cam")
    B, N = points.shape[0], points.shape[1]
    out = torch.empty(B, N, 2, dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        check(_lib.lib().whmr_project_weak(_p(points), _p(cam), B, N, float(focal), float(img_w), float(img_h),
                                           _p(out), _stream()))
    return out


def perspective_projection(points, rotation, translation, focal_length, camera_center, retain_z=False,
                           distortion=None):
    points = _req(points, "points")
    B, N = points.shape[0], points.shape[1]
    dev = points.device
    rot_batch = 0
    if rotation is not None:
        rotation = _req(rotation, "rotation")
        rot_batch = rotation.shape[0] if rotation.dim() == 3 else 1
        if rot_batch not in (1, B):
            raise ValueError("rotation batch %d must be 1 or %d" % (rot_batch, B))
    translation = _req(translation, "translation")
    distortion = _req(distortion, "distortion")
    focal_dev, focal_scalar = None, 0.0
    if torch.is_tensor(focal_length) and focal_length.numel() > 1:
        focal_dev = _req(focal_length.reshape(-1), "focal_length")
    else:
        focal_scalar = float(focal_length)
    if not torch.is_tensor(camera_center):
        camera_center = torch.as_tensor(camera_center, dtype=torch.float32)
    camera_center = _req(camera_center.to(dev), "camera_center")
    out = torch.empty(B, N, 3 if retain_z else 2, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().whmr_perspective_projection(_p(points), _p(rotation), rot_batch, _p(translation),
                                                     _p(focal_dev), focal_scalar, _p(camera_center),
                                                     _p(distortion), B, N, int(bool(retain_z)), _p(out),
                                                     _stream()))
    return out


def estimate_translation(S, joints_2d, focal_length=5000., img_size=(224., 224.), first_joint=25):
    """utils/geometry.py:386-408 on the device (no host round trip): S [B,N,3], joints_2d [B,N,3] = (x, y, conf);
    joints [first_joint, N) take part.  -> [B,3]."""
    S, joints_2d = _req(S, "S"), _req(joints_2d, "joints_2d")
    B, N = S.shape[0], S.shape[1]
    if joints_2d.shape != (B, N, 3) or S.shape[2] != 3:
        raise ValueError("estimate_translation: S %s / joints_2d %s" % (tuple(S.shape), tuple(joints_2d.shape)))
    out = torch.empty(B, 3, dtype=torch.float32, device=S.device)
    with torch.cuda.device(S.device):
        check(_lib.lib().whmr_estimate_translation(_p(S), _p(joints_2d), B, N, min(int(first_joint), N), float(focal_length),
                                                   float(img_size[0]), float(img_size[1]), _p(out), _stream()))
    return out


def project_full(points, cam, bbox_height, center, orig_shape, Tz, want_px=False):
    """models/whmr.py:147-173 fused.  -> (kp_norm [B,N,2], focal [B], cam_t [B,3], kp_px or None)"""
    points = _req(points, "points")
    B, N = points.shape[0], points.shape[1]
    dev = points.device
    cam, bbox_height, center = _req(cam, "cam"), _req(bbox_height, "bbox_height"), _req(center, "center")
    orig_shape, Tz = _req(orig_shape, "orig_shape"), _req(Tz, "Tz")
    kp = torch.empty(B, N, 2, dtype=torch.float32, device=dev)
    px = torch.empty(B, N, 2, dtype=torch.float32, device=dev) if want_px else None
    focal = torch.empty(B, dtype=torch.float32, device=dev)
    cam_t = torch.empty(B, 3, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().whmr_project_full(_p(points), _p(cam), _p(bbox_height), _p(center), _p(orig_shape), _p(Tz),
                                           B, N, _p(kp), _p(px), _p(focal), _p(cam_t), _stream()))
    return kp, focal, cam_t, px


def project_weak_full(points, cam, bbox_height, center, orig_shape, Tz, focal, img_w, img_h):
    """weak projection + predicted-focal block in one launch -> (kp_2d, kp_2d_w, focal_length, cam_t)"""
    points = _req(points, "points")
    B, N = points.shape[0], points.shape[1]
    dev = points.device
    cam, bbox_height, center = _req(cam, "cam"), _req(bbox_height, "bbox_height"), _req(center, "center")
    orig_shape, Tz = _req(orig_shape, "orig_shape"), _req(Tz, "Tz")
    kp = torch.empty(B, N, 2, dtype=torch.float32, device=dev)
    kpw = torch.empty(B, N, 2, dtype=torch.float32, device=dev)
    fl = torch.empty(B, dtype=torch.float32, device=dev)
    cam_t = torch.empty(B, 3, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().whmr_project_weak_full(_p(points), _p(cam), _p(bbox_height), _p(center), _p(orig_shape), _p(Tz),
                                                B, N, float(focal), float(img_w), float(img_h), _p(kp), _p(kpw), None,
                                                _p(fl), _p(cam_t), _stream()))
    return kp, kpw, fl, cam_t


def project_crop(points, cam, center, scale, img_focal, img_center, crop_size, img_w, img_h, distortion=None):
    points = _req(points, "points")
    B, N = points.shape[0], points.shape[1]
    dev = points.device
    args = [_req(x, n) for x, n in ((cam, "cam"), (center, "center"), (scale, "scale"),
                                    (img_focal, "img_focal"), (img_center, "img_center"))]
    distortion = _req(distortion, "distortion")
    full = torch.empty(B, N, 2, dtype=torch.float32, device=dev)
    crop = torch.empty(B, N, 2, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().whmr_project_crop(_p(points), *[_p(a) for a in args], _p(distortion), B, N,
                                           float(crop_size), float(img_w), float(img_h), _p(full), _p(crop), _stream()))
    return full, crop


# ----------------------------------------------------------------------------------------------
# sampling
# ----------------------------------------------------------------------------------------------
def is_host_map(feat):
    """True for a feature map in PAGE-LOCKED host memory (`tensor.pin_memory()`): the sampling kernels read such a map in
    place through unified addressing -- only the sectors the taps touch cross PCIe, never the whole map."""
    return torch.is_tensor(feat) and feat.device.type == "cpu" and feat.is_pinned()


def _feat_layout(feat, layout):
    """-> (tensor whose memory is contiguous in the kernel's layout, layout, B, C, H, W).
    A [B,C,H,W] tensor in torch.channels_last memory format IS an NHWC array in memory: it is
    handed to the NHWC kernel as is (no copy), so a backbone run in channels_last gets the 4x lower
    sampling traffic for free.
    A pinned host tensor (is_host_map) is passed through as it is: it must already be fp32 and contiguous in the layout
    named (nothing is copied or converted on the host: there is no CPU path)."""
    if is_host_map(feat):
        if feat.requires_grad and torch.is_grad_enabled():
            raise NotImplementedError("a host-resident (pinned) feature map is read in place and receives no gradient: "
                                      "inference only -- move the map to the device to train through the sampling")
        if feat.dtype != torch.float32 or feat.dim() != 4:
            raise _lib.WhmrError("a host-resident feature map must be a pinned fp32 [B,C,H,W] / [B,H,W,C] tensor")
        if layout == LAYOUT_NCHW and not feat.is_contiguous() and feat.is_contiguous(memory_format=torch.channels_last):
            B, Cc, H, W = feat.shape
            return feat, LAYOUT_NHWC, B, Cc, H, W
        if not feat.is_contiguous():
            raise _lib.WhmrError("a host-resident feature map must be contiguous (or channels_last): nothing is copied on the host")
        if layout == LAYOUT_NCHW:
            B, Cc, H, W = feat.shape
        else:
            B, H, W, Cc = feat.shape
        return feat, layout, B, Cc, H, W
    if not torch.is_tensor(feat) or not feat.is_cuda:
        _req(feat, "im_feat")
    if layout == LAYOUT_NCHW:
        B, Cc, H, W = feat.shape
        if feat.dtype == torch.float32 and not feat.is_contiguous() and \
                feat.is_contiguous(memory_format=torch.channels_last) and feat.data_ptr() % 4 == 0:
            return feat, LAYOUT_NHWC, B, Cc, H, W
        return _req(feat, "im_feat"), LAYOUT_NCHW, B, Cc, H, W
    B, H, W, Cc = feat.shape
    return _req(feat, "im_feat"), LAYOUT_NHWC, B, Cc, H, W


def sample_bilinear(feat, points, layout=LAYOUT_NCHW):
    """grid_sample(feat, points[:, :, None, :], align_corners=True)[..., 0]  ->  [B,C,N]
    `feat`: CUDA tensor, or a pinned host tensor gathered in place (is_host_map); the result lives on `points.device`."""
    feat, layout, B, Cc, H, W = _feat_layout(feat, layout)
    points = _req(points, "points", align=8)
    shared = points.dim() == 2 or (points.shape[0] == 1 and B != 1)   # one [N,2] grid for every body
    N = points.shape[-2]
    if (not shared and points.shape[0] != B) or points.shape[-1] != 2:
        raise ValueError("points %s do not match feature batch %d" % (tuple(points.shape), B))
    out = torch.empty(B, Cc, N, dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        check(_lib.lib().whmr_sample_bilinear(_p(feat), int(layout), B, Cc, H, W, _p(points), int(shared), N,
                                              _p(out), _stream()))
    return out


def project_sample(feat, p, cam, focal, img_w, img_h, layout=LAYOUT_NCHW):
    """MAF_Extractor.forward's projection + sampling.  -> (point_feat [B,C,N], points2d [B,N,2])
    `feat`: CUDA tensor, or a pinned host tensor gathered in place (is_host_map); the results live on `p.device`."""
    feat, layout, B, Cc, H, W = _feat_layout(feat, layout)
    p = _req(p, "p")
    cam = _req(cam, "cam")
    N = p.shape[1]
    pts2d = torch.empty(B, N, 2, dtype=torch.float32, device=p.device)
    out = torch.empty(B, Cc, N, dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        check(_lib.lib().whmr_project_sample(_p(feat), int(layout), B, Cc, H, W, _p(p), _p(cam), N, float(focal),
                                             float(img_w), float(img_h), _p(pts2d), _p(out), _stream()))
    return out, pts2d


class MafMlp:
    """Device-side state of the fused sampling + `reduce_dim` kernel (whmr_maf_mlp_create): the tf32 hi|lo split of the
    module's conv0..2 weights, re-split whenever a parameter's storage or version counter changes (optimizer step,
    load_state_dict, .to())."""

    def __init__(self, dims, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.WhmrError("MafMlp needs a CUDA device, got %s (no CPU fallback)" % (self.device,))
        self.dims = tuple(int(d) for d in dims)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(_lib.lib().whmr_maf_mlp_create(*self.dims, C.byref(self._h)))
        self._key = None

    @staticmethod
    def supported(dims):
        return len(dims) == 4 and dims[0] % 64 == 0 and dims[1] % 64 == 0 and dims[2] % 64 == 0 and dims[3] % 16 == 0 \
            and min(dims) >= 16 and dims[1] + dims[2] + dims[3] <= 256

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().whmr_maf_mlp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_weights(self, convs):
        """convs: the three Conv1d modules (weight [C_out, C_in_total, 1], bias [C_out] or None)."""
        ts = []
        for c in convs:
            ts += [c.weight, c.bias]
        key = tuple((None if t is None else (t.data_ptr(), t._version, tuple(t.shape))) for t in ts)
        if key == self._key:
            return
        c0, c1, c2, c3 = self.dims
        want = [(c1, c0, 1), (c2, c1 + c0, 1), (c3, c2 + c0, 1)]
        for c, w in zip(convs, want):
            if tuple(c.weight.shape) != w:
                raise ValueError("conv weight %s does not match the MLP widths %s" % (tuple(c.weight.shape), self.dims))
        ts = [None if t is None else _req(t.detach(), "conv parameter") for t in ts]
        with torch.cuda.device(self.device):
            check(_lib.lib().whmr_maf_mlp_set_weights(self._h, *[_p(t) for t in ts], _stream()))
        self._key = key

    def sample(self, feat, points, layout=LAYOUT_NCHW, want_point_feat=True):
        """MAF_Extractor.sampling in one launch -> (mesh_align_feat [B, C3*N], point_feat [B,C0,N] or None)"""
        feat, layout, B, Cc, H, W = _feat_layout(feat, layout)
        if Cc != self.dims[0]:
            raise ValueError("feature maps have %d channels, the MLP expects %d" % (Cc, self.dims[0]))
        points = _req(points, "points", align=8)
        shared = points.dim() == 2 or (points.shape[0] == 1 and B != 1)
        N = points.shape[-2]
        if (not shared and points.shape[0] != B) or points.shape[-1] != 2:
            raise ValueError("points %s do not match feature batch %d" % (tuple(points.shape), B))
        out = torch.empty(B, self.dims[3] * N, dtype=torch.float32, device=feat.device)
        pf = torch.empty(B, Cc, N, dtype=torch.float32, device=feat.device) if want_point_feat else None
        with torch.cuda.device(feat.device):
            check(_lib.lib().whmr_sample_reduce(self._h, _p(feat), int(layout), B, H, W, _p(points), int(shared), N,
                                                _p(out), _p(pf), _stream()))
        return out, pf

    def project_sample(self, feat, p, cam, focal, img_w, img_h, layout=LAYOUT_NCHW, want_point_feat=True,
                       want_points2d=False):
        """MAF_Extractor.forward in one launch (weak projection + sampling + MLP)"""
        feat, layout, B, Cc, H, W = _feat_layout(feat, layout)
        if Cc != self.dims[0]:
            raise ValueError("feature maps have %d channels, the MLP expects %d" % (Cc, self.dims[0]))
        p = _req(p, "p")
        cam = _req(cam, "cam")
        N = p.shape[1]
        out = torch.empty(B, self.dims[3] * N, dtype=torch.float32, device=feat.device)
        pf = torch.empty(B, Cc, N, dtype=torch.float32, device=feat.device) if want_point_feat else None
        pts2d = torch.empty(B, N, 2, dtype=torch.float32, device=feat.device) if want_points2d else None
        with torch.cuda.device(feat.device):
            check(_lib.lib().whmr_project_sample_reduce(self._h, _p(feat), int(layout), B, H, W, _p(p), _p(cam), N,
                                                        float(focal), float(img_w), float(img_h), _p(pts2d), _p(out),
                                                        _p(pf), _stream()))
        return out, pf, pts2d


# ----------------------------------------------------------------------------------------------
# metrics
# ----------------------------------------------------------------------------------------------
def joint_errors(pred, gt, want_pa=True):
    """-> (mpjpe [n], pa_mpjpe [n] or None) in the units of the inputs."""
    pred, gt = _req(pred, "pred"), _req(gt, "gt")
    n, J = pred.shape[0], pred.shape[1]
    mp = torch.empty(n, dtype=torch.float32, device=pred.device)
    pa = torch.empty(n, dtype=torch.float32, device=pred.device) if want_pa else None
    with torch.cuda.device(pred.device):
        check(_lib.lib().whmr_joint_errors(_p(pred), _p(gt), n, J, _p(mp), _p(pa), _stream()))
    return mp, pa


def vertex_errors(pred, gt):
    """PVE per frame (evaluate/eval.py:208-209): mean_v ||pred - gt||, pred/gt [n,V,3] -> [n]."""
    pred, gt = _req(pred, "pred"), _req(gt, "gt")
    if pred.shape != gt.shape:
        raise ValueError("vertex_errors: shapes differ %s vs %s" % (tuple(pred.shape), tuple(gt.shape)))
    n, V = pred.shape[0], pred.shape[1]
    out = torch.empty(n, dtype=torch.float32, device=pred.device)
    with torch.cuda.device(pred.device):
        check(_lib.lib().whmr_vertex_errors(_p(pred), _p(gt), n, V, _p(out), _stream()))
    return out


# ----------------------------------------------------------------------------------------------
# torch.library custom ops (the "PyTorch custom op" boundary named by the north star).  The
# module-level drop-ins (smpl.SMPL, geometry.projection, maf_extractor.MAF_Extractor) call these,
# so the ops show up by name in profiler traces and can be captured in CUDA graphs.
# ----------------------------------------------------------------------------------------------
@torch.library.custom_op("whmr::smpl_lbs", mutates_args=(), device_types="cuda")
def smpl_lbs(handle: int, betas: torch.Tensor, pose: torch.Tensor, pose_is_rotmat: bool) -> \
        tuple[torch.Tensor, torch.Tensor]:
    v, j, _ = _HANDLES[handle].forward(betas, pose, pose_is_rotmat)
    return v, j


@smpl_lbs.register_fake
def _(handle, betas, pose, pose_is_rotmat):
    h = _HANDLES[handle]
    B = betas.shape[0]
    return betas.new_empty(B, h.V, 3), betas.new_empty(B, h.J, 3)


@torch.library.custom_op("whmr::smpl_lbs_readout", mutates_args=(), device_types="cuda")
def smpl_lbs_readout(handle: int, readout: int, betas: torch.Tensor, pose: torch.Tensor, pose_is_rotmat: bool) -> \
        tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """SMPL forward + all read-outs of a Readout table (flat, group-major) in one chunked pass."""
    v, j, _, flat, _ws = _HANDLES[handle].forward(betas, pose, pose_is_rotmat, readout=_READOUTS[readout])
    return v, j, flat


@smpl_lbs_readout.register_fake
def _(handle, readout, betas, pose, pose_is_rotmat):
    h, ro = _HANDLES[handle], _READOUTS[readout]
    B = betas.shape[0]
    return betas.new_empty(B, h.V, 3), betas.new_empty(B, h.J, 3), betas.new_empty(B * ro.R * 3)


def _smpl_need_rotmat(pose_is_rotmat):
    if not pose_is_rotmat:
        raise NotImplementedError("whmr SMPL backward is implemented for rotation-matrix input (pose2rot=False, the "
                                  "model path models/whmr.py:132-137); axis-angle input is used for ground truth only")


def _smpl_setup(ctx, inputs, output):
    handle, betas, pose, pose_is_rotmat = inputs
    ctx.save_for_backward(betas, pose)
    ctx.handle, ctx.rotmat = handle, pose_is_rotmat


def _smpl_bwd(ctx, g_v, g_j):
    _smpl_need_rotmat(ctx.rotmat)
    betas, pose = ctx.saved_tensors
    g_betas, g_pose = _HANDLES[ctx.handle].backward(betas, pose, g_v, g_j)
    return None, g_betas, g_pose.view_as(pose), None


smpl_lbs.register_autograd(_smpl_bwd, setup_context=_smpl_setup)


def _smpl_ro_setup(ctx, inputs, output):
    handle, readout, betas, pose, pose_is_rotmat = inputs
    ctx.save_for_backward(betas, pose)
    ctx.handle, ctx.readout, ctx.rotmat = handle, readout, pose_is_rotmat


def _smpl_ro_bwd(ctx, g_v, g_j, g_flat):
    _smpl_need_rotmat(ctx.rotmat)
    betas, pose = ctx.saved_tensors
    h, ro = _HANDLES[ctx.handle], _READOUTS[ctx.readout]
    B = betas.shape[0]
    if g_flat is not None:   # fold the read-out gradients into those of the vertices / chain joints
        g_v = g_v.contiguous().clone() if g_v is not None else torch.zeros(B, h.V, 3, dtype=torch.float32, device=betas.device)
        g_j = g_j.contiguous().clone() if g_j is not None else torch.zeros(B, h.J, 3, dtype=torch.float32, device=betas.device)
        ro.backward(g_flat.contiguous(), g_v, g_j)
    g_betas, g_pose = h.backward(betas, pose, g_v, g_j)
    return None, None, g_betas, g_pose.view_as(pose), None


smpl_lbs_readout.register_autograd(_smpl_ro_bwd, setup_context=_smpl_ro_setup)


@torch.library.custom_op("whmr::smpl_lbs_readout_deferred", mutates_args=(), device_types="cuda")
def smpl_lbs_readout_deferred(handle: int, readout: int, betas: torch.Tensor, pose: torch.Tensor,
                              pose_is_rotmat: bool) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """As smpl_lbs_readout, but the finishing pass of the read-outs (regressor rows, joint copies) is left to
    `readout_finish` when the batch is a single chunk: then the 4th output (scratch) is non-empty.  The vertex
    one-hot rows (markers, picks, down-sampled meshes) are complete on return either way."""
    v, j, _, flat, ws = _HANDLES[handle].forward(betas, pose, pose_is_rotmat, readout=_READOUTS[readout],
                                                 defer_finish=True)
    return v, j, flat, ws


@smpl_lbs_readout_deferred.register_fake
def _(handle, readout, betas, pose, pose_is_rotmat):
    h, ro = _HANDLES[handle], _READOUTS[readout]
    B = betas.shape[0]
    return (betas.new_empty(B, h.V, 3), betas.new_empty(B, h.J, 3), betas.new_empty(B * ro.R * 3),
            betas.new_empty(0, dtype=torch.uint8))


@torch.library.custom_op("whmr::smpl_regressor", mutates_args=(), device_types="cuda")
def smpl_regressor(handle: int, readout: int, betas: torch.Tensor, rotmat: torch.Tensor, cam: torch.Tensor,
                   orthonormalize: bool, defer: bool) -> \
        tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """The SMPL part of Regressor.forward / forward_init in one chunked pass (models/whmr.py:128-190): optional
    unbiased_gram_schmidt of the predicted rotations (eval mode), SMPL forward, every read-out of the table, `pose`
    (rotation_matrix_to_angle_axis) and `theta` = cat(cam, betas, pose).
    -> (verts, chain joints, flat read-outs, read-out scratch (non-empty iff the finishing pass was deferred),
        rotmat used [B,J,3,3], pose [B,3J], theta [B,3+NB+3J]).  Differentiable w.r.t. betas / rotmat through verts, joints
    and the read-outs when orthonormalize is False (training); pose / theta / rotmat carry no gradient."""
    v, j, _, flat, ws, R, pose, theta = _HANDLES[handle].forward(
        betas, rotmat, True, readout=_READOUTS[readout], defer_finish=defer,
        glue={"gram_schmidt": orthonormalize, "cam": cam})
    return v, j, flat, ws, R, pose, theta


@smpl_regressor.register_fake
def _(handle, readout, betas, rotmat, cam, orthonormalize, defer):
    h, ro = _HANDLES[handle], _READOUTS[readout]
    B = betas.shape[0]
    return (betas.new_empty(B, h.V, 3), betas.new_empty(B, h.J, 3), betas.new_empty(B * ro.R * 3),
            betas.new_empty(0, dtype=torch.uint8), betas.new_empty(B, h.J, 3, 3), betas.new_empty(B, h.J * 3),
            betas.new_empty(B, 3 + h.NB + h.J * 3))


def _smpl_reg_setup(ctx, inputs, output):
    handle, readout, betas, rotmat, cam, orthonormalize, defer = inputs
    ctx.save_for_backward(betas, rotmat)
    ctx.handle, ctx.readout, ctx.orth = handle, readout, orthonormalize


def _smpl_reg_bwd(ctx, g_v, g_j, g_flat, g_ws, g_R, g_pose, g_theta):
    if ctx.orth:
        raise NotImplementedError("whmr::smpl_regressor: no backward through unbiased_gram_schmidt (the reference applies it "
                                  "in eval mode only, models/whmr.py:129-130)")
    betas, pose = ctx.saved_tensors
    h, ro = _HANDLES[ctx.handle], _READOUTS[ctx.readout]
    B = betas.shape[0]
    if g_flat is not None:
        g_v = g_v.contiguous().clone() if g_v is not None else torch.zeros(B, h.V, 3, dtype=torch.float32, device=betas.device)
        g_j = g_j.contiguous().clone() if g_j is not None else torch.zeros(B, h.J, 3, dtype=torch.float32, device=betas.device)
        ro.backward(g_flat.contiguous(), g_v, g_j)
    g_betas, g_pose_in = h.backward(betas, pose, g_v, g_j)
    return None, None, g_betas, g_pose_in.view_as(pose), None, None, None


smpl_regressor.register_autograd(_smpl_reg_bwd, setup_context=_smpl_reg_setup)


@torch.library.custom_op("whmr::readout_finish", mutates_args=("flat",), device_types="cuda")
def readout_finish(readout: int, joints: torch.Tensor, flat: torch.Tensor, scratch: torch.Tensor) -> None:
    ro = _READOUTS[readout]
    with torch.cuda.device(flat.device):
        check(_lib.lib().whmr_readout_finish(ro._h, _p(joints), joints.shape[0], _p(scratch), _p(flat), _stream()))


def readout_finish_multi(readout, joints, flats, scratches, proj=None):
    """The deferred finishing passes of several smpl_lbs_readout_deferred calls (same table, same batch) in one launch;
    writes the regressor rows into each `flat` in place.  (Plain function: inference-side scheduling, no autograd.)

    proj: fold the projections of one read-out group (the 49 joints) into the same launch -- dict(group=name, cams=[cam
    [B,3] or None per call], full=[bool per call], bbox_height, center, orig_shape, Tz, focal, img_w, img_h).  Returns one
    (kp_2d, kp_2d_w, focal_length, cam_t) tuple per call (kp_2d only / None where not full / no camera)."""
    ro = _READOUTS[readout]
    n = len(flats)
    if n == 0:
        return []
    B = joints[0].shape[0]
    dev = flats[0].device
    arr = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])  # noqa: E731
    outs = [None] * n
    fp = None
    keep = []
    if proj is not None:
        gi = ro.names.index(proj['group'])
        fp = _lib.FinishProjection()
        fp.row0, fp.n_points = int(sum(ro.sizes[:gi])), int(ro.sizes[gi])
        fp.focal, fp.img_w, fp.img_h = float(proj['focal']), float(proj['img_w']), float(proj['img_h'])
        if any(proj['full']):
            bb = [_req(proj[k], k) for k in ('bbox_height', 'center', 'orig_shape', 'Tz')]
            keep += bb
            fp.bbox_height, fp.center, fp.orig_shape, fp.Tz = [t.data_ptr() for t in bb]
        N = fp.n_points
        for i in range(n):
            cam = proj['cams'][i]
            if cam is None:
                continue
            cam = _req(cam, "cam")
            keep.append(cam)
            fp.cam[i] = cam.data_ptr()
            kp = torch.empty(B, N, 2, dtype=torch.float32, device=dev)
            fp.kp_weak[i] = kp.data_ptr()
            if proj['full'][i]:
                kpw = torch.empty(B, N, 2, dtype=torch.float32, device=dev)
                fl = torch.empty(B, dtype=torch.float32, device=dev)
                ct = torch.empty(B, 3, dtype=torch.float32, device=dev)
                fp.full[i], fp.kp_norm[i], fp.focal_out[i], fp.cam_t_out[i] = 1, kpw.data_ptr(), fl.data_ptr(), ct.data_ptr()
                outs[i] = (kp, kpw, fl, ct)
            else:
                outs[i] = (kp, None, None, None)
    with torch.cuda.device(dev):
        check(_lib.lib().whmr_readout_finish_project_multi(ro._h, n, arr(joints), arr(scratches), arr(flats), B,
                                                           None if fp is None else C.byref(fp), _stream()))
    return outs


@torch.library.custom_op("whmr::sample_bilinear", mutates_args=(), device_types="cuda")
def sample_bilinear_op(feat: torch.Tensor, points: torch.Tensor, layout: int) -> torch.Tensor:
    return sample_bilinear(feat, points, layout)


@sample_bilinear_op.register_fake
def _(feat, points, layout):
    Cc = feat.shape[1] if layout == LAYOUT_NCHW else feat.shape[3]
    return feat.new_empty(feat.shape[0], Cc, points.shape[-2])


def _sample_backward(g_out, feat, points, layout):
    """gradient of sample_bilinear w.r.t. feat (same shape and memory layout as feat); points [B,N,2] or shared [N,2]"""
    feat_k, lay, B, Cc, H, W = _feat_layout(feat, layout)
    g_feat = torch.zeros_like(feat_k)            # preserves channels_last strides
    g_out = _req(g_out, "grad_out")
    points = _req(points, "points", align=8)
    shared = points.dim() == 2 or (points.shape[0] == 1 and B != 1)
    N = points.shape[-2]
    with torch.cuda.device(feat.device):
        check(_lib.lib().whmr_sample_bilinear_backward(_p(g_out), int(lay), B, Cc, H, W, _p(points), int(shared), N,
                                                       _p(g_feat), _stream()))
    return g_feat


def _sample_setup(ctx, inputs, output):
    feat, points, layout = inputs
    ctx.save_for_backward(feat, points)
    ctx.layout = layout


def _sample_bwd(ctx, g):
    feat, points = ctx.saved_tensors
    return _sample_backward(g, feat, points.detach(), ctx.layout), None, None


# the reference detaches the sampling points (models/whmr.py:586-591): the gradient goes to the feature maps only
sample_bilinear_op.register_autograd(_sample_bwd, setup_context=_sample_setup)


@torch.library.custom_op("whmr::project_sample", mutates_args=(), device_types="cuda")
def project_sample_op(feat: torch.Tensor, p: torch.Tensor, cam: torch.Tensor, focal: float, img_w: float, img_h: float,
                      layout: int) -> tuple[torch.Tensor, torch.Tensor]:
    """MAF_Extractor.forward: weak projection of the mesh points + sampling -> (point_feat [B,C,N], points2d [B,N,2])"""
    return project_sample(feat, p, cam, focal, img_w, img_h, layout)


@project_sample_op.register_fake
def _(feat, p, cam, focal, img_w, img_h, layout):
    Cc = feat.shape[1] if layout == LAYOUT_NCHW else feat.shape[3]
    return feat.new_empty(feat.shape[0], Cc, p.shape[1]), feat.new_empty(feat.shape[0], p.shape[1], 2)


def _psample_setup(ctx, inputs, output):
    feat, p, cam, focal, img_w, img_h, layout = inputs
    ctx.save_for_backward(feat, output[1])
    ctx.layout = layout


def _psample_bwd(ctx, g_feat_out, g_pts):
    feat, pts2d = ctx.saved_tensors
    return _sample_backward(g_feat_out, feat, pts2d.detach(), ctx.layout), None, None, None, None, None, None


project_sample_op.register_autograd(_psample_bwd, setup_context=_psample_setup)


@torch.library.custom_op("whmr::project_weak", mutates_args=(), device_types="cuda")
def project_weak_op(points: torch.Tensor, cam: torch.Tensor, focal: float, img_w: float, img_h: float) -> torch.Tensor:
    return project_weak(points, cam, focal, img_w, img_h)


@project_weak_op.register_fake
def _(points, cam, focal, img_w, img_h):
    return points.new_empty(points.shape[0], points.shape[1], 2)


@torch.library.custom_op("whmr::project_weak_backward", mutates_args=(), device_types="cuda")
def project_weak_backward(points: torch.Tensor, cam: torch.Tensor, g_out: torch.Tensor, focal: float, img_w: float,
                          img_h: float) -> tuple[torch.Tensor, torch.Tensor]:
    points, cam, g_out = _req(points, "points"), _req(cam, "cam"), _req(g_out, "grad_out")
    B, N = points.shape[0], points.shape[1]
    g_points = torch.empty_like(points)
    g_cam = torch.empty(B, 3, dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        check(_lib.lib().whmr_project_weak_backward(_p(points), _p(cam), _p(g_out), B, N, float(focal), float(img_w),
                                                    float(img_h), _p(g_points), _p(g_cam), _stream()))
    return g_points, g_cam


@project_weak_backward.register_fake
def _(points, cam, g_out, focal, img_w, img_h):
    return torch.empty_like(points), cam.new_empty(points.shape[0], 3)


def _pw_setup(ctx, inputs, output):
    points, cam, focal, img_w, img_h = inputs
    ctx.save_for_backward(points, cam)
    ctx.consts = (focal, img_w, img_h)


def _pw_bwd(ctx, g):
    points, cam = ctx.saved_tensors
    g_points, g_cam = project_weak_backward(points, cam, g.contiguous(), *ctx.consts)
    return g_points, g_cam, None, None, None


project_weak_op.register_autograd(_pw_bwd, setup_context=_pw_setup)


@torch.library.custom_op("whmr::project_weak_full", mutates_args=(), device_types="cuda")
def project_weak_full_op(points: torch.Tensor, cam: torch.Tensor, bbox_height: torch.Tensor, center: torch.Tensor,
                         orig_shape: torch.Tensor, Tz: torch.Tensor, focal: float, img_w: float, img_h: float) -> \
        tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """weak projection + predicted-focal block (models/whmr.py:142-173) -> (kp_2d, kp_2d_w, focal_length, cam_t)"""
    return project_weak_full(points, cam, bbox_height, center, orig_shape, Tz, focal, img_w, img_h)


@project_weak_full_op.register_fake
def _(points, cam, bbox_height, center, orig_shape, Tz, focal, img_w, img_h):
    B, N = points.shape[0], points.shape[1]
    return points.new_empty(B, N, 2), points.new_empty(B, N, 2), points.new_empty(B), points.new_empty(B, 3)


@torch.library.custom_op("whmr::project_full_backward", mutates_args=(), device_types="cuda")
def project_full_backward(points: torch.Tensor, cam: torch.Tensor, bbox_height: torch.Tensor, center: torch.Tensor,
                          orig_shape: torch.Tensor, Tz: torch.Tensor, focal: float, img_w: float, img_h: float,
                          g_kp_weak: torch.Tensor, g_kp_norm: torch.Tensor, g_focal: torch.Tensor,
                          g_cam_t: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    points = _req(points, "points")
    B, N = points.shape[0], points.shape[1]
    args = [_req(x, n) for x, n in ((cam, "cam"), (bbox_height, "bbox_height"), (center, "center"),
                                    (orig_shape, "orig_shape"), (Tz, "Tz"))]
    gs = [_req(x, n) for x, n in ((g_kp_weak, "g_kp_weak"), (g_kp_norm, "g_kp_norm"), (g_focal, "g_focal"),
                                  (g_cam_t, "g_cam_t"))]
    g_points = torch.empty_like(points)
    g_cam = torch.empty(B, 3, dtype=torch.float32, device=points.device)
    g_Tz = torch.empty(B, dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        check(_lib.lib().whmr_project_full_backward(_p(points), *[_p(a) for a in args], B, N, float(focal), float(img_w),
                                                    float(img_h), *[_p(g) for g in gs], _p(g_points), _p(g_cam), _p(g_Tz),
                                                    _stream()))
    return g_points, g_cam, g_Tz


@project_full_backward.register_fake
def _(points, cam, bbox_height, center, orig_shape, Tz, focal, img_w, img_h, g_kp_weak, g_kp_norm, g_focal, g_cam_t):
    return torch.empty_like(points), cam.new_empty(points.shape[0], 3), cam.new_empty(points.shape[0])


def _pwf_setup(ctx, inputs, output):
    points, cam, bbox_height, center, orig_shape, Tz, focal, img_w, img_h = inputs
    ctx.save_for_backward(points, cam, bbox_height, center, orig_shape, Tz)
    ctx.consts = (focal, img_w, img_h)


def _pwf_bwd(ctx, g_kp, g_kpw, g_focal, g_cam_t):
    points, cam, bbox_height, center, orig_shape, Tz = ctx.saved_tensors
    B, N = points.shape[0], points.shape[1]
    z = lambda g, shape: g.contiguous() if g is not None else points.new_zeros(shape)  # noqa: E731
    g_points, g_cam, g_Tz = project_full_backward(points, cam, bbox_height, center, orig_shape, Tz, *ctx.consts,
                                                  z(g_kp, (B, N, 2)), z(g_kpw, (B, N, 2)), z(g_focal, (B,)),
                                                  z(g_cam_t, (B, 3)))
    return g_points, g_cam, None, None, None, g_Tz, None, None, None


project_weak_full_op.register_autograd(_pwf_bwd, setup_context=_pwf_setup)


# ----------------------------------------------------------------------------------------------
# utils/geometry.py:310-341 perspective_projection as a differentiable op (the drop-in of whmr_b200.geometry): the
# reference's training graph back-propagates through it into the joints (stage != 1), the translation (pred_cam_t -> Tz)
# and the predicted focal length (-> Tz), models/whmr.py:147-173, core/trainer.py:518.
# ----------------------------------------------------------------------------------------------
@torch.library.custom_op("whmr::perspective_projection", mutates_args=(), device_types="cuda")
def perspective_projection_op(points: torch.Tensor, rotation: torch.Tensor, translation: torch.Tensor, focal: torch.Tensor,
                              camera_center: torch.Tensor, retain_z: bool) -> torch.Tensor:
    """rotation: [0] (none), [1,3,3] or [B,3,3]; translation: [0] (none) or [B,3]; focal: [1] or [B]."""
    return perspective_projection(points, rotation if rotation.numel() else None,
                                  translation if translation.numel() else None,
                                  focal if focal.numel() > 1 else float(focal), camera_center, retain_z)


@perspective_projection_op.register_fake
def _(points, rotation, translation, focal, camera_center, retain_z):
    return points.new_empty(points.shape[0], points.shape[1], 3 if retain_z else 2)


@torch.library.custom_op("whmr::perspective_projection_backward", mutates_args=(), device_types="cuda")
def perspective_projection_backward(points: torch.Tensor, rotation: torch.Tensor, translation: torch.Tensor,
                                    focal: torch.Tensor, g_out: torch.Tensor, retain_z: bool) -> \
        tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    points, g_out = _req(points, "points"), _req(g_out, "grad_out")
    B, N = points.shape[0], points.shape[1]
    dev = points.device
    rot = _req(rotation, "rotation") if rotation.numel() else None
    tr = _req(translation, "translation") if translation.numel() else None
    f_dev = _req(focal.reshape(-1), "focal_length") if focal.numel() > 1 else None
    g_points = torch.empty_like(points)
    g_tr = torch.empty(B, 3, dtype=torch.float32, device=dev)
    g_f = torch.empty(B, dtype=torch.float32, device=dev)
    g_c = torch.empty(B, 2, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().whmr_perspective_projection_backward(
            _p(points), _p(rot), 0 if rot is None else (rot.shape[0] if rot.dim() == 3 else 1), _p(tr), _p(f_dev),
            0.0 if f_dev is not None else float(focal), _p(g_out), B, N, int(bool(retain_z)), _p(g_points), _p(g_tr),
            _p(g_f), _p(g_c), _stream()))
    return g_points, g_tr, g_f, g_c


@perspective_projection_backward.register_fake
def _(points, rotation, translation, focal, g_out, retain_z):
    B = points.shape[0]
    return torch.empty_like(points), points.new_empty(B, 3), points.new_empty(B), points.new_empty(B, 2)


def _pp_setup(ctx, inputs, output):
    points, rotation, translation, focal, camera_center, retain_z = inputs
    ctx.save_for_backward(points, rotation, translation, focal)
    ctx.retain_z = retain_z
    ctx.center_shape = tuple(camera_center.shape)


def _pp_bwd(ctx, g):
    points, rotation, translation, focal = ctx.saved_tensors
    if rotation.numel() and ctx.needs_input_grad[1]:
        raise NotImplementedError("whmr::perspective_projection: no gradient w.r.t. the rotation (the reference passes an "
                                  "identity, models/whmr.py:157-163)")
    g_points, g_tr, g_f, g_c = perspective_projection_backward(points, rotation, translation, focal, g.contiguous(),
                                                               ctx.retain_z)
    g_focal = None
    if ctx.needs_input_grad[3]:
        g_focal = g_f if focal.numel() > 1 else g_f.sum().reshape(focal.shape)
    return (g_points, None, g_tr if translation.numel() else None, g_focal,
            g_c.reshape(ctx.center_shape) if ctx.needs_input_grad[4] else None, None)


perspective_projection_op.register_autograd(_pp_bwd, setup_context=_pp_setup)
