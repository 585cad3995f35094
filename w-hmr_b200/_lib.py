"""ctypes binding of libwhmr_b200.so (C ABI declared in include/whmr_b200.h).

There is no fallback: if the library is missing, or a compute entry point is called without a
CUDA device, the call raises.  `build()` compiles the library in-tree with nvcc for sm_100a.
"""
import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwhmr_b200.so")
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]

_lock = threading.Lock()
_lib = None


class WhmrError(RuntimeError):
    pass


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> w-hmr_b200/libwhmr_b200.so"""
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    hdr = os.path.join(INCLUDE, "whmr_b200.h")
    if not force and os.path.exists(LIB_PATH):
        newest = max(os.path.getmtime(p) for p in srcs + [hdr])
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("WHMR_NVCC_EXTRA", "").split()   # e.g. -DWHMR_FUSED_COLLECTOR (experiments)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", LIB_PATH, os.path.join(CSRC, "whmr_b200.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise WhmrError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr[-8000:]))
    if verbose:
        print(r.stderr)
    return LIB_PATH


class SmplGlue(C.Structure):
    _fields_ = [("gram_schmidt", C.c_int32), ("rotmat_out", C.c_void_p), ("pose_aa_out", C.c_void_p),
                ("theta_out", C.c_void_p), ("cam", C.c_void_p), ("root_pose", C.c_void_p)]


class FinishProjection(C.Structure):     # whmr_finish_projection
    _fields_ = [("row0", C.c_int32), ("n_points", C.c_int32), ("focal", C.c_float), ("img_w", C.c_float),
                ("img_h", C.c_float), ("bbox_height", C.c_void_p), ("center", C.c_void_p), ("orig_shape", C.c_void_p),
                ("Tz", C.c_void_p), ("cam", C.c_void_p * 8), ("full", C.c_int32 * 8), ("kp_weak", C.c_void_p * 8),
                ("kp_norm", C.c_void_p * 8), ("focal_out", C.c_void_p * 8), ("cam_t_out", C.c_void_p * 8)]


class SmplModelDesc(C.Structure):
    _fields_ = [("n_verts", C.c_int32), ("n_joints", C.c_int32), ("n_betas", C.c_int32),
                ("v_template", C.c_void_p), ("shapedirs", C.c_void_p), ("posedirs", C.c_void_p),
                ("J_regressor", C.c_void_p), ("lbs_weights", C.c_void_p), ("parents", C.c_void_p)]


_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes).  Must list every symbol include/whmr_b200.h declares
# (tests/test_abi_cpu.py parses the header and checks both directions).
SIGNATURES = {
    "whmr_abi_version": (C.c_int, []),
    "whmr_last_error": (C.c_char_p, []),
    "whmr_launch_count": (C.c_uint64, []),
    "whmr_launch_count_reset": (None, []),
    "whmr_debug_set_trap_buffer": (None, [_vp]),
    "whmr_smpl_create": (C.c_int, [C.POINTER(SmplModelDesc), _i, C.POINTER(_vp)]),
    "whmr_smpl_destroy": (C.c_int, [_vp]),
    "whmr_smpl_set_gemm_mode": (C.c_int, [_vp, _i]),
    "whmr_smpl_get_info": (C.c_int, [_vp] + [C.POINTER(C.c_int32)] * 5),
    "whmr_smpl_workspace_bytes": (_sz, [_vp, _i]),
    "whmr_smpl_forward": (C.c_int, [_vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "whmr_smpl_forward_readout": (C.c_int, [_vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i,
                                            C.POINTER(C.c_int), _vp, _sz, _vp]),
    "whmr_smpl_forward_regressor": (C.c_int, [_vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i,
                                              C.POINTER(C.c_int), _vp, _vp, _sz, _vp]),
    "whmr_readout_workspace_bytes": (_sz, [_vp, _i]),
    "whmr_smpl_chunk_bodies": (C.c_int, [_vp]),
    "whmr_smpl_is_fused": (C.c_int, [_vp]),
    "whmr_readout_finish": (C.c_int, [_vp, _vp, _i, _vp, _vp, _vp]),
    "whmr_readout_finish_multi": (C.c_int, [_vp, _i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _i, _vp]),
    "whmr_readout_finish_project_multi": (C.c_int, [_vp, _i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _i,
                                                    C.POINTER(FinishProjection), _vp]),
    "whmr_smpl_stage_chain": (C.c_int, [_vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "whmr_smpl_stage_pose_blend": (C.c_int, [_vp, _i, _vp, _sz, _vp]),
    "whmr_smpl_stage_skin": (C.c_int, [_vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "whmr_smpl_set_probe_events": (C.c_int, [_vp, _vp, _vp, _vp]),
    "whmr_smpl_reserve": (C.c_int, [_vp, _i]),
    "whmr_smpl_forward_host": (C.c_int, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "whmr_batch_rodrigues": (C.c_int, [_vp, _i, _vp, _vp]),
    "whmr_rot6d_to_rotmat": (C.c_int, [_vp, _i, _vp, _vp]),
    "whmr_unbiased_gram_schmidt": (C.c_int, [_vp, _i, _vp, _vp]),
    "whmr_rotmat_to_axis_angle": (C.c_int, [_vp, _i, _vp, _vp]),
    "whmr_batch_rodrigues_quat": (C.c_int, [_vp, _i, _vp, _vp]),
    "whmr_readout_create": (C.c_int, [_i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, C.POINTER(_vp)]),
    "whmr_readout_destroy": (C.c_int, [_vp]),
    "whmr_readout_apply": (C.c_int, [_vp, _vp, _vp, _i, _vp, _vp]),
    "whmr_project_weak": (C.c_int, [_vp, _vp, _i, _i, _f, _f, _f, _vp, _vp]),
    "whmr_perspective_projection": (C.c_int, [_vp, _vp, _i, _vp, _vp, _f, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "whmr_project_full": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "whmr_project_weak_full": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "whmr_project_crop": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _f, _vp, _vp, _vp]),
    "whmr_sample_bilinear": (C.c_int, [_vp, _i, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp]),
    "whmr_project_sample": (C.c_int, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _f, _f, _f, _vp, _vp, _vp]),
    "whmr_maf_mlp_create": (C.c_int, [_i, _i, _i, _i, C.POINTER(_vp)]),
    "whmr_maf_mlp_destroy": (C.c_int, [_vp]),
    "whmr_maf_mlp_set_weights": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "whmr_sample_reduce": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp, _vp]),
    "whmr_project_sample_reduce": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _f, _f, _f, _vp, _vp, _vp, _vp]),
    "whmr_estimate_translation": (C.c_int, [_vp, _vp, _i, _i, _i, _f, _f, _f, _vp, _vp]),
    "whmr_gather_vertices": (C.c_int, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "whmr_smpl_backward_workspace_bytes": (C.c_size_t, [_vp, _i]),
    "whmr_smpl_backward": (C.c_int, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "whmr_readout_backward": (C.c_int, [_vp, _vp, _i, _vp, _vp, _vp]),
    "whmr_project_weak_backward": (C.c_int, [_vp, _vp, _vp, _i, _i, _f, _f, _f, _vp, _vp, _vp]),
    "whmr_perspective_projection_backward": (C.c_int, [_vp, _vp, _i, _vp, _vp, _f, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "whmr_project_full_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "whmr_sample_bilinear_backward": (C.c_int, [_vp, _i, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp]),
    "whmr_joint_errors": (C.c_int, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "whmr_vertex_errors": (C.c_int, [_vp, _vp, _i, _i, _vp, _vp]),
}


def lib():
    """The loaded library; raises WhmrError (never falls back) when it is not built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise WhmrError(
                        "libwhmr_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; "
                        "g.build()'`; there is no CPU or PyTorch fallback for this path." % LIB_PATH)
                l = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(l, name)
                    fn.restype = res
                    fn.argtypes = args
                if l.whmr_abi_version() != 1:
                    raise WhmrError("ABI version mismatch")
                _lib = l
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().whmr_last_error().decode("utf-8", "replace")
        raise WhmrError("whmr_b200 error %d: %s" % (rc, msg))


def launch_count():
    return int(lib().whmr_launch_count())
