"""The C-ABI library: loads on a CPU-only box, exports exactly the symbols include/whmr_b200.h
declares, reports errors through return codes (no compute happens without a GPU), and the product
path refuses CPU tensors instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "whmr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(whmr_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from whmr_b200 import _lib
    return _lib


def test_header_and_binding_agree(lib):
    names = _header_functions()
    assert len(names) >= 25
    assert names == sorted(lib.SIGNATURES), (set(names) ^ set(lib.SIGNATURES))


def test_library_exports_every_declared_symbol(lib):
    dll = ctypes.CDLL(lib.LIB_PATH)
    for n in _header_functions():
        assert hasattr(dll, n), n
    assert dll.whmr_abi_version() == 1


def test_library_is_self_contained(lib):
    import subprocess
    out = subprocess.run(["ldd", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libcuda.so" not in out and "libcudart" not in out   # static cudart, no torch


def test_argument_errors_are_return_codes(lib):
    L = lib.lib()
    assert L.whmr_project_weak(None, None, 4, 4, 1000.0, 256.0, 256.0, None, None) == 1      # WHMR_E_INVALID
    assert b"null pointer" in L.whmr_last_error()
    assert L.whmr_project_weak(None, None, 0, 4, 1000.0, 256.0, 256.0, None, None) == 0      # empty batch is fine
    assert L.whmr_sample_bilinear(None, 7, 1, 1, 4, 4, None, 0, 1, None, None) == 1          # bad layout
    assert L.whmr_smpl_workspace_bytes(None, 8) == 0
    with pytest.raises(lib.WhmrError):
        lib.check(L.whmr_joint_errors(None, None, 3, 1000, None, None, None))


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_gpu_means_loud_failure_not_fallback(lib, smpl_model):
    from whmr_b200 import ops
    from whmr_b200.smpl import SMPL
    m = SMPL(model=smpl_model)                         # constructing on CPU is fine (DataLoader workers do it)
    assert m.v_template.shape == (6890, 3) and m.faces.shape == (13776, 3)
    with pytest.raises(lib.WhmrError):
        m(betas=torch.zeros(2, 10), body_pose=torch.zeros(2, 69), global_orient=torch.zeros(2, 3))
    with pytest.raises(lib.WhmrError):
        ops.sample_bilinear(torch.zeros(1, 4, 8, 8), torch.zeros(1, 3, 2))
    with pytest.raises(lib.WhmrError):                 # cudaMalloc fails -> WHMR_E_CUDA, surfaced as an exception
        ops.SmplHandle(smpl_model['v_template'], smpl_model['shapedirs'], smpl_model['posedirs'],
                       smpl_model['J_regressor'], smpl_model['weights'], smpl_model['parents'], "cuda:0")


def test_smpl_module_state_dict_names(smpl_model):
    """buffer / parameter names of smplx.SMPL + the PARE wrapper, so reference checkpoints load strict."""
    from whmr_b200.smpl import SMPL
    keys = set(SMPL(model=smpl_model, batch_size=2, create_transl=True).state_dict())
    for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "parents", "lbs_weights", "faces_tensor",
              "J_regressor_extra", "vertex_joint_selector.extra_joints_idxs", "betas", "global_orient",
              "body_pose", "transl"):
        assert k in keys, k


def test_maf_extractor_state_dict_names():
    from whmr_b200.maf_extractor import MAF_Extractor
    keys = set(MAF_Extractor(mesh_downsampling=None).state_dict())
    assert keys == {"conv0.weight", "conv0.bias", "conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "Dmap"}


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "w-hmr_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in src.replace("no oracle", ""), f


def test_joint_map_matches_reference_table():
    from whmr_b200 import constants as c
    assert len(c.JOINT_MAP_49) == 49 and c.JOINT_MAP_49[8] == 0 and c.JOINT_MAP_49[0] == 24
    assert c.H36M_TO_J14 == (6, 5, 4, 1, 2, 3, 16, 15, 14, 11, 12, 13, 8, 10)
    assert len(c.vertex_joint_selector_ids()) == 21
    assert np.array_equal(np.asarray(c.SMPL_PARENTS)[1:] < np.arange(1, 24), np.ones(23, bool))


def test_modules_deepcopy_and_pickle_without_native_handles():
    """copy.deepcopy (EMA copies) and pickle (spawn-start DataLoader workers build an SMPL per dataset,
    datasets/base_dataset.py:145) carry the drop-in modules; native handles are left behind and rebuilt lazily."""
    import copy
    import pickle

    import numpy as np
    import torch
    import whmr_b200.synthetic as syn
    from whmr_b200.maf_extractor import MAF_Extractor
    from whmr_b200.regressor import BodyModelHead
    from whmr_b200.smpl import SMPL
    model = syn.make_smpl_model(seed=0)
    smpl = SMPL(model=model)
    smpl._dev['cuda:9'] = (object(), object())          # stands in for a live (SmplHandle, Readout) pair
    ext = MAF_Extractor(mesh_downsampling=None)
    ext._mlp = object()
    head = BodyModelHead(smpl, model['Dmap0'], model['Dmap1'], model['ssm'], model['J_regressor_h36m'])
    head._ro[('cuda:9', True)] = object()
    for clone in (copy.deepcopy, lambda m: pickle.loads(pickle.dumps(m))):
        s2, e2, h2 = clone(smpl), clone(ext), clone(head)
        assert s2._dev == {} and e2._mlp is None and h2._ro == {} and h2.smpl._dev == {}
        assert torch.equal(s2.v_template, smpl.v_template) and torch.equal(s2.joint_map, smpl.joint_map)
        assert all(torch.equal(a, b) for a, b in zip(e2.state_dict().values(), ext.state_dict().values()))
        assert np.array_equal(h2._ssm, head._ssm) and h2.fuse_projection == head.fuse_projection
    assert len(smpl._dev) == 1 and ext._mlp is not None and len(head._ro) == 1     # the originals keep theirs
