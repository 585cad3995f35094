#!/usr/bin/env python
"""Golden SMPL bodies produced by EXECUTING the reference's own in-tree SMPL code.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_golden_smpl.py

What runs, unmodified, from the reference tree (models/smpl_webuser/ -- the original SMPL loader the
batched smplx/pare wrapper is a twin of):
  * serialization.ready_arguments  (:78-111)  v_shaped = shapedirs.dot(betas) + v_template,
                                              J = J_regressor . v_shaped,
                                              v_posed = v_shaped + posedirs.dot(lrotmin(pose))
  * posemapper.lrotmin             (:36-39)   numpy branch: cv2.Rodrigues(p) - I for joints 1..23
  * verts.verts_core -> lbs.verts_core -> lbs.global_rigid_transformation (lbs.py:27-79) with xp = numpy:
                                              chain, rest-pose removal, T = A.dot(weights.T), skinning, Jtr
`chumpy` (an autodiff array library, environment.yml:49) is not installed.  The functions above
take the array module as the `xp` argument and have explicit numpy branches, so they never need
chumpy's arithmetic; the files only `import chumpy` at module level.  A stand-in module that forwards
`array`/`vstack` to numpy and `MatVecMult(m, v)` to `m.dot(v)` satisfies those imports.  `models/__init__.py`
imports `pare`, so `models` is registered as a bare namespace pointing at the reference directory and
only `models.smpl_webuser.*` is loaded.

The model is the seeded synthetic SMPL-shaped model of whmr_b200.synthetic (seed 0, 'random'
weights; the SMPL weights are licence-gated and absent); its checksum is stored so a consumer can tell
if its generator drifted.  Inputs: the 8 real pose/shape rows of the reference's vendored fixtures
(real_pose_shape.npz), theta = 0, and seeded random poses up to |theta| ~ pi.
Output: smpl_webuser_outputs.npz  (pose [n,72], betas [n,10], verts [n,6890,3], Jtr [n,24,3]; computed
in float64 by the reference code, stored as float32 -- 6e-8 m rounding against a 1e-5 m tolerance).
"""
import hashlib
import os
import sys
import types

import numpy as np

REF = os.environ.get('WHMR_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def install_shims():
    ch = types.ModuleType('chumpy')
    ch.array = np.asarray
    ch.zeros = np.zeros
    ch.vstack = np.vstack

    class Ch(object):   # base class named at import time by posemapper.Rodrigues; never instantiated
        pass
    ch.Ch = Ch
    chch = types.ModuleType('chumpy.ch')
    chch.MatVecMult = lambda mtx, vec: mtx.dot(vec)
    ch.ch = chch
    sys.modules['chumpy'] = ch
    sys.modules['chumpy.ch'] = chch
    models = types.ModuleType('models')
    models.__path__ = [os.path.join(REF, 'models')]
    sys.modules['models'] = models


def numpy_xp():
    """The `xp` array module handed to the reference's functions: numpy itself, except that `concatenate`
    promotes Python scalars to 1-element arrays the way chumpy's does (lbs.py:56 concatenates `(J[i,:], 0)`,
    which plain numpy rejects).  Being a different object from the `chumpy` stand-in, it selects the
    `cv2.Rodrigues` branch at lbs.py:36-38."""
    xp = types.ModuleType('numpy_xp')
    for name in ('vstack', 'hstack', 'dstack', 'array', 'zeros'):
        setattr(xp, name, getattr(np, name))
    xp.concatenate = lambda seq, *a, **k: np.concatenate([np.atleast_1d(x) for x in seq], *a, **k)
    return xp


def model_checksum(m):
    h = hashlib.sha256()
    for k in ('v_template', 'shapedirs', 'posedirs', 'weights', 'J_regressor', 'parents'):
        h.update(np.ascontiguousarray(m[k]).tobytes())
    return h.hexdigest()


def main():
    sys.path.insert(0, ROOT)
    install_shims()
    from models.smpl_webuser import serialization, verts   # the reference's files
    import whmr_b200.synthetic as syn

    m = syn.make_smpl_model(seed=0, weights='random')
    V = m['v_template'].shape[0]
    f64 = lambda a: np.asarray(a, dtype=np.float64)  # noqa: E731
    base = {   # the pickle schema serialization.ready_arguments expects
        'v_template': f64(m['v_template']),
        'shapedirs': f64(m['shapedirs']),
        'posedirs': f64(m['posedirs']).T.reshape(V, 3, -1),   # smplx stores reshape(-1,207).T of this
        'weights': f64(m['weights']),
        'J_regressor': f64(m['J_regressor']),
        'kintree_table': m['kintree_table'],
        'J': f64(m['J_regressor']).dot(f64(m['v_template'])),   # the pickle's rest joints; recomputed from betas (:104-107)
        'bs_type': 'lrotmin', 'bs_style': 'lbs',
    }
    rng = np.random.default_rng(7)
    real = np.load(os.path.join(HERE, 'real_pose_shape.npz'))
    pose = [f64(real['pose_aa'])[:8], np.zeros((1, 72))]
    betas = [f64(real['betas'])[:8], np.zeros((1, 10))]
    n_rand = 7
    p = rng.normal(0, 0.3, size=(n_rand, 72)); p[:, :3] = rng.normal(0, 1.0, size=(n_rand, 3))
    p[0, 3:6] = (np.pi - 1e-3, 0, 0)          # a joint at ~180 degrees
    p[1, 6:9] = (0, 0, 1e-6)                  # and one at ~0
    pose.append(p); betas.append(np.clip(rng.normal(0, 1, size=(n_rand, 10)), -3, 3))
    pose = np.concatenate(pose); betas = np.concatenate(betas)

    xp = numpy_xp()
    verts_out, jtr_out = [], []
    for i in range(pose.shape[0]):
        dd = dict(base)
        dd['pose'] = pose[i].copy(); dd['betas'] = betas[i].copy()
        dd = serialization.ready_arguments(dd)
        v, Jtr = verts.verts_core(pose=dd['pose'], v=dd['v_posed'], J=dd['J'], weights=dd['weights'],
                                  kintree_table=dd['kintree_table'], bs_style=dd['bs_style'], want_Jtr=True, xp=xp)
        verts_out.append(v); jtr_out.append(Jtr)
    out = dict(pose=pose.astype(np.float32), betas=betas.astype(np.float32),
               verts=np.stack(verts_out).astype(np.float32), Jtr=np.stack(jtr_out).astype(np.float32),
               model_sha256=np.array(model_checksum(m)), model_seed=np.array(0))
    path = os.path.join(HERE, 'smpl_webuser_outputs.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, {k: getattr(v, 'shape', None) for k, v in out.items()}, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
