"""Per-body float64 NumPy restatement of the ORIGINAL SMPL math that IS in the reference
tree (models/smpl_webuser/) -- TEST INFRASTRUCTURE.

Independent second derivation used to cross-pin `smpl_oracle.py` (whose source, smplx, is
absent).  chumpy is not installed, so the `xp == numpy` branches are followed:
  * rotation: cv2.Rodrigues            models/smpl_webuser/lbs.py:36-38, posemapper.py:36-39
  * v_shaped = shapedirs.dot(betas)+v_template         serialization.py:102, verts.py:42-45
  * J = J_regressor . v_shaped                         serialization.py:104-107
  * v_posed = v_shaped + posedirs.dot(lrotmin(pose))   serialization.py:108, posemapper.py:36-39
  * chain / rest-pose removal / skinning               lbs.py:27-79
Input model arrays use the smplx in-memory layout (posedirs [207, V*3]) and are re-viewed
as the pickle layout ([V,3,207]) here.
"""
import numpy as np


def _rodrigues(r):
    import cv2
    return cv2.Rodrigues(np.asarray(r, dtype=np.float64).reshape(3, 1))[0]


def lrotmin(pose):
    """posemapper.py:36-39: concat over joints 1.. of (Rodrigues(p) - I).ravel()."""
    p = np.asarray(pose, dtype=np.float64).ravel()[3:]
    return np.concatenate([(_rodrigues(pp) - np.eye(3)).ravel() for pp in p.reshape(-1, 3)]).ravel()


def global_rigid_transformation(pose, J, parents):
    """lbs.py:27-60 with xp = numpy."""
    pose = np.asarray(pose, dtype=np.float64).reshape(-1, 3)
    nj = pose.shape[0]
    with_zeros = lambda x: np.vstack((x, np.array([[0.0, 0.0, 0.0, 1.0]])))  # noqa: E731
    results = {0: with_zeros(np.hstack((_rodrigues(pose[0]), J[0].reshape(3, 1))))}
    for i in range(1, nj):
        p = int(parents[i])
        results[i] = results[p].dot(with_zeros(np.hstack((
            _rodrigues(pose[i]), (J[i] - J[p]).reshape(3, 1)))))
    pack = lambda x: np.hstack([np.zeros((4, 3)), x.reshape(4, 1)])  # noqa: E731
    results_global = [results[i] for i in range(nj)]
    results2 = [results_global[i] - pack(results_global[i].dot(np.concatenate((J[i], [0.0]))))
                for i in range(nj)]
    return np.dstack(results2), results_global


def smpl_body(model, pose_aa, betas):
    """One body: returns (verts [V,3], posed joints Jtr [24,3]) in float64.
    verts.py:32-90 / serialization.py:78-137 + lbs.py:63-79 (verts_core)."""
    f64 = lambda a: np.asarray(a, dtype=np.float64)  # noqa: E731
    v_template = f64(model['v_template'])
    V = v_template.shape[0]
    shapedirs = f64(model['shapedirs'])
    posedirs = f64(model['posedirs']).T.reshape(V, 3, -1)   # pickle layout [V,3,207]
    weights = f64(model['weights'])
    Jreg = f64(model['J_regressor'])
    parents = np.asarray(model['parents'])
    betas = f64(betas)
    v_shaped = shapedirs.dot(betas) + v_template
    J = Jreg.dot(v_shaped)
    v_posed = v_shaped + posedirs.dot(lrotmin(pose_aa))
    A, A_global = global_rigid_transformation(pose_aa, J, parents)
    T = A.dot(weights.T)                                   # [4,4,V]
    rest_shape_h = np.vstack((v_posed.T, np.ones((1, V))))
    v = (T[:, 0, :] * rest_shape_h[0, :].reshape(1, -1) +
         T[:, 1, :] * rest_shape_h[1, :].reshape(1, -1) +
         T[:, 2, :] * rest_shape_h[2, :].reshape(1, -1) +
         T[:, 3, :] * rest_shape_h[3, :].reshape(1, -1)).T
    Jtr = np.vstack([g[:3, 3] for g in A_global])
    return v[:, :3], Jtr
