"""The body-model part of the evaluation pass (evaluate/eval.py:150-228, BASELINE config 5): ground-truth SMPL
from axis-angle pose, predicted SMPL from rotation matrices (or predicted vertices handed in), H36M 17-joint
regression -> 14 joints -> pelvis-centred (:198-219), MPJPE (:222), PA-MPJPE (utils/pose_utils.py:10-75, on the
device instead of a per-frame NumPy SVD loop on the host) and PVE (:208-209).

Frames are independent, so a pass is sharded contiguously over ranks (`dist.shard_bounds`), each rank walks its
shard in chunks (82.7 KB of vertices per frame bounds the chunk), and the only collective is the gather of the
three per-frame error vectors at the end (`dist.all_gather_rows`; 35,515 x 12 B for 3DPW).
"""
import numpy as np
import torch

from . import constants, dist as wdist, ops
from .smpl import SMPL


class EvalPass:
    def __init__(self, smpl, J_regressor_h36m, joint_mapper=None, smpl_male=None, smpl_female=None):
        """smpl: whmr_b200.smpl.SMPL (shared with the model, core/trainer.py:66); J_regressor_h36m [17,V]
        (data/J_regressor_h36m.npy); joint_mapper: H36M_TO_J14 (default) or H36M_TO_J17 (mpi-inf-3dhp, :150).
        smpl_male / smpl_female: the gendered models of the 3DPW protocol (core/trainer.py:58-64, 783-791): with them and a
        per-frame `gender` vector the ground truth comes from the male model, overwritten by the female one where
        gender == 1, exactly as the reference selects it."""
        assert isinstance(smpl, SMPL)
        self.smpl = smpl
        self.smpl_male, self.smpl_female = smpl_male, smpl_female
        self._J = np.asarray(J_regressor_h36m, dtype=np.float64)
        self._map = list(constants.H36M_TO_J14 if joint_mapper is None else joint_mapper)
        self._ro = {}

    def _readout(self, device):
        ro = self._ro.get(str(device))
        if ro is None:
            import scipy.sparse as sp
            V, J = self.smpl.v_template.shape[0], self.smpl.J_regressor.shape[0]
            h = sp.hstack([sp.csr_matrix(self._J), sp.csr_matrix((self._J.shape[0], J))]).tocsr()
            n17, nk = h.shape[0], len(self._map)
            sub = np.full(n17 + nk, -1, dtype=np.int32)
            sub[n17:] = 0                                   # minus the pelvis = regressed joint 0 (:201-203, :214-216)
            ro = ops.Readout([('j17', h), ('kp', h[self._map])], V, J, device, sub_rows=sub)
            self._ro[str(device)] = ro
        return ro

    def joints(self, betas, pose, pose_is_rotmat, model=None):
        """SMPL forward + pelvis-centred evaluation joints in one pass -> (vertices [n,V,3], kp [n,14,3])."""
        h, _ = (model or self.smpl)._state(betas.device)
        ro = self._readout(betas.device)
        verts, _, flat = ops.smpl_lbs_readout(h.id, ro.id, betas, pose, bool(pose_is_rotmat))
        return verts, ro.split(flat, betas.shape[0])['kp']

    def joints_from_vertices(self, verts):
        return self._readout(verts.device).apply(verts)['kp']

    def __call__(self, gt_pose, gt_betas, pred_rotmat=None, pred_betas=None, pred_vertices=None, gender=None):
        """gt_pose [n,72] axis-angle, gt_betas [n,10]; prediction either as (pred_rotmat [n,24,3,3], pred_betas)
        or as pred_vertices [n,V,3] (the model's `global_verts`, :181).  gender [n] (0 male, 1 female): ground truth from
        the gendered models (core/trainer.py:783-791).  -> dict of per-frame errors in metres:
        mpjpe, pa_mpjpe, pve  (the reference multiplies by 1000 when printing, :262-266)."""
        aa = gt_pose.reshape(gt_pose.shape[0], -1)
        if gender is not None:
            if self.smpl_male is None or self.smpl_female is None:
                raise ValueError("EvalPass: gender given but no smpl_male / smpl_female models")
            vm, km = self.joints(gt_betas, aa, False, self.smpl_male)
            vf, kf = self.joints(gt_betas, aa, False, self.smpl_female)
            fem = (gender.to(vm.device) == 1).view(-1, 1, 1)
            gt_verts, gt_kp = torch.where(fem, vf, vm), torch.where(fem, kf, km)    # kp is linear in the vertices
        else:
            gt_verts, gt_kp = self.joints(gt_betas, aa, False)
        if pred_vertices is None:
            pred_vertices, pred_kp = self.joints(pred_betas, pred_rotmat.reshape(pred_rotmat.shape[0], -1, 3, 3), True)
        else:
            pred_kp = self.joints_from_vertices(pred_vertices)
        mp, pa = ops.joint_errors(pred_kp, gt_kp)
        return {'mpjpe': mp, 'pa_mpjpe': pa, 'pve': ops.vertex_errors(pred_vertices, gt_verts)}

    def run_sharded(self, gt_pose, gt_betas, pred_rotmat, pred_betas, chunk=4096, device=None):
        """Whole pass over host (or device) arrays of N frames: this rank evaluates rows shard_bounds(N, rank,
        world) in chunks and every rank returns the gathered [N] error vectors."""
        rank, world = wdist.world_info()
        N = gt_pose.shape[0]
        lo, hi = wdist.shard_bounds(N, rank, world)
        dev = torch.device(device) if device is not None else self.smpl.v_template.device
        T = lambda a: (a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))).to(dev)  # noqa: E731
        parts = {'mpjpe': [], 'pa_mpjpe': [], 'pve': []}
        for a in range(lo, hi, chunk):
            b = min(hi, a + chunk)
            r = self(T(gt_pose[a:b]), T(gt_betas[a:b]), T(pred_rotmat[a:b]), T(pred_betas[a:b]))
            for k in parts:
                parts[k].append(r[k])
        empty = torch.empty(0, dtype=torch.float32, device=dev)
        local = torch.stack([torch.cat(parts[k]) if parts[k] else empty for k in ('mpjpe', 'pa_mpjpe', 'pve')], dim=1)
        full = wdist.all_gather_rows(local, N)
        return {'mpjpe': full[:, 0], 'pa_mpjpe': full[:, 1], 'pve': full[:, 2]}
