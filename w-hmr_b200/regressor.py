"""The body-model part of `Regressor.forward` / `Regressor.forward_init` (models/whmr.py:128-209 and
:225-269) -- everything between the regressor MLP's outputs (pred_rotmat, pred_shape, pred_cam) and
the 17-key result dict -- as one module over the sm_100a kernels.

Per call the reference issues: 1 SMPL forward (~100 launches), `projection`, the predicted-focal
`perspective_projection`, the H36M matmul, two dense down-sampling matmuls (75 MFLOP/body), a
marker gather, a dense vertices2joints and a VertexJointSelector.  Here that is 7 launches:
chain, pose-blend, skin, read-out (short rows), read-out (long rows), weak projection, full projection.
"""
import numpy as np
import torch
import torch.nn as nn

from . import constants, ops
from .smpl import SMPL


class BodyModelHead(nn.Module):
    def __init__(self, smpl, Dmap0, Dmap1, ssm, J_regressor_h36m=None):
        """smpl: whmr_b200.smpl.SMPL; Dmap0 [1723,V], Dmap1 [431,1723] (dense or scipy sparse,
        models/whmr.py:77-98); ssm [67] marker vertex ids (:100); J_regressor_h36m [17,V] optional
        (passed per call as `J_regressor` in the reference)."""
        super().__init__()
        assert isinstance(smpl, SMPL)
        self.smpl = smpl
        import scipy.sparse as sp
        self._D0 = sp.csr_matrix(Dmap0, dtype=np.float64)
        self._D10 = (sp.csr_matrix(Dmap1, dtype=np.float64) @ self._D0).tocsr()
        self._ssm = np.asarray(ssm, dtype=np.int64)
        self._h36m = None if J_regressor_h36m is None else np.asarray(J_regressor_h36m, dtype=np.float64)
        self._ro = {}
        self.probe = None    # bench.py: callable(name) recording a CUDA event after each enqueued op
        # optional torch.cuda.Stream: the read-out finishing pass and the joint projections of a call run on it,
        # concurrently with whatever the caller enqueues next on the main stream (the feature sampling only needs
        # the markers, which the skinning kernel itself writes).  The caller joins it (RegressorLoop.step does).
        self.side_stream = None
        # Gradient routing of the reference's training graph (models/whmr.py:142-165): cfg.TRAIN.STAGE == 1 -> `projection`
        # sees the joints, the predicted-focal block sees joints.detach(); any other stage the opposite; pred_cam is
        # always detached inside the predicted-focal block (s and pred_cam_t).  None = no detach (forward-only use:
        # one fused launch for both projections).
        self.train_stage = None

    def _mark(self, name):
        if self.probe is not None:
            self.probe(name)

    def set_h36m_regressor(self, J_regressor):
        J = J_regressor.detach().cpu().numpy() if torch.is_tensor(J_regressor) else np.asarray(J_regressor)
        if J.ndim == 3:          # the reference expands it to [B,17,V] per batch (core/trainer.py:775)
            J = J[0]
        self._h36m = J.astype(np.float64)
        self._ro = {}

    def _readout(self, device, with_h36m):
        key = (str(device), bool(with_h36m))
        ro = self._ro.get(key)
        if ro is None:
            import scipy.sparse as sp
            s = self.smpl
            c = lambda b: b.detach().cpu().numpy()  # noqa: E731
            V, J = s.v_template.shape[0], s.J_regressor.shape[0]
            vid = c(s.vertex_joint_selector.extra_joints_idxs)
            pick = lambda idx, off=0: sp.csr_matrix((np.ones(len(idx)), (np.arange(len(idx)), np.asarray(idx) + off)),  # noqa: E731
                                                    shape=(len(idx), V + J))
            widen = lambda m: sp.hstack([sp.csr_matrix(m), sp.csr_matrix((m.shape[0], J))]).tocsr()  # noqa: E731
            src54 = sp.vstack([pick(np.arange(J), V), pick(vid), widen(c(s.J_regressor_extra).astype(np.float64))]).tocsr()
            groups = [
                ('joints', src54[s.joint_map.numpy()]),                                  # models/smpl.py:74-76
                ('smpl_kp_3d', sp.vstack([widen(c(s.J_regressor).astype(np.float64)), pick(vid)])),  # whmr.py:186-187
                ('sub_verts', widen(self._D0)),                                           # :182
                ('temp_verts', widen(self._D10)),                                         # :183
                ('markers', pick(self._ssm)),                                             # :184
            ]
            sub = None
            if with_h36m:
                if self._h36m is None:
                    raise ValueError("J_regressor requested but no H36M regressor was given")
                n0 = sum(g.shape[0] for _, g in groups)
                h = widen(self._h36m)
                groups.append(('h36m_j17', h))
                groups.append(('kp_3d_h36m', h[list(constants.H36M_TO_J14)]))            # :179
                n_rows = n0 + h.shape[0] + len(constants.H36M_TO_J14)
                sub = np.full(n_rows, -1, dtype=np.int32)
                sub[n0 + h.shape[0]:] = n0                                                # minus pelvis = J17 row 0 (:178,180)
            ro = ops.Readout(groups, V, J, device, sub_rows=sub)
            self._ro[key] = ro
        return ro

    def _assemble(self, r, verts, rot, pred_rotmat, pred_shape, pred_cam, bbox_height, center, orig_shape, Tz, J_regressor,
                  scale):
        """projections of the 49 joints + the result dict of Regressor.forward / forward_init (models/whmr.py:142-208)"""
        B = rot.shape[0]
        pred_joints = r['joints']
        if bbox_height is not None and self.train_stage is not None and torch.is_grad_enabled():
            f, w, hgt = constants.FOCAL_LENGTH, float(constants.IMG_RES_WIDTH), float(constants.IMG_RES_HEIGHT)
            st1 = self.train_stage == 1
            kp_2d = ops.project_weak_op(pred_joints if st1 else pred_joints.detach(), pred_cam, f, w, hgt)
            _, kp_w, focal, cam_t = ops.project_weak_full_op(pred_joints.detach() if st1 else pred_joints, pred_cam.detach(),
                                                             bbox_height, center, orig_shape, Tz, f, w, hgt)
            self._mark('project_weak_full')
        elif bbox_height is not None:   # Regressor.forward: weak + predicted-focal projection, one launch
            kp_2d, kp_w, focal, cam_t = ops.project_weak_full_op(
                pred_joints, pred_cam, bbox_height, center, orig_shape, Tz, constants.FOCAL_LENGTH,
                float(constants.IMG_RES_WIDTH), float(constants.IMG_RES_HEIGHT))
            self._mark('project_weak_full')
        else:                         # forward_init: weak projection only
            kp_2d = ops.project_weak_op(pred_joints, pred_cam, constants.FOCAL_LENGTH,
                                        float(constants.IMG_RES_WIDTH), float(constants.IMG_RES_HEIGHT))
            self._mark('project_weak')
        out = {
            'verts': verts, 'sub_verts': r['sub_verts'], 'temp_verts': r['temp_verts'], 'kp_2d': kp_2d,
            'kp_3d': r['kp_3d_h36m'] if J_regressor else pred_joints,
            'smpl_kp_3d': r['smpl_kp_3d'], 'rotmat': rot, 'pred_cam': pred_cam, 'pred_shape': pred_shape,
            'pred_pose': pred_rotmat.reshape(B, -1), 'pelvis': r['smpl_kp_3d'][:, :1, :], 'markers': r['markers'],
            'joints49': pred_joints,
        }
        if bbox_height is not None:
            out.update(kp_2d_w=kp_w, focal_length=focal, pred_cam_t=cam_t, scale=scale)
        return out

    # -- deferred schedule (RegressorLoop): inside WHMR.forward's loop (models/whmr.py:550-651) the next iteration reads
    #    only the markers / camera / shape / pose of a Regressor result; the regressor-row read-outs and the joint
    #    projections are read by the caller after the loop.  `begin` runs the SMPL kernels (vertices, one-hot read-outs
    #    such as the markers are complete on return), `complete_all` finishes every pending call in one launch.
    def begin(self, pred_rotmat, pred_shape, J_regressor=None):
        if torch.is_tensor(J_regressor):
            if self._h36m is None:
                self.set_h36m_regressor(J_regressor)
            J_regressor = True
        B = pred_rotmat.shape[0]
        dev = pred_rotmat.device
        h, _ = self.smpl._state(dev)
        rot = pred_rotmat.reshape(B, -1, 3, 3)
        ro = self._readout(dev, bool(J_regressor))
        verts, joints24, flat, scratch = ops.smpl_lbs_readout_deferred(h.id, ro.id, pred_shape, rot, True)
        r = ro.split(flat, B)
        return {'ro': ro, 'verts': verts, 'joints24': joints24, 'flat': flat, 'scratch': scratch, 'r': r, 'rot': rot,
                'pred_rotmat': pred_rotmat, 'pred_shape': pred_shape, 'J': bool(J_regressor), 'markers': r['markers']}

    def complete_all(self, states, cams, bbox_height=None, center=None, orig_shape=None, Tz=None, full=None, scale=None):
        """states: list from `begin`; cams: pred_cam per state; full[i]: evaluate the predicted-focal block for state i
        (Regressor.forward) or only the weak projection (forward_init); None entries in cams: no projection (the global
        SMPL call, models/whmr.py:641-651).  -> list of result dicts."""
        pend = [s for s in states if s['scratch'].numel()]
        by_ro = {}
        for s in pend:
            by_ro.setdefault(s['ro'].id, []).append(s)
        for rid, group in by_ro.items():
            for i in range(0, len(group), 8):
                g = group[i:i + 8]
                ops.readout_finish_multi(rid, [s['joints24'] for s in g], [s['flat'] for s in g], [s['scratch'] for s in g])
        outs = []
        for i, s in enumerate(states):
            if cams[i] is None:
                outs.append({'verts': s['verts'], 'r': s['r']})
                continue
            use_full = bool(full[i]) if full is not None else bbox_height is not None
            outs.append(self._assemble(s['r'], s['verts'], s['rot'], s['pred_rotmat'], s['pred_shape'], cams[i],
                                       bbox_height if use_full else None, center, orig_shape, Tz, s['J'], scale))
        return outs

    def forward(self, pred_rotmat, pred_shape, pred_cam, bbox_height=None, center=None, orig_shape=None,
                Tz=None, J_regressor=None, scale=None):
        """pred_rotmat [B,24,3,3]; pred_shape [B,10]; pred_cam [B,3].  With bbox_height/center/
        orig_shape/Tz the predicted-focal block (:147-173) is evaluated too (Regressor.forward);
        without them only the weak projection (forward_init).  J_regressor: None, True (use the
        stored H36M regressor) or a tensor [17,V] / [B,17,V]."""
        if torch.is_tensor(J_regressor):
            if self._h36m is None:
                self.set_h36m_regressor(J_regressor)
            J_regressor = True
        B = pred_rotmat.shape[0]
        dev = pred_rotmat.device
        h, _ = self.smpl._state(dev)
        rot = pred_rotmat.reshape(B, -1, 3, 3)
        self._mark('pre_smpl')
        ro = self._readout(dev, bool(J_regressor))
        side = self.side_stream
        main = torch.cuda.current_stream(dev)
        if side is None:
            verts, joints24, flat = ops.smpl_lbs_readout(h.id, ro.id, pred_shape, rot, True)
        else:
            verts, joints24, flat, scratch = ops.smpl_lbs_readout_deferred(h.id, ro.id, pred_shape, rot, True)
            side.wait_stream(main)
            torch.cuda.set_stream(side)     # finishing pass + projections below go to the side stream
            if scratch.numel():
                ops.readout_finish(ro.id, joints24, flat, scratch)
            if not torch.cuda.is_current_stream_capturing():   # allocator bookkeeping for cross-stream use
                for t in (scratch, flat, joints24, pred_cam):
                    t.record_stream(side)
        r = ro.split(flat, B)
        self._mark('skin_readout')
        out = self._assemble(r, verts, rot, pred_rotmat, pred_shape, pred_cam, bbox_height, center, orig_shape, Tz,
                             J_regressor, scale)
        if side is not None:
            if not torch.cuda.is_current_stream_capturing():
                for t in out.values():
                    if torch.is_tensor(t) and t.is_cuda:
                        t.record_stream(main)
            torch.cuda.set_stream(main)
        return out
