"""The body-model part of `Regressor.forward` / `Regressor.forward_init` (models/whmr.py:128-209 and
:225-269) -- everything between the regressor MLP's outputs (pred_rotmat, pred_shape, pred_cam) and
the 17-key result dict -- as one module over the sm_100a kernels.

Per call the reference issues: unbiased_gram_schmidt (eval mode, ~25 launches), 1 SMPL forward (~100 launches),
`projection`, the predicted-focal `perspective_projection`, rotation_matrix_to_angle_axis (~40 launches), the `theta`
concatenation, the H36M matmul, two dense down-sampling matmuls (75 MFLOP/body), a marker gather, a dense
vertices2joints and a VertexJointSelector.  Here that is 4 launches: chain (with the rotation glue folded in),
fused pose-blend + skinning (+ read-out emits), read-out finishing pass, weak + full projection.
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
        self.fuse_projection = True   # complete_all: the joint projections ride in the read-out finishing launch
        # optional torch.cuda.Stream: the read-out finishing pass and the joint projections of a call run on it,
        # concurrently with whatever the caller enqueues next on the main stream (the feature sampling only needs
        # the markers, which the skinning kernel itself writes).  The caller joins it (RegressorLoop.step does).
        self.side_stream = None
        # Gradient routing of the reference's training graph (models/whmr.py:142-165): cfg.TRAIN.STAGE == 1 -> `projection`
        # sees the joints, the predicted-focal block sees joints.detach(); any other stage the opposite; pred_cam is
        # always detached inside the predicted-focal block (s and pred_cam_t).  None = the reference's configured default
        # (configs/pymaf_config.yaml:26, TRAIN.STAGE: 2) whenever a gradient is required; without autograd both
        # projections run as one fused launch (no routing to do).
        self.train_stage = None
        self.default_train_stage = 2

    def __getstate__(self):
        """copy.deepcopy / pickle: without the native read-out tables, streams and probes (rebuilt / reset on first use)."""
        d = self.__dict__.copy()
        d['_ro'] = {}
        d['probe'] = None
        d['side_stream'] = None
        return d

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
                  scale, pose=None, theta=None, projected=None):
        """projections of the 49 joints + the result dict of Regressor.forward / forward_init (models/whmr.py:142-208)"""
        B = rot.shape[0]
        pred_joints = r['joints']
        needs_grad = torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad
                                                     for t in (pred_joints, pred_cam, Tz))
        if projected is not None:     # deferred schedule: computed inside the finishing launch (complete_all)
            kp_2d, kp_w, focal, cam_t = projected
        elif bbox_height is not None and needs_grad:
            f, w, hgt = constants.FOCAL_LENGTH, float(constants.IMG_RES_WIDTH), float(constants.IMG_RES_HEIGHT)
            st1 = (self.default_train_stage if self.train_stage is None else self.train_stage) == 1
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
        if pose is not None:      # models/whmr.py:174,190 (and :237,253 in forward_init)
            out.update(pose=pose, theta=theta)
        if bbox_height is not None:
            out.update(kp_2d_w=kp_w, focal_length=focal, pred_cam_t=cam_t, scale=scale)
        return out

    # -- deferred schedule (RegressorLoop): inside WHMR.forward's loop (models/whmr.py:550-651) the next iteration reads
    #    only the markers / camera / shape / pose of a Regressor result; the regressor-row read-outs and the joint
    #    projections are read by the caller after the loop.  `begin` runs the SMPL kernels (vertices, one-hot read-outs
    #    such as the markers are complete on return), `complete_all` finishes every pending call in one launch.
    def begin(self, pred_rotmat, pred_shape, J_regressor=None, pred_cam=None, orthonormalize=False):
        """orthonormalize: unbiased_gram_schmidt of the predicted rotations first (Regressor.forward in eval mode,
        models/whmr.py:129-130); pred_cam: the head of `theta` (zeros if None)."""
        if torch.is_tensor(J_regressor):
            if self._h36m is None:
                self.set_h36m_regressor(J_regressor)
            J_regressor = True
        B = pred_rotmat.shape[0]
        dev = pred_rotmat.device
        h, _ = self.smpl._state(dev)
        rot = pred_rotmat.reshape(B, -1, 3, 3)
        ro = self._readout(dev, bool(J_regressor))
        cam = pred_cam if pred_cam is not None else rot.new_zeros(B, 3)
        verts, joints24, flat, scratch, rot_used, pose, theta = ops.smpl_regressor(h.id, ro.id, pred_shape, rot, cam,
                                                                                   bool(orthonormalize), True)
        r = ro.split(flat, B)
        return {'ro': ro, 'verts': verts, 'joints24': joints24, 'flat': flat, 'scratch': scratch, 'r': r, 'rot': rot_used,
                'pred_rotmat': pred_rotmat, 'pred_shape': pred_shape, 'J': bool(J_regressor), 'markers': r['markers'],
                'pose': pose, 'theta': theta}

    def complete_all(self, states, cams, bbox_height=None, center=None, orig_shape=None, Tz=None, full=None, scale=None):
        """states: list from `begin`; cams: pred_cam per state; full[i]: evaluate the predicted-focal block for state i
        (Regressor.forward) or only the weak projection (forward_init); None entries in cams: no projection (the global
        SMPL call, models/whmr.py:641-651).  -> list of result dicts."""
        use_full = [cams[i] is not None and (bool(full[i]) if full is not None else bbox_height is not None)
                    for i in range(len(states))]
        pend = [i for i, s in enumerate(states) if s['scratch'].numel()]
        by_ro = {}
        for i in pend:
            by_ro.setdefault(states[i]['ro'].id, []).append(i)
        projected = {}
        for rid, idxs in by_ro.items():
            for k in range(0, len(idxs), 8):
                g = idxs[k:k + 8]
                # the joint projections ride in the finishing launch (models/whmr.py:142-173, 237): no launch of their own
                proj = None
                if self.fuse_projection and any(cams[i] is not None for i in g):
                    proj = dict(group='joints', cams=[cams[i] for i in g], full=[use_full[i] for i in g],
                                bbox_height=bbox_height, center=center, orig_shape=orig_shape, Tz=Tz,
                                focal=constants.FOCAL_LENGTH, img_w=float(constants.IMG_RES_WIDTH),
                                img_h=float(constants.IMG_RES_HEIGHT))
                res = ops.readout_finish_multi(rid, [states[i]['joints24'] for i in g], [states[i]['flat'] for i in g],
                                               [states[i]['scratch'] for i in g], proj)
                for i, r_ in zip(g, res):
                    if r_ is not None:
                        projected[i] = r_
        outs = []
        for i, s in enumerate(states):
            if cams[i] is None:
                outs.append({'verts': s['verts'], 'r': s['r'], 'pose': s['pose'], 'rotmat': s['rot']})
                continue
            outs.append(self._assemble(s['r'], s['verts'], s['rot'], s['pred_rotmat'], s['pred_shape'], cams[i],
                                       bbox_height if use_full[i] else None, center, orig_shape, Tz, s['J'], scale,
                                       pose=s['pose'], theta=s['theta'], projected=projected.get(i)))
        return outs

    def forward(self, pred_rotmat, pred_shape, pred_cam, bbox_height=None, center=None, orig_shape=None,
                Tz=None, J_regressor=None, scale=None, is_train=False, orthonormalize=None):
        """pred_rotmat [B,24,3,3]; pred_shape [B,10]; pred_cam [B,3].  With bbox_height/center/
        orig_shape/Tz the predicted-focal block (:147-173) is evaluated too (Regressor.forward);
        without them only the weak projection (forward_init).  J_regressor: None, True (use the
        stored H36M regressor) or a tensor [17,V] / [B,17,V].
        is_train: Regressor.forward's flag -- in eval mode the predicted rotations go through
        unbiased_gram_schmidt first (:129-130); forward_init never does.  `orthonormalize` overrides.
        The dict carries every tensor of the reference's (incl. `pose`, `theta`; `rotmat` = the rotations used)."""
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
        if orthonormalize is None:
            orthonormalize = bbox_height is not None and not is_train
        verts, joints24, flat, scratch, rot, pose, theta = ops.smpl_regressor(
            h.id, ro.id, pred_shape, rot, pred_cam, bool(orthonormalize), side is not None)
        if side is not None:
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
                             J_regressor, scale, pose=pose, theta=theta)
        if side is not None:
            # tensors the side stream still reads: under stream capture record_stream() is unavailable, so their blocks must
            # not return to the capture pool before the caller joins the side stream -- the result keeps them alive
            out['_side_keep'] = (joints24, flat, scratch)
            if not torch.cuda.is_current_stream_capturing():
                for t in out.values():
                    if torch.is_tensor(t) and t.is_cuda:
                        t.record_stream(main)
            torch.cuda.set_stream(main)
        return out
