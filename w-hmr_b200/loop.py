"""The body-model hot path of `WHMR.forward`'s iterative regressor loop (models/whmr.py:550-651)
with the unchanged PyTorch parts (backbone, deconvs, Tz head, regressor / reduce_dim MLPs) factored
out: per iteration the caller supplies what those MLPs would produce (rotmats, betas, camera) and
gets back what they would consume (sampled point features) plus the Regressor result tensors.

    init      : forward_init        -> SMPL + read-outs + weak projection                 (:550)
    iter 0    : grid sampling of the 7x9 grid on feature level 0, then Regressor.forward  (:596-602)
    iter 1, 2 : MAF_Extractor.forward(markers of the previous output, previous camera) on level i,
                then Regressor.forward                                                     (:604-612)
    global    : 5th SMPL call with the re-estimated global orientation + H36M joints       (:641-651)

`RegressorLoop.step` launches 14 kernels on the deferred schedule (5 x {chain with the rotation glue folded in, fused
pose-blend + skinning}, 3 sampling, ONE finishing pass for the five read-outs with the four joint projections folded into
it) and 22 on the immediate one; `capture()` wraps the step in a CUDA graph so a replay costs one launch from the host.  Every call also produces
the reference's `pose` / `theta` (rotation_matrix_to_angle_axis, :174,190) and, in eval mode, orthonormalises the
predicted rotations first (unbiased_gram_schmidt, :129-130) -- inside the chain kernel, no extra launch.
"""
import numpy as np
import torch

from . import constants, ops
from .regressor import BodyModelHead
from .smpl import SMPL
from .synthetic import grid_points

VITPOSE_LEVELS = ((32, 24), (64, 48), (128, 96))   # models/whmr.py:325-331,543 deconv outputs @256 ch
RES50_LEVELS = ((14, 14), (28, 28), (56, 56))


class RegressorLoop:
    def __init__(self, model, device, backbone='vitpose', gemm_mode=None, with_h36m=True, extractors=None):
        """extractors: optional list of three `MAF_Extractor` modules (models/whmr.py:333-335).  With them the loop also runs
        their `reduce_dim` MLP -- fused into the sampling launch (maf_fused_tc.cuh) -- and returns 'ref_features'
        (3 x [B, 32*N], what `Regressor.forward` consumes, models/whmr.py:597-612) instead of the raw 'point_feats'."""
        self.device = torch.device(device)
        self.extractors = None if extractors is None else [e.to(self.device).eval() for e in extractors]
        self.smpl = SMPL(model=model, gemm_mode=gemm_mode).to(self.device)
        self.head = BodyModelHead(self.smpl, model['Dmap0'], model['Dmap1'], model['ssm'],
                                  model.get('J_regressor_h36m'))
        self.with_h36m = bool(with_h36m) and model.get('J_regressor_h36m') is not None
        self.grid = torch.from_numpy(grid_points(backbone)).to(self.device)   # [63,2] shared by all bodies
        self.backbone = backbone
        self.levels = VITPOSE_LEVELS if backbone == 'vitpose' else RES50_LEVELS
        self.layout = ops.LAYOUT_NCHW
        self._graph = None
        # Optional: run the read-out finishing pass + joint projections of call i on a side stream so they overlap
        # with the sampling / SMPL kernels that follow.  Measured (tools/overlap_test.py, B=256): 0.447 -> 0.604 ms
        # per graph replay -- the persistent 1-CTA/SM tcgen05 kernels wait for SMs still holding side-stream CTAs
        # and lose more in their tails than the overlap gains -- so it is off by default.
        self.overlap = False
        self._side = torch.cuda.Stream(device=self.device)
        # finishing passes of the five read-outs + the joint projections after the loop, in one + four launches
        # (inference only: under autograd the immediate schedule is used); 0.389 -> see profiles/r01_notes.md
        self.defer = True
        self._needs_grad = False
        self.is_train = False   # eval mode: Regressor.forward orthonormalises the predicted rotations (models/whmr.py:129-130)

    def step(self, feats, params, bbox):
        """feats: 3 feature maps [B,256,H_i,W_i]; params: 5 dicts {rotmat [B,24,3,3], betas [B,10],
        cam [B,3]} (init, iter0, iter1, iter2, global); bbox: {bbox_height, center, orig_shape, Tz}.
        Returns the last Regressor output dict + 'point_feats' (3 x [B,256,N]) + global outputs."""
        J = True if self.with_h36m else None
        p = params
        main = torch.cuda.current_stream(self.device)
        needs_grad = torch.is_grad_enabled() and (any(t.requires_grad for q in p for t in q.values()) or
                                                  any(f.requires_grad for f in feats))
        self._needs_grad = needs_grad
        if self.defer and self.head.probe is None and not self.overlap and not needs_grad:
            return self._step_deferred(feats, params, bbox, J)
        self.head.side_stream = self._side if (self.overlap and self.head.probe is None) else None
        out = self.head(p[0]['rotmat'], p[0]['betas'], p[0]['cam'], J_regressor=J)           # forward_init
        point_feats = []
        keep = [out]      # every iteration's result stays alive until the side stream has been joined (see BodyModelHead)
        for it in range(3):
            pf = self._sample(it, feats, out['markers'], p[it]['cam'])
            self.head._mark('sample_l%d' % it)
            point_feats.append(pf)
            q = p[it + 1]
            out = self.head(q['rotmat'], q['betas'], q['cam'], bbox['bbox_height'], bbox['center'],
                            bbox['orig_shape'], bbox['Tz'], J_regressor=J, is_train=self.is_train)   # :598 / :608
            keep.append(out)
        g = p[4]
        h, _ = self.smpl._state(self.device)
        self.head._mark('pre_smpl')
        ro = self.head._readout(self.device, self.with_h36m)
        gverts, gjoints24, flat, _, _, gpose, _ = ops.smpl_regressor(h.id, ro.id, g['betas'], g['rotmat'], g['cam'],
                                                                     False, False)   # :632-644
        r = ro.split(flat, gverts.shape[0])
        self.head._mark('skin_readout')
        if self.head.side_stream is not None:
            main.wait_stream(self._side)     # join: everything in the result dict is complete on the main stream
        res = dict(out)
        res['_keep'] = keep
        res['point_feats' if self.extractors is None else 'ref_features'] = point_feats
        res['global_verts'] = gverts
        res['global_kp_3d'] = r['kp_3d_h36m'] if self.with_h36m else r['joints']            # :646-651
        res['global_pose'] = gpose                                                          # :632-633
        return res

    def _sample(self, it, feats, markers, cam):
        if self.extractors is not None:          # one launch: sampling (+ projection) + reduce_dim MLP
            ext = self.extractors[it]
            ext.layout = self.layout
            # inference: no grad_fn is needed, so the extractor takes its fused kernel; a training step (parameters or
            # inputs of the loop require a gradient) keeps autograd on and gets the sampling op + PyTorch MLP
            with torch.set_grad_enabled(self._needs_grad):
                if it == 0:
                    return ext.sampling(self.grid, im_feat=feats[0])[0]                      # :596-597
                ext.im_feat, ext.cam = feats[it], cam                                        # :564,593
                return ext(markers, None, None, None, None)[0]                               # :606
        if ops.is_host_map(feats[it]):   # pinned host map: gathered in place over PCIe (ops.is_host_map)
            if it == 0:
                return ops.sample_bilinear(feats[0], self.grid, self.layout)
            return ops.project_sample(feats[it], markers.detach(), cam.detach(), constants.FOCAL_LENGTH,
                                      float(constants.IMG_RES_WIDTH), float(constants.IMG_RES_HEIGHT), self.layout)[0]
        if it == 0:
            return ops.sample_bilinear_op(feats[0], self.grid, self.layout)                  # :596-597
        pf, _ = ops.project_sample_op(feats[it], markers, cam, constants.FOCAL_LENGTH,       # :606
                                      float(constants.IMG_RES_WIDTH), float(constants.IMG_RES_HEIGHT),
                                      self.layout)   # projection fused into the sampling launch
        return pf

    def _step_deferred(self, feats, params, bbox, J):
        """Same results, 14 launches instead of 22: the next iteration needs only the markers (written by the SMPL
        kernel itself) and the camera, so the five finishing passes of the read-outs AND the four joint projections run
        as ONE launch after the loop (models/whmr.py:550-651 returns everything at the end as well)."""
        p = params
        states = [self.head.begin(p[0]['rotmat'], p[0]['betas'], J, p[0]['cam'])]             # forward_init
        point_feats = []
        for it in range(3):
            point_feats.append(self._sample(it, feats, states[-1]['markers'], p[it]['cam']))
            q = p[it + 1]
            states.append(self.head.begin(q['rotmat'], q['betas'], J, q['cam'], orthonormalize=not self.is_train))
        # global call, :632-644: the axis-angle of [global_rotmat | last body rotations] IS global_pose
        states.append(self.head.begin(p[4]['rotmat'], p[4]['betas'], J, p[4]['cam']))
        outs = self.head.complete_all(states, [p[0]['cam'], p[1]['cam'], p[2]['cam'], p[3]['cam'], None],
                                      bbox['bbox_height'], bbox['center'], bbox['orig_shape'], bbox['Tz'],
                                      full=[False, True, True, True, False])
        res = dict(outs[3])
        res['point_feats' if self.extractors is None else 'ref_features'] = point_feats
        res['global_verts'] = outs[4]['verts']
        res['global_kp_3d'] = outs[4]['r']['kp_3d_h36m'] if self.with_h36m else outs[4]['r']['joints']
        res['global_pose'] = outs[4]['pose']
        return res

    # ---- CUDA-graph replay ------------------------------------------------------------------------
    def capture(self, feats, params, bbox, warmup=2):
        """Capture `step` on static input tensors; returns (graph, static outputs dict)."""
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.step(feats, params, bbox)
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = self.step(feats, params, bbox)
        return g, out


def make_loop_inputs(B, device, backbone='vitpose', seed=1, rank=0, channels=256, dtype=torch.float32):
    """Synthetic inputs of one loop pass: feature maps ~ N(0,1) generated on the device,
    5 parameter sets, bbox quantities.  Reproducible per (seed, rank)."""
    from . import synthetic as syn
    dev = torch.device(device)
    levels = VITPOSE_LEVELS if backbone == 'vitpose' else RES50_LEVELS
    g = torch.Generator(device=dev).manual_seed(1000 * seed + rank)
    feats = [torch.randn(B, channels, h, w, generator=g, device=dev, dtype=dtype) for h, w in levels]
    params = []
    base = syn.make_bodies(B, seed=seed, rank=rank)
    for i in range(5):
        b = syn.make_bodies(B, seed=seed + 17 * (i + 1), rank=rank, with_real_rows=(i == 0))
        params.append({'rotmat': torch.from_numpy(b['rotmat']).to(dev), 'betas': torch.from_numpy(b['betas']).to(dev),
                       'cam': torch.from_numpy(b['cam']).to(dev)})
    bbox = {k: torch.from_numpy(np.ascontiguousarray(base[k])).to(dev)
            for k in ('bbox_height', 'center', 'orig_shape', 'Tz')}
    return feats, params, bbox
