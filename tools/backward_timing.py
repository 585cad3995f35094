#!/usr/bin/env python
"""Forward + backward time of the differentiable SMPL / BodyModelHead ops at training batch sizes (eager, CUDA events)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whmr_b200.synthetic as syn  # noqa: E402
from whmr_b200 import _lib  # noqa: E402
from whmr_b200.regressor import BodyModelHead  # noqa: E402
from whmr_b200.smpl import SMPL  # noqa: E402

dev = torch.device("cuda:0")
model = syn.make_smpl_model(seed=0)
smpl = SMPL(model=model).to(dev)
head = BodyModelHead(smpl, model['Dmap0'], model['Dmap1'], model['ssm'], model['J_regressor_h36m'])
head.train_stage = 1


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1000.0


for B in [int(x) for x in (sys.argv[1:] or ["64", "256", "1024"])]:
    b = syn.make_bodies(B, seed=5)
    T = lambda a, g=False: torch.from_numpy(a).to(dev).requires_grad_(g)  # noqa: E731
    rm, be, cam, tz = T(b['rotmat'], True), T(b['betas'], True), T(b['cam'], True), T(b['Tz'], True)
    args = (T(b['bbox_height']), T(b['center']), T(b['orig_shape']))

    def fwd():
        return head(rm, be, cam, args[0], args[1], args[2], tz, J_regressor=True)

    def fwd_bwd():
        o = fwd()
        (o['verts'].sum() + o['kp_3d'].pow(2).sum() + o['kp_2d'].sum() + o['kp_2d_w'].sum() + o['smpl_kp_3d'].sum()).backward()
        rm.grad = be.grad = cam.grad = tz.grad = None

    with torch.no_grad():
        t_f = timeit(fwd)
    t_fb = timeit(fwd_bwd)
    _lib.lib().whmr_launch_count_reset()
    fwd_bwd()
    torch.cuda.synchronize()
    print("B=%d: BodyModelHead forward %.1f us (eager) | forward+loss+backward %.1f us | %d whmr kernel launches"
          % (B, t_f, t_fb, _lib.launch_count()))
