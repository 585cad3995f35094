"""One BodyModelHead forward + loss + backward at B (default 64) for `ncu --metrics gpu__time_duration.sum`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whmr_b200.synthetic as syn  # noqa: E402
from whmr_b200.loop import RegressorLoop  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
model = syn.make_smpl_model(seed=0)
loop = RegressorLoop(model, dev)
head = loop.head
head.train_stage = 2
b = syn.make_bodies(B, seed=5)
T = lambda a, g=False: torch.from_numpy(a).to(dev).requires_grad_(g)  # noqa: E731
rm, be, cam, tz = T(b["rotmat"], True), T(b["betas"], True), T(b["cam"], True), T(b["Tz"], True)
bb = (T(b["bbox_height"]), T(b["center"]), T(b["orig_shape"]))
for it in range(3):
    o = head(rm, be, cam, bb[0], bb[1], bb[2], tz, J_regressor=True, is_train=True)
    (o["verts"].pow(2).sum() + o["kp_3d"].pow(2).sum() + o["kp_2d"].pow(2).sum() + o["kp_2d_w"].pow(2).sum() +
     o["smpl_kp_3d"].pow(2).sum()).backward()
    torch.cuda.synchronize()
    if it == 1:
        torch.cuda.profiler.start()
torch.cuda.profiler.stop()
