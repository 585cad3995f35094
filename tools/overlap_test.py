import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whmr_b200.synthetic as syn
from whmr_b200.loop import RegressorLoop, make_loop_inputs
dev = torch.device("cuda:0")
model = syn.make_smpl_model(seed=0)
loop = RegressorLoop(model, dev)
feats, params, bbox = make_loop_inputs(256, dev)
def timeit(fn, n=200):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for ov in (False, True):
    loop.overlap = ov
    loop.step(feats, params, bbox); torch.cuda.synchronize()
    t_eager = timeit(lambda: loop.step(feats, params, bbox), 50)
    g, _ = loop.capture(feats, params, bbox)
    t_graph = timeit(g.replay)
    print("overlap", ov, "eager %.3f ms  graph %.3f ms" % (t_eager, t_graph))
