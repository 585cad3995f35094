#!/usr/bin/env python
"""Where the B=256 loop step goes: graph replays of the full step and of the step with kernel classes removed
(the removed op returns tensors cached from a full run, so everything downstream still has valid inputs)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whmr_b200.synthetic as syn  # noqa: E402
from whmr_b200 import ops  # noqa: E402
from whmr_b200.loop import RegressorLoop, make_loop_inputs  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
model = syn.make_smpl_model(seed=0)
loop = RegressorLoop(model, dev)
feats, params, bbox = make_loop_inputs(B, dev)


def timeit(fn, n=300):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1000.0


def graph_time():
    g, _ = loop.capture(feats, params, bbox)
    return timeit(g.replay)


orig = {k: getattr(ops, k) for k in ('project_weak_full_op', 'project_weak_op', 'sample_bilinear_op', 'project_sample_op',
                                     'smpl_lbs_readout')}
cache = {}


def cached(name):
    fn = orig[name]

    def record(*a):
        r = fn(*a)
        cache.setdefault(name, []).append(r)
        return r
    return record


# one eager run that records every op's outputs in call order
for k in orig:
    setattr(ops, k, cached(k))
loop.step(feats, params, bbox)
torch.cuda.synchronize()
for k in orig:
    setattr(ops, k, orig[k])


def replay_of(name):
    it = {'i': 0}

    def f(*a):
        r = cache[name][it['i'] % len(cache[name])]
        it['i'] += 1
        return r
    return f


t_full = graph_time()
print("full step: %.1f us" % t_full)
for label, names in (("without projections (4 launches)", ['project_weak_full_op', 'project_weak_op']),
                     ("without sampling (3 launches)", ['sample_bilinear_op', 'project_sample_op']),
                     ("without SMPL+read-outs (15 launches)", ['smpl_lbs_readout']),
                     ("sampling only", ['smpl_lbs_readout', 'project_weak_full_op', 'project_weak_op']),
                     ("SMPL+read-outs only", ['sample_bilinear_op', 'project_sample_op', 'project_weak_full_op', 'project_weak_op'])):
    for k in names:
        setattr(ops, k, replay_of(k))
    t = graph_time()
    for k in names:
        setattr(ops, k, orig[k])
    print("%-40s %.1f us   (delta %.1f)" % (label, t, t_full - t))
