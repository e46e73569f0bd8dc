"""Diagnostic (not a pytest file): where does the e2e (host in / host out) step time go?"""
import sys, time
import numpy as np, torch
sys.path[:0] = ["chainer-faster-rcnn_b200", "oracle"]
import frcnn_oracle as orc
from frcnn_b200.engine import Engine
eng = Engine(orc.make_params(seed=1234), anchors=orc.generate_anchors(ratios=(0.5, 1, 2), scales=(8, 16, 32)))
plan = eng.plan(600, 1000)
host = torch.from_numpy(orc.make_image(600, 1000, seed=0)[0]).pin_memory()
plan.forward(host.cuda()); torch.cuda.synchronize()
res = torch.empty((300, 84), dtype=torch.float32).pin_memory()
def t(fn, n=20):
    torch.cuda.synchronize(); s = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - s) / n * 1e3
print("h2d 7.2MB pinned      %.3f ms" % t(lambda: plan.x_in.copy_(host, non_blocking=True)))
print("graph replay          %.3f ms" % t(lambda: plan.graph.replay()))
print("d2h boxes + sync      %.3f ms" % t(lambda: (res.copy_(plan.boxes, non_blocking=True), torch.cuda.synchronize())))
def step():
    plan.x_in.copy_(host, non_blocking=True); plan.graph.replay(); res.copy_(plan.boxes, non_blocking=True); torch.cuda.synchronize()
print("full e2e step         %.3f ms" % t(step))
def step_sync_each():
    plan.x_in.copy_(host, non_blocking=True); torch.cuda.synchronize(); a = time.perf_counter()
    plan.graph.replay(); torch.cuda.synchronize(); b = time.perf_counter()
    return b - a
print("replay after sync     %.3f ms" % (np.mean([step_sync_each() for _ in range(10)]) * 1e3))
