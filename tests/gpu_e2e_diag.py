"""Diagnostic (not a pytest file): why is a single graph replay slow after an H2D copy / idle gap?"""
import sys, time
import numpy as np, torch
sys.path[:0] = ["chainer-faster-rcnn_b200", "oracle"]
import frcnn_oracle as orc
from frcnn_b200.engine import Engine
eng = Engine(orc.make_params(seed=1234), anchors=orc.generate_anchors(ratios=(0.5, 1, 2), scales=(8, 16, 32)))
plan = eng.plan(600, 1000)
host = torch.from_numpy(orc.make_image(600, 1000, seed=0)[0]).pin_memory()
dev_img = host.cuda()
other = torch.empty_like(dev_img)
plan.forward(dev_img); torch.cuda.synchronize()

def timed_replay(label, pre):
    res = []
    for _ in range(8):
        pre()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a = time.perf_counter()
        e0.record(); plan.graph.replay(); e1.record()
        torch.cuda.synchronize()
        b = time.perf_counter()
        res.append(((b - a) * 1e3, e0.elapsed_time(e1)))
    print("%-38s wall %s | gpu %s" % (label, " ".join("%.2f" % r[0] for r in res), " ".join("%.2f" % r[1] for r in res)))

timed_replay("A: nothing before", lambda: None)
timed_replay("B: H2D into x_in", lambda: plan.x_in.copy_(host, non_blocking=True))
timed_replay("C: D2D into x_in", lambda: plan.x_in.copy_(dev_img, non_blocking=True))
timed_replay("D: H2D into other buffer", lambda: other.copy_(host, non_blocking=True))
timed_replay("E: 5 ms host sleep", lambda: time.sleep(0.005))
timed_replay("F: 50 ms host sleep", lambda: time.sleep(0.05))
# eager (no graph) single shot
def eager(label, pre):
    res = []
    for _ in range(5):
        pre(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a = time.perf_counter(); e0.record(); plan._run(); e1.record(); torch.cuda.synchronize(); b = time.perf_counter()
        res.append(((b - a) * 1e3, e0.elapsed_time(e1)))
    print("%-38s wall %s | gpu %s" % (label, " ".join("%.2f" % r[0] for r in res), " ".join("%.2f" % r[1] for r in res)))
eager("G: eager after nothing", lambda: None)
eager("H: eager after H2D", lambda: plan.x_in.copy_(host, non_blocking=True))
# back-to-back
torch.cuda.synchronize(); a = time.perf_counter()
for _ in range(20): plan.graph.replay()
torch.cuda.synchronize(); print("I: 20 back-to-back replays: %.3f ms each" % ((time.perf_counter() - a) / 20 * 1e3))
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,pstate,power.draw", "--format=csv"], capture_output=True, text=True).stdout)
