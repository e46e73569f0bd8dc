"""Diagnostic (not a pytest): per-layer conv timing with and without the spin kernel in front, same process."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "chainer-faster-rcnn_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import frcnn_oracle as orc  # noqa: E402
import bench  # noqa: E402
from frcnn_b200.engine import Engine  # noqa: E402

anchors = orc.generate_anchors(ratios=(0.5, 1, 2), scales=(8, 16, 32))
eng = Engine(orc.make_params(seed=1234), precision="bf16x3", anchors=anchors, use_graph=True)
plan = eng.plan(600, 1000)
x = torch.from_numpy(orc.make_image(600, 1000, seed=0)[0]).cuda()
for _ in range(5):
    plan.forward(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30):
    plan.forward(x)
e1.record()
torch.cuda.synchronize()
print("graph replay: %.4f ms/image" % (e0.elapsed_time(e1) / 30))
for spin in (False, True, False, True):
    t = bench.conv_layer_table(plan, torch, reps=5, spin=spin)
    conv = sum(r[1] for r in t if " k3" in r[0] or "->64 k1" in r[0])
    print("spin=%s: conv stack %.4f ms, all GEMMs %.4f ms; conv1_2 %.4f conv4_2 %.4f fc6 %.4f" % (
        spin, conv, sum(r[1] for r in t), t[1][1], t[8][1], t[15][1]))
