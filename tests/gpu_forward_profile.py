"""Diagnostic (not a pytest): per-kernel durations of one 600x1000 forward (eager launches, warm) from the torch profiler."""
import os
os.environ.setdefault("FRCNN_PDL", "0")          # per-kernel durations: a PDL kernel's duration would include its wait for the previous one
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "chainer-faster-rcnn_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import frcnn_oracle as orc  # noqa: E402
from frcnn_b200.engine import Engine  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

anchors = orc.generate_anchors(ratios=(0.5, 1, 2), scales=(8, 16, 32))
eng = Engine(orc.make_params(seed=1234), precision="bf16x3", anchors=anchors, use_graph=False, with_detect=True)
plan = eng.plan(600, 1000)
x = torch.from_numpy(orc.make_image(600, 1000, seed=0)[0]).cuda()
for _ in range(5):
    plan.forward(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(10):
        plan.forward(x)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 10.0, e.count / 10) for e in prof.key_averages()]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print("sum of kernel time per image: %.1f us" % tot)
for k, t, c in rows[:24]:
    print("%-70s %8.1f us  x%.0f" % (k[:70], t, c))
