"""Target for compute-sanitizer (not a pytest): one small pass through every kernel family.
    compute-sanitizer --tool memcheck python tests/gpu_sanitizer_target.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "chainer-faster-rcnn_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import frcnn_oracle as orc  # noqa: E402
from frcnn_b200.engine import Engine, StreamRunner  # noqa: E402
from frcnn_b200.resnet_engine import ResNetEngine  # noqa: E402
from frcnn_b200.train_engine import RcnnTrainer, RpnTrainer  # noqa: E402

anchors = orc.generate_anchors(ratios=(0.5, 1, 2), scales=(8, 16, 32))
params = orc.make_params(seed=1234)
H, W = 75, 101                                            # ragged everything
x = torch.from_numpy(orc.make_image(H, W, seed=0)[0]).cuda()
eng = Engine(params, precision="bf16x3", anchors=anchors, use_graph=False, with_detect=True, det_conf=0.05)
prob, boxes, plan = eng(x)
print("forward", prob.shape)
raw = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (60, 81, 3), dtype=np.uint8)).pin_memory()
from frcnn_b200 import preprocess  # noqa: E402
s, Hh, Ww = preprocess.plan_size(60, 81)
plan2 = eng.plan(Hh, Ww)
print("stream", StreamRunner(plan2, src_hw=(60, 81), depth=2).run([raw, raw, raw]))
gt = torch.tensor([[5, 8, 60, 70, 3], [30, 10, 95, 60, 7]], dtype=torch.float32).cuda()
tr = RpnTrainer(params, H, W, anchors, subsample="device")
print("rpn step", tr.step(x, gt).cpu().numpy())
rc = RcnnTrainer(params, H, W, anchors, post_n=50)
np.random.seed(0)
try:
    print("rcnn step", rc.step(x, gt).cpu().numpy())
except Exception as e:                                    # tiny image: the sampler may keep nothing
    print("rcnn step skipped:", str(e)[:80])
rp = orc.make_resnet_params(50, seed=1)
re_ = ResNetEngine(rp, 50, precision="bf16x3", anchors=anchors, use_graph=False, post_n=50)
p2, b2, _ = re_(torch.from_numpy(orc.make_image(96, 131, seed=2)[0]).cuda())
print("resnet", p2.shape)
# round 2 additions: the host-array front end (pinned chunked upload, one result block), the host NMS paths
# (one-kernel zero-copy path and the chip-wide pipeline), frcnn_linear with ragged shapes
from frcnn_b200 import ops  # noqa: E402
res, _ = eng.call_host(orc.make_image(H, W, seed=3)[0])
print("host call", res["count"], res["prob"].shape)
for n in (1, 65, 300, 2049):
    d = np.random.default_rng(n).uniform(0, 200, size=(n, 5)).astype(np.float32)
    d[:, 2:4] += d[:, 0:2]
    print("cpu_nms_host", n, len(ops.cpu_nms_host(d, 0.5)))
xa = torch.randn((1, 77, 192), device="cuda")
hi = xa.to(torch.bfloat16)
act = ops.Act(hi, (xa - hi.float()).to(torch.bfloat16))
wh, wl = ops.pack_conv_weights(torch.randn((96, 192), device="cuda") * 0.05)
y, y32 = ops.linear(act, wh, wl, ops.pad_bias(torch.zeros(96, device="cuda"), 96), True, ld_f32=96)
print("linear", y32.shape)
torch.cuda.synchronize()
print("SANITIZER_TARGET_DONE")
