"""Diagnostic (not a pytest): time one train_rpn.py step (config #5 shape: 600x1000) by phase."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "chainer-faster-rcnn_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import frcnn_oracle as orc  # noqa: E402
from frcnn_b200.train_engine import RpnTrainer  # noqa: E402

H, W = 600, 1000
anchors = orc.generate_anchors(ratios=(0.5, 1, 2), scales=(8, 16, 32))
params = orc.make_params(seed=1234)
x = torch.from_numpy(orc.make_image(H, W, seed=0)[0]).cuda()
gt = torch.tensor([[100, 120, 400, 380, 3], [500, 200, 900, 560, 7], [50, 50, 200, 180, 1], [600, 30, 780, 150, 5]], dtype=torch.float32).cuda()
for prec in ("bf16x3", "bf16"):
    tr = RpnTrainer(params, H, W, anchors, precision=prec, subsample="device")
    for _ in range(3):
        tr.step(x, gt)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    n = 10
    tf = tb = tu = 0.0
    for _ in range(n):
        ev[0].record()
        losses = tr.forward(x, gt)
        ev[1].record()
        tr.backward()
        ev[2].record()
        tr.update()
        ev[3].record()
        torch.cuda.synchronize()
        tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2]); tu += ev[2].elapsed_time(ev[3])
    print("%s: forward %.3f ms  backward %.3f ms  update+repack %.3f ms  total %.3f ms  (%.1f steps/s)  losses %s  mem %.1f GB" % (
        prec, tf / n, tb / n, tu / n, (tf + tb + tu) / n, 1e3 * n / (tf + tb + tu), tr.last_losses.cpu().numpy(),
        torch.cuda.max_memory_allocated() / 2**30), flush=True)
    if prec == "bf16x3":
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            tr.step(x, gt)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
    del tr
    torch.cuda.empty_cache()
