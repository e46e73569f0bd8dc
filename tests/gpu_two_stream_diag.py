"""Diagnostic (not a pytest): throughput of K independent images in flight on K streams (one ForwardPlan each,
shared weights) versus one stream.  Usage: python tests/gpu_two_stream_diag.py [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "chainer-faster-rcnn_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import frcnn_oracle as orc  # noqa: E402
from frcnn_b200.engine import Engine, ForwardPlan  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
anchors = orc.generate_anchors(ratios=(0.5, 1, 2), scales=(8, 16, 32))
params = orc.make_params(seed=1234)
for prec in ("bf16x3",):
    eng = Engine(params, precision=prec, anchors=anchors, use_graph=True)
    imgs = [torch.from_numpy(orc.make_image(600, 1000, seed=i)[0]).cuda() for i in range(4)]
    from frcnn_b200 import ops
    for nstream, cap in ((1, 0), (2, 0), (3, 0), (4, 0), (3, 74), (4, 74)):
        ops.set_conv_max_ctas(cap)
        plans = [ForwardPlan(eng.weights, 600, 1000, anchors=anchors) for _ in range(nstream)]
        streams = [torch.cuda.Stream() for _ in range(nstream)]
        for p in plans:
            p.forward(imgs[0])
        torch.cuda.synchronize()

        def run(n):
            for i in range(n):
                k = i % nstream
                with torch.cuda.stream(streams[k]):
                    plans[k].forward(imgs[i % 4])
        run(6)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams:
            s.wait_event(e0)
        run(steps)
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("%s streams=%d max_ctas=%d: %.3f ms/img  %.1f img/s" % (prec, nstream, cap, ms / steps, 1e3 * steps / ms), flush=True)
        del plans
    ops.set_conv_max_ctas(0)
