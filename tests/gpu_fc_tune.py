"""Diagnostic (not a pytest): fc6 / fc7 / conv5 GEMM time vs (cta_group, BN)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "chainer-faster-rcnn_b200"))
import torch  # noqa: E402
from frcnn_b200 import ops  # noqa: E402


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        torch.cuda._sleep(200000)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


for (name, H, W, Cin, Cout, k) in [("fc6", 1, 300, 25088, 4096, 1), ("fc7", 1, 300, 4096, 4096, 1), ("conv5", 38, 63, 512, 512, 3),
                                   ("conv4_2", 75, 125, 512, 512, 3)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((H, W, Cin), device="cuda", generator=g)
    xa = ops.Act(x.to(torch.bfloat16), (x - x.to(torch.bfloat16).float()).to(torch.bfloat16))
    w = torch.randn((Cout, Cin, k, k), device="cuda", generator=g) * 0.01
    hi, lo = ops.pack_conv_weights(w, cin_pad=Cin)
    b = ops.pad_bias(torch.zeros(Cout, device="cuda"), Cout)
    out = None
    for cg in (1, 2):
        for bn in (64, 128, 256):
            ops.set_conv_cta_group(cg)
            ops.set_conv_tile(bn, 0, 0)
            try:
                y, _ = ops.conv2d(xa, hi, lo, b, k, True)
                t = timeit(lambda: ops.conv2d(xa, hi, lo, b, k, True, out=y))
                print("%-8s cg=%d bn=%3d: %.1f us" % (name, cg, bn, 1e3 * t), flush=True)
            except Exception as e:
                print("%-8s cg=%d bn=%3d: %s" % (name, cg, bn, str(e)[:80]), flush=True)
    ops.set_conv_cta_group(0)
    ops.set_conv_tile(0, 0, 0)
    y, _ = ops.conv2d(xa, hi, lo, b, k, True)
    print("%-8s auto: %.1f us" % (name, 1e3 * timeit(lambda: ops.conv2d(xa, hi, lo, b, k, True, out=y))), flush=True)
