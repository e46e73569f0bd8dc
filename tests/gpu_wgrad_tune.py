"""Diagnostic (not a pytest): weight-gradient split-K GEMM time vs BN / splits for the VGG layer shapes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "chainer-faster-rcnn_b200"))
import torch  # noqa: E402
from frcnn_b200 import ops, train_ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(1)


def tb(planes, c, h, w):
    t = train_ops.TBuf(planes, c, h, w, "cuda")
    t.hi.copy_(torch.randn(t.hi.shape, device="cuda", generator=g).to(torch.bfloat16))
    t.lo.copy_((torch.randn(t.lo.shape, device="cuda", generator=g) * 0.004).to(torch.bfloat16))
    return t


def timeit(fn, n=5):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for (name, H, W, cin, cout) in [("conv1_2", 600, 1000, 64, 64), ("conv2_2", 300, 500, 128, 128), ("conv3_2", 150, 250, 256, 256),
                                ("conv4_2", 75, 125, 512, 512), ("conv5_2", 38, 63, 512, 512)]:
    a, b = tb(1, cout, H, W), tb(3, cin, H, W)
    kb = a.Kp // 64
    for bn in (64, 128, 256):
        if bn > max(64, cin):
            continue
        for mult in (1, 2, 4):
            tiles = 9 * ((cout + 127) // 128) * ((cin + bn - 1) // bn)
            s = max(1, min((mult * 148 + tiles - 1) // tiles, kb // 8))
            ops.set_conv_tile(bn, 0, 0)
            t = timeit(lambda: train_ops.gemm_nt_splitk(a.hi[0], a.lo[0], b.hi, b.lo, groups=9, row_stride=a.Wp, splits=s))
            gf = 2.0 * H * W * cin * cout * 9 / 1e9
            print("%-8s bn=%3d splits=%3d: %.1f us  (%.0f TF alg)" % (name, bn, s, 1e3 * t, gf / t), flush=True)
    ops.set_conv_tile(0, 0, 0)
    del a, b
