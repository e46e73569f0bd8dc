"""Target for ncu captures (not a pytest): python tests/gpu_ncu_target.py <conv|prep|wgrad|conv1|fc6|forward>."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "chainer-faster-rcnn_b200"))
import torch  # noqa: E402
from frcnn_b200 import ops, train_ops  # noqa: E402

what = sys.argv[1]
g = torch.Generator(device="cuda").manual_seed(1)


def act(h, w, c):
    x = torch.randn((h, w, c), device="cuda", generator=g)
    hi = x.to(torch.bfloat16)
    return ops.Act(hi, (x - hi.float()).to(torch.bfloat16))


if what == "conv":                    # conv4_2: 75x125x512 -> 512, 3x3, halo + CTA pairs
    x = act(75, 125, 512)
    w = torch.randn((512, 512, 3, 3), device="cuda", generator=g) * 0.01
    hi, lo = ops.pack_conv_weights(w, cin_pad=512)
    b = ops.pad_bias(torch.zeros(512, device="cuda"), 512)
    y, _ = ops.conv2d(x, hi, lo, b, 3, True)
    for _ in range(4):
        ops.conv2d(x, hi, lo, b, 3, True, out=y)
elif what == "prep":                  # the conv1_2-level gradient re-layout: pooled routing + mask + transposed copy
    H, W, C = 600, 1000, 64
    yv = act(H, W, C)
    yv.hi.clamp_(min=0)
    p = ops.maxpool2x2_ceil(yv)
    gsrc = act(300, 500, C)
    out = act(H, W, C)
    tb = train_ops.TBuf(1, C, H, W, "cuda")
    tb3 = train_ops.TBuf(3, C, H, W, "cuda")
    for _ in range(4):
        train_ops.grad_prepare(H, W, C, g=gsrc, y=yv, p=p, out=out, tbuf=tb)
        train_ops.grad_prepare(H, W, C, g=yv, tbuf=tb3)
elif what == "wgrad":                 # conv3_2 weight gradient: 256 x 256 x 9 taps over 150x250 pixels
    H, W, C = 150, 250, 256
    a = act(H, W, C)
    tb = train_ops.TBuf(1, C, H, W, "cuda")
    tb3 = train_ops.TBuf(3, C, H, W, "cuda")
    train_ops.grad_prepare(H, W, C, g=a, tbuf=tb)
    train_ops.grad_prepare(H, W, C, g=a, tbuf=tb3)
    for _ in range(4):
        parts = train_ops.gemm_nt_splitk(tb.hi[0], tb.lo[0], tb3.hi, tb3.lo, groups=9, row_stride=tb.Wp, splits=9)
elif what == "conv1":                 # conv1_1 on the compact image: 600x1000x3 -> 64 through the sliding-window tensor map
    H, W = 600, 1000
    x = torch.randn((3, H, W), device="cuda", generator=g) * 60
    w = torch.randn((64, 3, 3, 3), device="cuda", generator=g) * 0.2
    wh, wl = ops.pack_conv_weights_c8(w)
    b = ops.pad_bias(torch.zeros(64, device="cuda"), 64)
    xc8 = ops.pack_image_c8(x)
    y = ops.conv3x3_c8(xc8, H, W, wh, wl, b, True)
    for _ in range(3):
        ops.pack_image_c8(x, out=xc8)
        ops.conv3x3_c8(xc8, H, W, wh, wl, b, True, out=y)
elif what == "fc6":                   # fc6 as frcnn_linear: 300 RoIs x 25088 -> 4096 (swapped operands, N = 160 tiles, 2 K splits)
    R, K, N = 300, 25088, 4096
    x = act(1, R, K)
    w = torch.randn((N, K), device="cuda", generator=g) * 0.01
    hi, lo = ops.pack_conv_weights(w)
    b = ops.pad_bias(torch.zeros(N, device="cuda"), N)
    work = ops.linear_workspace(R, K, N, "cuda")
    y, _ = ops.linear(x, hi, lo, b, True, work=work)
    for _ in range(3):
        ops.linear(x, hi, lo, b, True, out=y, work=work)
elif what == "forward":               # two eager 600x1000 forwards incl. the per-class NMS: the launch list of one image
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import frcnn_oracle as orc
    from frcnn_b200.engine import Engine
    os.environ.setdefault("FRCNN_PDL", "0")
    eng = Engine(orc.make_params(seed=1234), precision="bf16x3", anchors=orc.generate_anchors(ratios=(0.5, 1, 2), scales=(8, 16, 32)),
                 use_graph=False, with_detect=True)
    plan = eng.plan(600, 1000)
    xi = torch.from_numpy(orc.make_image(600, 1000, seed=0)[0]).cuda()
    for _ in range(2):
        plan.forward(xi)
        torch.cuda.synchronize()
    print("launches per image:", plan.n_launches)
torch.cuda.synchronize()
print("done", what)
