"""ResNet-50/101/152 trunk for the forward engine (BASELINE config #4).

The reference's trunk class is /root/reference models/resnet.py:11-45: a thin subclass of the un-vendored
chainer.links.model.vision.resnet.ResNetLayers that returns 'res5' (stride 32, 2048 channels), so the detector is
`FasterRCNN(trunk_class=ResNet, rpn_in_ch=2048, feat_stride=32)` with fc6 reading 2048*7*7 inputs.  Restated structure:

    conv1 7x7/2 p3 (+bias) - bn1 - relu - max_pooling_2d(3, stride=2)      [pad 0, cover_all]
    res2 (64-64-256, stride 1) x3, res3 (128-128-512, /2) x4, res4 (256-256-1024, /2) x23 (101), res5 (512-512-2048, /2) x3
    block 'a'  (BottleneckA): conv1 1x1/s - bn - relu - conv2 3x3 - bn - relu - conv3 1x1 - bn;  shortcut conv4 1x1/s - bn4
    block 'bN' (BottleneckB): same main path, identity shortcut;  out = relu(main + shortcut)

On the device every BatchNormalization (test mode) is folded into the preceding convolution at pack time, every
convolution is the tcgen05 kernel of the VGG path (1x1 = plain GEMM tiles, 3x3 = shared-halo implicit GEMM), conv1 is an
im2col GEMM (K = 147 -> 160), the residual add + ReLU is the epilogue of conv3 (frcnn_conv2d_res), and a stride-2 1x1
convolution is frcnn_subsample2x (shared by conv1 and the projection shortcut) followed by a stride-1 GEMM.
"""
import numpy as np
import torch

from . import ops
from .engine import Engine, ForwardPlan, PackedWeights, _t

RESNET_BLOCKS = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}
RESNET_STAGES = (("res2", 64, 64, 256, 1), ("res3", 256, 128, 512, 2), ("res4", 512, 256, 1024, 2), ("res5", 1024, 512, 2048, 2))
BN_EPS = 2e-5                  # chainer.links.BatchNormalization default
CONV1_KPAD = 160               # 7*7*3 = 147 -> 160 (16-byte rows)


def block_list(n_layers):
    out = []
    for (stage, cin, mid, cout, stride), n in zip(RESNET_STAGES, RESNET_BLOCKS[n_layers]):
        out.append((stage, "a", cin, mid, cout, stride, True))
        for i in range(1, n):
            out.append((stage, "b%d" % i, cout, mid, cout, 1, False))
    return out


def fold_batchnorm(W, gamma, beta, mean, var, bias=None, eps=BN_EPS):
    """Test-mode BN after a convolution == the convolution with W*s and bias beta + (b - mean)*s, s = gamma/sqrt(var+eps).
    float64 host arithmetic (one-time, at weight-pack time), float32 results."""
    s = np.asarray(gamma, np.float64) / np.sqrt(np.asarray(var, np.float64) + eps)
    b0 = np.zeros_like(s) if bias is None else np.asarray(bias, np.float64)
    Wf = (np.asarray(W, np.float64) * s[:, None, None, None]).astype(np.float32)
    bf = (np.asarray(beta, np.float64) + (b0 - np.asarray(mean, np.float64)) * s).astype(np.float32)
    return Wf, bf


class ResNetPackedWeights(PackedWeights):
    def __init__(self, params, n_layers=101, precision="bf16x3", device="cuda", num_classes=21, n_anchors=9):
        self.n_layers = n_layers
        super(ResNetPackedWeights, self).__init__(params, precision, device, num_classes, n_anchors, pool_chw=(2048, 7, 7))

    def _folded(self, params, conv, bn, bias=None):
        return fold_batchnorm(params[conv + "/W"], params[bn + "/gamma"], params[bn + "/beta"], params[bn + "/avg_mean"],
                              params[bn + "/avg_var"], bias)

    def _pack_trunk(self, params, P):
        dev = self.device
        Wf, bf = self._folded(params, "trunk/conv1", "trunk/bn1", params.get("trunk/conv1/b"))
        hi, lo = ops.pack_conv_weights_im2col_general(_t(Wf, dev), CONV1_KPAD, precision=self.precision)
        self.convs["conv1"] = (hi, lo, ops.pad_bias(_t(bf, dev), 64))
        for stage, blk, cin, mid, cout, stride, proj in block_list(self.n_layers):
            base = "trunk/%s/%s" % (stage, blk)
            for ci in ((1, 2, 3, 4) if proj else (1, 2, 3)):
                Wf, bf = self._folded(params, base + "/conv%d" % ci, base + "/bn%d" % ci)
                self.convs["%s/%s/conv%d" % (stage, blk, ci)] = self._pack(_t(Wf, dev), _t(bf, dev), cin_pad=Wf.shape[1])


class ResNetForwardPlan(ForwardPlan):
    """ForwardPlan with the ResNet trunk (the RPN / ProposalLayer / RoI pool / head part is inherited unchanged)."""

    def _alloc_trunk(self, H, W):
        act = self._act
        Ho, Wo = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
        self.x_col = act(Ho, Wo, CONV1_KPAD)
        self.c1 = act(Ho, Wo, 64)
        h, w = (Ho - 2) // 2 + 1, (Wo - 2) // 2 + 1
        self.p1 = act(h, w, 64)
        self.blocks = []
        x = self.p1
        for stage, blk, cin, mid, cout, stride, proj in block_list(self.w.n_layers):
            B = dict(name="%s/%s" % (stage, blk), x=x, stride=stride, proj=proj)
            if stride == 2:
                h, w = (h + 1) // 2, (w + 1) // 2
                B["xs"] = act(h, w, cin)                 # the stride-2 1x1 convs read pixels (2h, 2w)
            B["y1"], B["y2"] = act(h, w, mid), act(h, w, mid)
            if proj:
                B["sc"] = act(h, w, cout)
            B["out"] = act(h, w, cout)
            self.blocks.append(B)
            x = B["out"]
        self.acts = [self.x_col, x]                      # acts[-1] is the feature map ('res5')
        return h, w

    def _run_trunk(self):
        cv = self.w.convs
        n = 0
        ops.pack_image_im2col_general(self.x_in, 7, 2, 3, CONV1_KPAD, out=self.x_col)
        hi, lo, b = cv["conv1"]
        ops.conv2d(self.x_col, hi, lo, b, 1, True, out=self.c1)
        ops.maxpool3x3s2_ceil(self.c1, out=self.p1)
        n += 3
        for B in self.blocks:
            x = B["x"]
            if B["stride"] == 2:
                ops.subsample2x(x, out=B["xs"])
                x = B["xs"]
                n += 1
            hi, lo, b = cv[B["name"] + "/conv1"]
            ops.conv2d(x, hi, lo, b, 1, True, out=B["y1"])
            hi, lo, b = cv[B["name"] + "/conv2"]
            ops.conv2d(B["y1"], hi, lo, b, 3, True, out=B["y2"])
            if B["proj"]:
                hi, lo, b = cv[B["name"] + "/conv4"]
                ops.conv2d(x, hi, lo, b, 1, False, out=B["sc"])
                res = B["sc"]
                n += 1
            else:
                res = B["x"]
            hi, lo, b = cv[B["name"] + "/conv3"]
            ops.conv2d_res(B["y2"], hi, lo, b, 1, True, res, out=B["out"])
            n += 3
        return self.blocks[-1]["out"], n


class ResNetEngine(Engine):
    """`Engine` for FasterRCNN(trunk_class=ResNet, rpn_in_ch=2048, feat_stride=32)."""
    supports_hwc_input = False         # the 7x7 first layer's im2col kernel reads dense (C,H,W) only

    def __init__(self, params, n_layers=101, precision="bf16x3", device="cuda", anchors=None, num_classes=21, n_anchors=9,
                 feat_stride=32, **plan_kwargs):
        self.weights = ResNetPackedWeights(params, n_layers, precision, device, num_classes, n_anchors)
        self.anchors, self.feat_stride = anchors, feat_stride
        self.plan_kwargs = plan_kwargs
        self.plans = {}
        import threading
        self._tls, self._lock = threading.local(), threading.Lock()

    def plan(self, H, W, **overrides):
        kw = dict(self.plan_kwargs)
        kw.update(overrides)
        key = (H, W, tuple(sorted(kw.items())))
        if key not in self.plans:
            self.plans[key] = ResNetForwardPlan(self.weights, H, W, anchors=self.anchors, feat_stride=self.feat_stride, **kw)
        return self.plans[key]


class TrunkOnly(object):
    """The trunk alone (models.resnet.ResNet.__call__): packs only trunk/* and runs ResNetForwardPlan's trunk hooks."""

    class _W(ResNetPackedWeights):
        def __init__(self, params, n_layers, precision, device):        # no RPN / head parameters here
            self.n_layers, self.precision, self.device = n_layers, precision, device
            self.convs = {}
            self._pack_trunk(params, None)

    class _Plan(ResNetForwardPlan):
        def __init__(self, weights, H, W):
            self.w, self.H, self.W = weights, H, W
            dev = weights.device
            x3 = weights.precision == "bf16x3"

            def act(h, w, c):
                hi = torch.empty((h, w, c), dtype=torch.bfloat16, device=dev)
                return ops.Act(hi, torch.empty_like(hi) if x3 else None)
            self._act = act
            self.x_in = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
            self._alloc_trunk(H, W)

    def __init__(self, params, n_layers, precision, device):
        self.weights = self._W(params, n_layers, precision, device)
        self.plans = {}

    def run(self, x_chw):
        _, H, W = x_chw.shape
        if (H, W) not in self.plans:
            self.plans[(H, W)] = self._Plan(self.weights, H, W)
        p = self.plans[(H, W)]
        p.x_in.copy_(x_chw)
        feat, _ = p._run_trunk()
        return feat
