"""One train_rpn.py step on the device: forward (trunk + RPN) -> AnchorTargetLayer -> rpn_loss -> backward through the
RPN heads, the RPN 3x3 conv and the 13 trunk convs -> [gradient all-reduce] -> WeightDecay + MomentumSGD.

Mirrors /root/reference train_rpn.py:144-179 with `model.rpn_train = True`: FasterRCNN.__call__ takes the RPN training
branch (models/faster_rcnn.py:114-116 -> models/region_proposal_network.py:117-156), the optimizer is
MomentumSGD(lr=0.001) (momentum 0.9 default) + WeightDecay(0.0005) (train_rpn.py:165-167).  Only parameters that receive a
gradient are updated (Chainer v1 `cleargrads`: fc6/fc7/cls_score/bbox_pred have grad None in RPN mode and are skipped).

Every dense contraction runs on the tcgen05 kernel of the forward path:
    forward conv     frcnn_conv2d                       (pools NOT fused here: backward needs the un-pooled maps)
    data gradient    frcnn_conv2d with the 180-degree rotated, in/out-swapped filter (frcnn_pack_conv_weights_dgrad)
    weight gradient  frcnn_gemm_nt_splitk over the zero-padded pixel axis (9 tap groups x K splits), reduced in fixed order
and the element-wise backward ops (ReLU mask, max-pool routing, hi/lo split, transposition) are fused into
frcnn_grad_prepare.  Gradients / master weights / momentum are float32 in three flat buffers (one NCCL all-reduce).
"""
import numpy as np
import torch

from . import ops, shard, train_ops
from ._lib import FrcnnError
from .engine import VGG16_LAYERS

TRUNK = [it for it in VGG16_LAYERS if it != "pool"]              # (name, cin, cout) x 13
POOL_AFTER = {VGG16_LAYERS[i - 1][0] for i, it in enumerate(VGG16_LAYERS) if it == "pool"}


class RpnTrainer(object):
    """Weights, optimizer state and buffers of the RPN-mode training step for one image shape."""

    def __init__(self, params, H, W, anchors, precision="bf16x3", lr=0.001, momentum=0.9, weight_decay=0.0005,
                 feat_stride=16, loss_lambda=1.0, delta=3.0, subsample="device", seed=0, device="cuda", process_group=None):
        if precision not in ops.PRECISIONS:
            raise FrcnnError("precision must be one of %s" % (ops.PRECISIONS,))
        self.x3 = precision == "bf16x3"
        self.precision, self.H, self.W = precision, H, W
        self.lr, self.momentum, self.weight_decay = lr, momentum, weight_decay
        self.feat_stride, self.loss_lambda, self.delta = feat_stride, loss_lambda, delta
        self.subsample, self.seed, self.pg = subsample, seed, process_group
        dev = self.device = torch.device(device)
        self.anchors = torch.from_numpy(np.ascontiguousarray(anchors, dtype=np.float64)).to(dev)
        self.A = int(self.anchors.shape[0])
        # ---- trainable parameters in the reference's naming; one flat fp32 master / momentum / gradient buffer
        self.layers = [dict(name="trunk/" + n, cin=ci, cout=co, k=3) for n, ci, co in TRUNK]
        self.layers.append(dict(name="RPN/rpn_conv_3x3", cin=512, cout=512, k=3))
        heads = [("RPN/rpn_cls_score", 2 * self.A), ("RPN/rpn_bbox_pred", 4 * self.A)]
        names, shapes = [], []
        for L in self.layers:
            names += [L["name"] + "/W", L["name"] + "/b"]
            shapes += [(L["cout"], L["cin"], L["k"], L["k"]), (L["cout"],)]
        # the twin heads are stored adjacently so that (W_cls | W_bbox) is ONE [6A, 512] matrix: cls W, bbox W, cls b, bbox b
        for n, co in heads:
            names.append(n + "/W")
            shapes.append((co, 512, 1, 1))
        for n, co in heads:
            names.append(n + "/b")
            shapes.append((co,))
        sizes = [int(np.prod(s)) for s in shapes]
        offs = [0] + [int(v) for v in np.cumsum(sizes)[:-1]]         # no padding: the twin heads must stay contiguous
        total = int(sum(sizes))
        self.w_flat = torch.zeros((total,), dtype=torch.float32, device=dev)
        self.v_flat = torch.zeros_like(self.w_flat)
        self.g_flat = torch.zeros_like(self.w_flat)
        self.index = {}
        for n, s, o, sz in zip(names, shapes, offs, sizes):
            self.index[n] = (o, sz, s)
            src = params[n]
            self.w_flat[o:o + sz] = torch.from_numpy(np.ascontiguousarray(src, dtype=np.float32).reshape(-1)).to(dev)
        self.n_heads = 6 * self.A
        self.head_ld = ops.round_up(self.n_heads, 32)
        # ---- geometry + activations (forward keeps every conv output; pooled maps where the graph pools)
        x3 = self.x3

        def act(h, w, c):
            hi = torch.empty((h, w, c), dtype=torch.bfloat16, device=dev)
            return ops.Act(hi, torch.empty_like(hi) if x3 else None)

        self.x_in = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
        self.x_col = act(H, W, 32)
        h, w = H, W
        for L in self.layers[:13]:
            L["H"], L["W"] = h, w
            L["y"] = act(h, w, L["cout"])
            short = L["name"].split("/")[1]
            L["pool"] = short in POOL_AFTER
            if L["pool"]:
                h, w = (h + 1) // 2, (w + 1) // 2
                L["p"] = act(h, w, L["cout"])
        self.fh, self.fw = h, w
        R = self.layers[13]
        R["H"], R["W"], R["pool"] = h, w, False
        R["y"] = act(h, w, 512)
        self.head_out = torch.empty((h * w, self.head_ld), dtype=torch.float32, device=dev)
        self.targets = train_ops.AnchorTargets(self.A, h, w, dev)
        self.zero_bias = torch.zeros((1024,), dtype=torch.float32, device=dev)
        # gradient activations (NHWC) and transposed buffers, shared between layers of identical shape
        self._acts, self._tbufs = {}, {}
        self.packed = {}
        self.repack()
        self.last_losses = None

    # ---------------------------------------------------------------- views
    def view(self, flat, name):
        o, sz, shape = self.index[name]
        return flat[o:o + sz].view(shape)

    def weights(self, name):
        return self.view(self.w_flat, name)

    def grads(self, name):
        return self.view(self.g_flat, name)

    def _heads_w(self, flat):
        o, _, _ = self.index["RPN/rpn_cls_score/W"]
        return flat[o:o + self.n_heads * 512].view(self.n_heads, 512)

    def _heads_b(self, flat):
        o, _, _ = self.index["RPN/rpn_cls_score/b"]
        return flat[o:o + self.n_heads]

    def _gact(self, h, w, c, tag=0):
        key = (h, w, c, tag)
        if key not in self._acts:
            hi = torch.empty((h, w, c), dtype=torch.bfloat16, device=self.device)
            self._acts[key] = ops.Act(hi, torch.empty_like(hi) if self.x3 else None)
        return self._acts[key]

    def _tbuf(self, planes, c, h, w):
        key = (planes, c, h, w)
        if key not in self._tbufs:
            self._tbufs[key] = train_ops.TBuf(planes, c, h, w, self.device, self.x3)
        return self._tbufs[key]

    # ---------------------------------------------------------------- weight (re)packing from the fp32 masters
    def repack(self):
        for i, L in enumerate(self.layers):
            w, b = self.weights(L["name"] + "/W"), self.weights(L["name"] + "/b")
            if i == 0:
                fwd = ops.pack_conv_weights_im2col(w, precision=self.precision)
                dg = None                                                   # the image needs no gradient
            else:
                fwd = ops.pack_conv_weights(w, cin_pad=L["cin"], precision=self.precision)
                dg = train_ops.pack_conv_weights_dgrad(w, cout_pad=L["cout"], x3=self.x3)
            self.packed[L["name"]] = (fwd, ops.pad_bias(b, L["cout"]), dg)
        wh, bh = self._heads_w(self.w_flat), self._heads_b(self.w_flat)
        fwd = ops.pack_conv_weights(wh, precision=self.precision)
        dg = train_ops.pack_conv_weights_dgrad(wh, cout_pad=self.head_ld, x3=self.x3)     # [1, 512, head_ld]
        self.packed["heads"] = (fwd, ops.pad_bias(bh, self.head_ld), dg)

    # ---------------------------------------------------------------- forward
    def forward(self, x_chw, gt_boxes, im_info=None, disable_pos=None):
        """x_chw (3,H,W) float32 CUDA; gt_boxes [G,5] float32 CUDA.  Returns losses float32[4] (device):
        rpn_loss_cls, rpn_loss_bbox, rpn_cls_accuracy, rpn_loss.  (The reference also runs the ProposalLayer here,
        region_proposal_network.py:122-124; its output does not enter the RPN loss and is skipped.)"""
        im_h, im_w = (self.H, self.W) if im_info is None else (int(im_info[0]), int(im_info[1]))
        self.x_in.copy_(x_chw, non_blocking=True)
        ops.pack_image_im2col(self.x_in, out=self.x_col)
        x = self.x_col
        for i, L in enumerate(self.layers[:13]):
            (hi, lo), b, _ = self.packed[L["name"]]
            ops.conv2d(x, hi, lo, b, 1 if i == 0 else 3, True, out=L["y"])
            x = L["y"]
            if L["pool"]:
                ops.maxpool2x2_ceil(L["y"], out=L["p"])
                x = L["p"]
        self.feat = x
        R = self.layers[13]
        (hi, lo), b, _ = self.packed[R["name"]]
        ops.conv2d(x, hi, lo, b, 3, True, out=R["y"])
        (hi, lo), b, _ = self.packed["heads"]
        ops.conv2d(R["y"], hi, lo, b, 1, False, out_act=False, ld_f32=self.head_ld, out_f32=self.head_out)
        if disable_pos is not None:
            mode, kw = train_ops.SUBSAMPLE_LIST, dict(disable_pos=disable_pos)
        elif self.subsample == "device":
            mode, kw = train_ops.SUBSAMPLE_DEVICE, dict(seed=self.seed)
            self.seed += 1
        else:
            mode, kw = train_ops.SUBSAMPLE_NONE, {}
        train_ops.anchor_targets(self.anchors, self.A, self.fh, self.fw, self.feat_stride, gt_boxes, im_h, im_w, mode=mode,
                                 work=self.targets, **kw)
        losses, self.head_grad, _ = train_ops.rpn_loss(self.head_out, None, self.anchors, self.A, self.fh, self.fw,
                                                       self.feat_stride, im_h, im_w, self.targets, delta=self.delta,
                                                       loss_lambda=self.loss_lambda, layout="nhwc", ld=self.head_ld)
        self.last_losses = losses
        return losses

    # ---------------------------------------------------------------- backward
    def _splits(self, groups, M, N, kb_total):
        # measured (tests/gpu_wgrad_tune.py): ~2 waves of 148 tiles, 4 when the M tile is half empty (64 output channels)
        tiles = groups * ((M + 127) // 128) * ((N + 127) // 128)
        target = (4 if M <= 64 else 2) * 148
        s = max(1, (target + tiles - 1) // tiles)
        return int(min(s, max(1, kb_total // 8)))

    def _wgrad(self, dyT, xT, M, N, groups, dw, m_out=None):
        """dw (flat view) = sum_pixels dY (x) X over the padded pixel axis; dyT: TBuf 1 plane [M_parts, Kp]; xT: TBuf."""
        K = dyT.Kp
        a_hi, a_lo = dyT.hi[0], (dyT.lo[0] if dyT.lo is not None else None)
        if groups == 9:
            b_hi, b_lo = xT.hi, xT.lo
        else:
            pl = 1 if xT.planes == 3 else 0
            b_hi, b_lo = xT.hi[pl], (xT.lo[pl] if xT.lo is not None else None)
        parts = train_ops.gemm_nt_splitk(a_hi, a_lo, b_hi, b_lo, groups=groups, row_stride=dyT.Wp,
                                         splits=self._splits(groups, M, N, K // 64))
        train_ops.wgrad_reduce(parts, m_out or M, N, dw)

    def backward(self, debug=None):
        """Fills g_flat with d(rpn_loss)/d(parameter) for every trainable parameter.  debug: a dict that receives, per
        layer name, float32 (C,H,W) copies of the incoming gradient ("g_in"), the masked / routed gradient ("dy") and the
        data gradient the layer passes on ("g_out") -- the buffers themselves are reused from layer to layer."""
        fh, fw, ld = self.fh, self.fw, self.head_ld
        R = self.layers[13]
        # ---- twin 1x1 heads: dY = head_grad [fh*fw, ld] fp32 (columns >= 6A are zero)
        dyh = self._gact(fh, fw, ld)
        dyhT = self._tbuf(1, ld, fh, fw)
        train_ops.grad_prepare(fh, fw, ld, g_f32=self.head_grad, out=dyh, tbuf=dyhT)
        midT = self._tbuf(3, 512, fh, fw)
        train_ops.grad_prepare(fh, fw, 512, g=R["y"], tbuf=midT)
        self._wgrad(dyhT, midT, ld, 512, 1, self._heads_w(self.g_flat), m_out=self.n_heads)
        train_ops.bias_grad(dyhT, self.n_heads, self._heads_b(self.g_flat))
        _, _, (dhi, dlo) = self.packed["heads"]
        g = self._gact(fh, fw, 512, tag=1)
        ops.conv2d(dyh, dhi, dlo, self.zero_bias, 1, False, out=g)
        if debug is not None:
            debug["heads"] = dict(dy=dyh.to_chw_f32().clone(), g_out=g.to_chw_f32().clone())
        # ---- RPN 3x3 conv, then the trunk, last layer first
        chain = [R] + self.layers[12::-1]
        for L in chain:
            h, w, co, ci = L["H"], L["W"], L["cout"], L["cin"]
            first = L is self.layers[0]
            dy = self._gact(h, w, co)
            dyT = self._tbuf(1, co, h, w)
            train_ops.grad_prepare(h, w, co, g=g, y=L["y"], p=L.get("p") if L["pool"] else None, out=dy, tbuf=dyT)
            if debug is not None:
                debug[L["name"]] = dict(g_in=g.to_chw_f32().clone(), dy=dy.to_chw_f32().clone())
            x = self._input_of(L)
            if first:
                xT = self._tbuf(1, 32, h, w)
                train_ops.grad_prepare(h, w, 32, g=x, tbuf=xT)
                dwc = torch.empty((co, 32), dtype=torch.float32, device=self.device)
                self._wgrad(dyT, xT, co, 32, 1, dwc)
                # im2col column k = tap*3 + c  ->  OIHW [co][c][tap]
                self.grads(L["name"] + "/W").copy_(dwc[:, :27].reshape(co, 9, 3).permute(0, 2, 1).reshape(co, 3, 3, 3))
            else:
                xT = self._tbuf(3, ci, h, w)
                train_ops.grad_prepare(h, w, ci, g=x, tbuf=xT)
                self._wgrad(dyT, xT, co, ci, 9, self.grads(L["name"] + "/W"))
            train_ops.bias_grad(dyT, co, self.grads(L["name"] + "/b"))
            if not first:
                _, _, (dhi, dlo) = self.packed[L["name"]]
                g = self._gact(h, w, ci, tag=1)
                ops.conv2d(dy, dhi, dlo, self.zero_bias, 3, False, out=g)
                if debug is not None:
                    debug[L["name"]]["g_out"] = g.to_chw_f32().clone()

    def _input_of(self, L):
        i = self.layers.index(L)
        if i == 0:
            return self.x_col
        if i == 13:
            return self.feat
        prev = self.layers[i - 1]
        return prev["p"] if prev["pool"] else prev["y"]

    # ---------------------------------------------------------------- optimizer
    def update(self):
        """[all-reduce SUM over ranks, like ParallelUpdater's addgrads] + WeightDecay + MomentumSGD, then repack."""
        shard.allreduce_sum_(self.g_flat, self.pg)
        train_ops.sgd_momentum(self.w_flat, self.v_flat, self.g_flat, self.lr, self.momentum, self.weight_decay)
        self.repack()

    def step(self, x_chw, gt_boxes, im_info=None, disable_pos=None):
        losses = self.forward(x_chw, gt_boxes, im_info, disable_pos)
        self.backward()
        self.update()
        return losses
