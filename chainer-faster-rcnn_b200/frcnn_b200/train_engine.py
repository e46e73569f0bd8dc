"""The reference's two training steps on the device.

RpnTrainer -- one train_rpn.py step: forward (trunk + RPN) -> AnchorTargetLayer -> rpn_loss -> backward through the RPN
heads, the RPN 3x3 conv and the 13 trunk convs -> [gradient all-reduce] -> WeightDecay + MomentumSGD.
RcnnTrainer -- one train_rcnn.py step: trunk -> frozen test-mode RPN -> ProposalTargetLayer -> RoI pool -> fc6/fc7 with
dropout -> head losses -> backward through the head, the RoI pooling and the trunk -> the same optimizer.

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

ROIS_PER_IMAGE, FG_FRACTION = 128, 0.25                          # models/proposal_target_layer.py:49-50
FG_THRESH, BG_THRESH_HI, BG_THRESH_LO = 0.5, 0.5, 0.1            # :46-48
TRUNK = [it for it in VGG16_LAYERS if it != "pool"]              # (name, cin, cout) x 13
POOL_AFTER = {VGG16_LAYERS[i - 1][0] for i, it in enumerate(VGG16_LAYERS) if it == "pool"}


class _ConvTrainer(object):
    """Shared by RpnTrainer / RcnnTrainer: flat float32 master / momentum / gradient buffers, the 13 trunk layers with
    their stored activations, weight (re)packing, the trunk forward, the generic conv-backward chain and the optimizer."""

    def _init_common(self, params, H, W, anchors, precision, lr, momentum, weight_decay, feat_stride, device, process_group,
                     extra_layers, extra_params, load_transform=None):
        if precision not in ops.PRECISIONS:
            raise FrcnnError("precision must be one of %s" % (ops.PRECISIONS,))
        self.x3 = precision == "bf16x3"
        self.precision, self.H, self.W = precision, H, W
        self.lr, self.momentum, self.weight_decay = lr, momentum, weight_decay
        self.feat_stride, self.pg = feat_stride, process_group
        dev = self.device = torch.device(device)
        self.anchors = torch.from_numpy(np.ascontiguousarray(anchors, dtype=np.float64)).to(dev)
        self.A = int(self.anchors.shape[0])
        # ---- trainable parameters in the reference's naming; one flat fp32 master / momentum / gradient buffer
        self.layers = [dict(name="trunk/" + n, cin=ci, cout=co, k=3) for n, ci, co in TRUNK] + list(extra_layers)
        names, shapes = [], []
        for L in self.layers:
            names += [L["name"] + "/W", L["name"] + "/b"]
            shapes += [(L["cout"], L["cin"], L["k"], L["k"]), (L["cout"],)]
        for n, sh in extra_params:              # stored in the given order with NO padding (adjacent twin heads stay one matrix)
            names.append(n)
            shapes.append(tuple(sh))
        sizes = [int(np.prod(s)) for s in shapes]
        offs = [0] + [int(v) for v in np.cumsum(sizes)[:-1]]
        total = int(sum(sizes))
        self.w_flat = torch.zeros((total,), dtype=torch.float32, device=dev)
        self.v_flat = torch.zeros_like(self.w_flat)
        self.g_flat = torch.zeros_like(self.w_flat)
        self.index = {}
        for n, s, o, sz in zip(names, shapes, offs, sizes):
            self.index[n] = (o, sz, s)
            src = np.ascontiguousarray(params[n], dtype=np.float32)
            if load_transform is not None:
                src = load_transform(n, src)
            self.w_flat[o:o + sz] = torch.from_numpy(np.ascontiguousarray(src).reshape(-1)).to(dev)
        # ---- geometry + activations (forward keeps every conv output; pooled maps where the graph pools)
        x3 = self.x3

        def act(h, w, c):
            hi = torch.empty((h, w, c), dtype=torch.bfloat16, device=dev)
            return ops.Act(hi, torch.empty_like(hi) if x3 else None)

        self._act = act
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
        self.zero_bias = torch.zeros((32768,), dtype=torch.float32, device=dev)
        # gradient activations (NHWC) and transposed buffers, shared between layers of identical shape
        self._acts, self._tbufs = {}, {}
        self.packed = {}
        self.last_losses = None
        # ---- multi-GPU gradient exchange (train_rpn.py:169-174): "fp32" = one exact SUM of the fp32 bucket (the parity
        # mode: bucket == sum of the per-image gradients bit for bit); "bf16" = BASELINE config #5: the bucket is rounded
        # to bf16 for the all-reduce (half the bytes), masters / momentum stay fp32.  The exchange runs on its own stream
        # in two buckets: the deep layers (conv4_1 ... heads, 90 % of the parameters) are reduced while conv3_3 ... conv1_1
        # still back-propagate; only the small shallow bucket is exposed.
        self.grad_dtype = "fp32"
        self.g_bf16 = None
        self.comm_stream = None
        self._comm_done = None
        self._bucket_lo = None              # flat offset where the early (deep-layer) bucket starts
        self.n_updates = 0                  # optimizer steps taken (the drop-in model syncs its Link params when this moves)
        self._ev_bwd = self._ev_comm = None


class RpnTrainer(_ConvTrainer):
    """Weights, optimizer state and buffers of the RPN-mode training step for one image shape."""

    def __init__(self, params, H, W, anchors, precision="bf16x3", lr=0.001, momentum=0.9, weight_decay=0.0005,
                 feat_stride=16, loss_lambda=1.0, delta=3.0, subsample="device", seed=0, device="cuda", process_group=None):
        A = len(anchors)
        heads = [("RPN/rpn_cls_score", 2 * A), ("RPN/rpn_bbox_pred", 4 * A)]
        # the twin heads are stored adjacently so that (W_cls | W_bbox) is ONE [6A, 512] matrix: cls W, bbox W, cls b, bbox b
        extra = [(n + "/W", (co, 512, 1, 1)) for n, co in heads] + [(n + "/b", (co,)) for n, co in heads]
        self._init_common(params, H, W, anchors, precision, lr, momentum, weight_decay, feat_stride, device, process_group,
                          [dict(name="RPN/rpn_conv_3x3", cin=512, cout=512, k=3)], extra)
        self.loss_lambda, self.delta, self.subsample, self.seed = loss_lambda, delta, subsample, seed
        self.n_heads = 6 * self.A
        self.head_ld = ops.round_up(self.n_heads, 32)
        R = self.layers[13]
        R["H"], R["W"], R["pool"] = self.fh, self.fw, False
        R["y"] = self._act(self.fh, self.fw, 512)
        self.head_out = torch.empty((self.fh * self.fw, self.head_ld), dtype=torch.float32, device=self.device)
        self.targets = train_ops.AnchorTargets(self.A, self.fh, self.fw, self.device)
        self.repack()

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

    def _tbuf(self, planes, c, h, w, tag=0):
        key = (planes, c, h, w, tag)
        if key not in self._tbufs:
            self._tbufs[key] = train_ops.TBuf(planes, c, h, w, self.device, self.x3)
        return self._tbufs[key]

    # ---------------------------------------------------------------- weight (re)packing from the fp32 masters
    def _repack_convs(self):
        for i, L in enumerate(self.layers):
            w, b = self.weights(L["name"] + "/W"), self.weights(L["name"] + "/b")
            if i == 0:
                fwd = ops.pack_conv_weights_im2col(w, precision=self.precision)
                dg = None                                                   # the image needs no gradient
            else:
                fwd = ops.pack_conv_weights(w, cin_pad=L["cin"], precision=self.precision)
                dg = train_ops.pack_conv_weights_dgrad(w, cout_pad=L["cout"], x3=self.x3)
            self.packed[L["name"]] = (fwd, ops.pad_bias(b, L["cout"]), dg)

    def repack(self):
        self._repack_convs()
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
        x = self._forward_trunk(x_chw)
        R = self.layers[13]
        (hi, lo), b, _ = self.packed[R["name"]]
        ops.conv2d(x, hi, lo, b, 3, True, out=R["y"])
        (hi, lo), b, _ = self.packed["heads"]
        ops.conv2d(R["y"], hi, lo, b, 1, False, out_act=False, ld_f32=self.head_ld, out_f32=self.head_out)
        if disable_pos is None and self.subsample == "numpy":
            disable_pos = self._numpy_subsample(gt_boxes, im_h, im_w)
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

    def _numpy_subsample(self, gt_boxes, im_h, im_w):
        """The reference's own subsampling (anchor_target_layer.py:148-168): labels before subsampling come back to the host
        and np.random.choice draws from NumPy's global RNG in the reference's order (a seeded run reproduces the reference's
        labels bit for bit); returns the positions to disable in the inside-compact numbering (int32 CUDA tensor)."""
        w = train_ops.anchor_targets(self.anchors, self.A, self.fh, self.fw, self.feat_stride, gt_boxes, im_h, im_w,
                                     mode=train_ops.SUBSAMPLE_NONE, work=self.targets)
        n_inside = int(w.counts[0].item())
        labels = w.labels_full[w.inds_inside[:n_inside].long()].cpu().numpy()
        num_fg = int(train_ops.RPN_FG_FRACTION * train_ops.RPN_BATCHSIZE)
        disable = [np.zeros((0,), np.int64)]
        fg_inds = np.where(labels == 1)[0]
        if len(fg_inds) > num_fg:
            d = np.random.choice(fg_inds, size=int(len(fg_inds) - num_fg), replace=False)
            labels[d] = -1
            disable.append(d)
        num_bg = train_ops.RPN_BATCHSIZE - np.sum(labels == 1)
        bg_inds = np.where(labels == 0)[0]
        if len(bg_inds) > num_bg:
            disable.append(np.random.choice(bg_inds, size=int(len(bg_inds) - num_bg), replace=False))
        return torch.from_numpy(np.concatenate(disable).astype(np.int32)).to(self.device)

    def _forward_trunk(self, x_chw):
        """The 13 VGG16 convolutions with every output kept (pools un-fused: backward needs the un-pooled maps)."""
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
        return x

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
        self._backward_chain([R] + self.layers[12::-1], g, debug)

    def _backward_chain(self, chain, g, debug=None):
        """g: gradient (ops.Act) w.r.t. the post-ReLU output of chain[0] -- at pooled resolution if that layer is followed
        by a pool.  Per layer: ReLU / pool backward + re-layout, weight + bias gradient, data gradient for the next one."""
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
            self._early_exchange(L["name"])
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

    # ---------------------------------------------------------------- gradient exchange + optimizer
    def set_grad_exchange(self, grad_dtype="fp32", overlap=True):
        """grad_dtype "fp32" (exact) or "bf16" (config #5); overlap: reduce the deep-layer bucket during the rest of backward."""
        if grad_dtype not in ("fp32", "bf16"):
            raise FrcnnError("grad_dtype must be fp32 or bf16")
        self.grad_dtype = grad_dtype
        if grad_dtype == "bf16" and self.g_bf16 is None:
            self.g_bf16 = torch.zeros(self.g_flat.shape, dtype=torch.bfloat16, device=self.device)
        self._overlap = bool(overlap)

    def _multi(self):
        return shard.dist.is_available() and shard.dist.is_initialized() and shard.dist.get_world_size(self.pg) > 1

    def _exchange(self, lo, hi):
        """All-reduce(SUM) of g_flat[lo:hi] on the communication stream, after everything queued so far on the compute stream."""
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream(device=self.device)
        ev = torch.cuda.Event()
        ev.record()
        self.comm_stream.wait_event(ev)
        with torch.cuda.stream(self.comm_stream):
            if self.grad_dtype == "bf16":
                train_ops.cast_f32_bf16(self.g_flat[lo:hi], self.g_bf16[lo:hi])
                shard.allreduce_sum_(self.g_bf16[lo:hi], self.pg)
            else:
                shard.allreduce_sum_(self.g_flat[lo:hi], self.pg)

    def _early_exchange(self, layer_name):
        """Called from the backward chain right after `layer_name`'s gradients are enqueued."""
        if not getattr(self, "_overlap", True) or not self._multi() or layer_name != "trunk/conv4_1":
            return
        lo = self.index["trunk/conv4_1/W"][0]
        if lo % 4 != 0:
            return
        self._bucket_lo = lo
        self._exchange(lo, self.g_flat.numel())

    def update(self):
        """[all-reduce SUM over ranks, like ParallelUpdater's addgrads] + WeightDecay + MomentumSGD, then repack."""
        if self._multi():
            hi = self._bucket_lo if self._bucket_lo is not None else self.g_flat.numel()
            self._ev_bwd = torch.cuda.Event(enable_timing=True)
            self._ev_bwd.record()                                  # backward fully enqueued on the compute stream
            self._exchange(0, hi)
            self._ev_comm = torch.cuda.Event(enable_timing=True)
            self._ev_comm.record(self.comm_stream)
            torch.cuda.current_stream().wait_event(self._ev_comm)
            self._bucket_lo = None
        if self.grad_dtype == "bf16" and self._multi():
            train_ops.sgd_momentum_bf16g(self.w_flat, self.v_flat, self.g_bf16, self.lr, self.momentum, self.weight_decay)
        else:
            train_ops.sgd_momentum(self.w_flat, self.v_flat, self.g_flat, self.lr, self.momentum, self.weight_decay)
        self.repack()
        self.n_updates += 1

    def last_exposed_exchange_ms(self):
        """Time between "backward done" on the compute stream and "exchange done" on the communication stream for the last
        step (what the optimizer had to wait for); call after a synchronize.  0 if the exchange finished first."""
        if self._ev_bwd is None or self._ev_comm is None:
            return 0.0
        return max(0.0, self._ev_bwd.elapsed_time(self._ev_comm))

    def export_params(self):
        """Trainable parameters as numpy arrays under the reference's names (checkpoint / write-back into the Links)."""
        return {n: self.weights(n).detach().cpu().numpy().copy() for n in self.index}

    def step(self, x_chw, gt_boxes, im_info=None, disable_pos=None):
        losses = self.forward(x_chw, gt_boxes, im_info, disable_pos)
        self.backward()
        self.update()
        return losses


class RcnnTrainer(RpnTrainer):
    """One train_rcnn.py step (`model.rcnn_train = True`, /root/reference models/faster_rcnn.py:114-173): trunk -> RPN in TEST
    mode (its weights are frozen here: proposals are plain arrays, faster_rcnn.py:117-120) -> ProposalTargetLayer ->
    RoI pooling of ALL proposals -> fc6 / fc7 with dropout -> cls_score | bbox_pred -> losses on the kept rows ->
    backward through the head, the RoI pooling and the 13 trunk convolutions -> WeightDecay + MomentumSGD.
    Trainable: trunk/*, fc6, fc7, cls_score, bbox_pred (the RPN links receive no gradient: cleargrads semantics).

    fc6/W is kept internally in the (h, w, c) input order the RoI-pool kernel emits (`export_params` converts back).
    Sampling (`sample="numpy"`): the proposals' best overlaps come back to the host and np.random.choice draws exactly as
    proposal_target_layer.py:99-129 does, so a seeded run reproduces the reference's kept set; `keep_inds=` pins it.
    Dropout masks are explicit inputs (uint8 [post_n, 4096] x 2) or drawn from a torch generator."""

    def __init__(self, params, H, W, anchors, precision="bf16x3", lr=0.001, momentum=0.9, weight_decay=0.0005,
                 feat_stride=16, num_classes=21, delta=1.0, post_n=300, pre_n=6000, nms_thresh=0.7, min_size=16,
                 sample="numpy", seed=0, dropout=True, device="cuda", process_group=None):
        nc = self.num_classes = num_classes
        heads = [("cls_score", nc), ("bbox_pred", 4 * nc)]
        extra = [("fc6/W", (4096, 512 * 49)), ("fc6/b", (4096,)), ("fc7/W", (4096, 4096)), ("fc7/b", (4096,))]
        extra += [(n + "/W", (co, 4096)) for n, co in heads] + [(n + "/b", (co,)) for n, co in heads]

        def to_hwc(name, a):                         # fc6 columns (c, h, w) -> (h, w, c)
            return a.reshape(4096, 512, 49).transpose(0, 2, 1) if name == "fc6/W" else a
        self._init_common(params, H, W, anchors, precision, lr, momentum, weight_decay, feat_stride, device, process_group,
                          [], extra, load_transform=to_hwc)
        self.delta, self.sample, self.dropout = delta, sample, dropout
        self.post_n, self.pre_n, self.nms_thresh, self.min_size = post_n, pre_n, nms_thresh, min_size
        self.n_heads = 5 * nc
        self.head_ld = ops.round_up(self.n_heads, 32)
        dev = self.device
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(seed)
        # frozen RPN (forward only)
        T = lambda k: torch.from_numpy(np.ascontiguousarray(params[k], dtype=np.float32)).to(dev)
        hi, lo = ops.pack_conv_weights(T("RPN/rpn_conv_3x3/W"), cin_pad=512, precision=precision)
        self.rpn3 = (hi, lo, ops.pad_bias(T("RPN/rpn_conv_3x3/b"), 512))
        wh = torch.cat([T("RPN/rpn_cls_score/W"), T("RPN/rpn_bbox_pred/W")], dim=0)
        bh = torch.cat([T("RPN/rpn_cls_score/b"), T("RPN/rpn_bbox_pred/b")], dim=0)
        self.rpn_ld = ops.round_up(6 * self.A, 32)
        hi, lo = ops.pack_conv_weights(wh, precision=precision)
        self.rpn_heads = (hi, lo, ops.pad_bias(bh, self.rpn_ld))
        fh, fw = self.fh, self.fw
        self.rpn_mid = self._act(fh, fw, 512)
        self.rpn_out = torch.empty((fh * fw, self.rpn_ld), dtype=torch.float32, device=dev)
        self.prop = ops.ProposalWorkspace(self.A, fh, fw, pre_n, post_n, dev)
        # head activations
        self.pool5 = self._act(1, post_n, 49 * 512)
        self.fc6 = self._act(1, post_n, 4096)
        self.fc7 = self._act(1, post_n, 4096)
        self.head_out = torch.empty((post_n, self.head_ld), dtype=torch.float32, device=dev)
        self.dfeat = torch.empty((fh * fw, 512), dtype=torch.float32, device=dev)
        self.roi_ws = None
        self.repack()

    # ---------------------------------------------------------------- parameters
    def _heads_w(self, flat):
        o, _, _ = self.index["cls_score/W"]
        return flat[o:o + self.n_heads * 4096].view(self.n_heads, 4096)

    def _heads_b(self, flat):
        o, _, _ = self.index["cls_score/b"]
        return flat[o:o + self.n_heads]

    def export_params(self):
        """Trainable parameters as numpy arrays in the reference's layouts (fc6/W back to (c, h, w) columns)."""
        out = {}
        for n in self.index:
            a = self.weights(n).detach().cpu().numpy().copy()
            out[n] = a.reshape(4096, 49, 512).transpose(0, 2, 1).reshape(4096, 512 * 49) if n == "fc6/W" else a
        return out

    def repack(self):
        self._repack_convs()
        for n in ("fc6", "fc7"):
            w, b = self.weights(n + "/W"), self.weights(n + "/b")
            fwd = ops.pack_conv_weights(w, precision=self.precision)
            dg = train_ops.pack_conv_weights_dgrad(w, cout_pad=4096, x3=self.x3)
            self.packed[n] = (fwd, ops.pad_bias(b, 4096), dg)
        wh, bh = self._heads_w(self.w_flat), self._heads_b(self.w_flat)
        fwd = ops.pack_conv_weights(wh, precision=self.precision)
        dg = train_ops.pack_conv_weights_dgrad(wh, cout_pad=self.head_ld, x3=self.x3)        # [1, 4096, head_ld]
        self.packed["heads"] = (fwd, ops.pad_bias(bh, self.head_ld), dg)

    # ---------------------------------------------------------------- ProposalTargetLayer sampling (host, like the reference)
    def _sample(self, max_ov, R):
        mo = max_ov[:R].cpu().numpy()
        n_fg_cap = int(FG_FRACTION * ROIS_PER_IMAGE)
        fg = np.where(mo >= FG_THRESH)[0]                                                    # :99
        n_fg = min(n_fg_cap, fg.size)
        if fg.size > 0:
            fg = np.random.choice(fg, size=n_fg, replace=False)                              # :105-110
        bg = np.where((mo < BG_THRESH_HI) & (mo >= BG_THRESH_LO))[0]                         # :113-114
        n_bg = min(ROIS_PER_IMAGE - n_fg, bg.size)
        if bg.size > 0:
            bg = np.random.choice(bg, size=n_bg, replace=False)                              # :119-126
        return np.concatenate([fg, bg]).astype(np.int32)

    # ---------------------------------------------------------------- forward
    def forward(self, x_chw, gt_boxes, im_info=None, keep_inds=None, masks=None):
        """Returns losses float32[4] (device): loss_cls, loss_bbox, cls_accuracy, loss_rcnn."""
        im_h, im_w = (self.H, self.W) if im_info is None else (int(im_info[0]), int(im_info[1]))
        feat = self._forward_trunk(x_chw)
        hi, lo, b = self.rpn3
        ops.conv2d(feat, hi, lo, b, 3, True, out=self.rpn_mid)
        hi, lo, b = self.rpn_heads
        ops.conv2d(self.rpn_mid, hi, lo, b, 1, False, out_act=False, ld_f32=self.rpn_ld, out_f32=self.rpn_out)
        ops.proposals(self.rpn_out, None, self.anchors, self.A, self.fh, self.fw, self.feat_stride, im_h, im_w, self.min_size,
                      self.pre_n, self.post_n, self.nms_thresh, layout="nhwc", ld=self.rpn_ld, cls_is_logits=True, work=self.prop)
        rois, count = self.prop.rois, self.prop.count
        self.max_ov, self.argmax = train_ops.roi_overlaps(rois, count, gt_boxes)
        if keep_inds is None:
            R = int(count.item())                    # the reference sizes its arrays from len(proposals) too
            keep_inds = torch.from_numpy(self._sample(self.max_ov, R)).to(self.device)
        self.keep = keep_inds.to(device=self.device, dtype=torch.int32).contiguous()
        if self.keep.numel() < 1:
            raise FrcnnError("ProposalTargetLayer kept no RoI (no proposal overlaps a ground-truth box by >= 0.1)")
        self.use_gt, self.ext, self.labels = train_ops.roi_targets(rois, gt_boxes, self.argmax, self.keep, self.num_classes)
        ops.roi_pool(feat, rois, count, 7, 7, 1.0 / self.feat_stride, out=self.pool5)
        if masks is None and self.dropout:
            masks = [(torch.rand((self.post_n, 4096), device=self.device, generator=self.gen) >= 0.5).to(torch.uint8) for _ in range(2)]
        self.masks = masks
        (hi, lo), b, _ = self.packed["fc6"]
        ops.conv2d(self.pool5, hi, lo, b, 1, True, out=self.fc6, m_valid=count)
        if masks is not None:
            train_ops.dropout_(self.fc6, masks[0])                                           # F.dropout(train=rcnn_train), :127
        (hi, lo), b, _ = self.packed["fc7"]
        ops.conv2d(self.fc6, hi, lo, b, 1, True, out=self.fc7, m_valid=count)
        if masks is not None:
            train_ops.dropout_(self.fc7, masks[1])                                           # :128
        (hi, lo), b, _ = self.packed["heads"]
        ops.conv2d(self.fc7, hi, lo, b, 1, False, out_act=False, ld_f32=self.head_ld, out_f32=self.head_out, m_valid=count)
        losses, self.head_grad = train_ops.rcnn_loss(self.head_out, self.keep, self.labels, self.ext, self.num_classes, self.delta)
        self.last_losses = losses
        return losses

    # ---------------------------------------------------------------- backward
    def _fc_wgrad(self, dyT, xT, M, N, dw, m_out=None):
        """dW [M][N] = dY^T X over the (padded) RoI axis; written straight into the gradient view when the shapes allow."""
        a_hi, a_lo = dyT.hi[0], (dyT.lo[0] if dyT.lo is not None else None)
        b_hi, b_lo = xT.hi[0], (xT.lo[0] if xT.lo is not None else None)
        if m_out is None and N % 32 == 0:
            train_ops.gemm_nt_splitk(a_hi, a_lo, b_hi, b_lo, groups=1, splits=1, out=dw.reshape(-1))
        else:
            parts = train_ops.gemm_nt_splitk(a_hi, a_lo, b_hi, b_lo, groups=1, splits=1)
            train_ops.wgrad_reduce(parts, m_out or M, N, dw)

    def backward(self, debug=None):
        R_cap, ld = self.post_n, self.head_ld
        twice = self.masks is not None               # ratio-0.5 dropout: dz = 2 * g * (y > 0), y = the post-dropout activation
        # ---- cls_score | bbox_pred: dY = head_grad [R_cap, ld] fp32 (zero rows for RoIs that were not kept)
        dyh = self._gact(1, R_cap, ld)
        dyhT = self._tbuf(1, ld, 1, R_cap)
        train_ops.grad_prepare(1, R_cap, ld, g_f32=self.head_grad, out=dyh, tbuf=dyhT)
        fc7T = self._tbuf(1, 4096, 1, R_cap)
        train_ops.grad_prepare(1, R_cap, 4096, g=self.fc7, tbuf=fc7T)
        self._fc_wgrad(dyhT, fc7T, ld, 4096, self._heads_w(self.g_flat), m_out=self.n_heads)
        train_ops.bias_grad(dyhT, self.n_heads, self._heads_b(self.g_flat))
        _, _, (dhi, dlo) = self.packed["heads"]
        g = self._gact(1, R_cap, 4096, tag=1)
        ops.conv2d(dyh, dhi, dlo, self.zero_bias, 1, False, out=g)
        # ---- fc7
        dy7 = self._gact(1, R_cap, 4096)
        dy7T = self._tbuf(1, 4096, 1, R_cap, tag=1)
        train_ops.grad_prepare(1, R_cap, 4096, g=g, y=self.fc7, out=dy7, tbuf=dy7T, times2=twice)
        if debug is not None:
            debug["fc7"] = dict(g_in=g.to_chw_f32().clone(), dy=dy7.to_chw_f32().clone())
        fc6T = self._tbuf(1, 4096, 1, R_cap)          # the buffer of fc7T: the heads' weight gradient has consumed it
        train_ops.grad_prepare(1, R_cap, 4096, g=self.fc6, tbuf=fc6T)
        self._fc_wgrad(dy7T, fc6T, 4096, 4096, self.grads("fc7/W"))
        train_ops.bias_grad(dy7T, 4096, self.grads("fc7/b"))
        _, _, (dhi, dlo) = self.packed["fc7"]
        g6 = self._gact(1, R_cap, 4096, tag=2)
        ops.conv2d(dy7, dhi, dlo, self.zero_bias, 1, False, out=g6)
        # ---- fc6
        dy6 = self._gact(1, R_cap, 4096)
        train_ops.grad_prepare(1, R_cap, 4096, g=g6, y=self.fc6, out=dy6, tbuf=dy7T, times2=twice)      # dy7T's buffer now holds dy6^T
        if debug is not None:
            debug["fc6"] = dict(g_in=g6.to_chw_f32().clone(), dy=dy6.to_chw_f32().clone())
        p5T = self._tbuf(1, 49 * 512, 1, R_cap)
        train_ops.grad_prepare(1, R_cap, 49 * 512, g=self.pool5, tbuf=p5T)
        self._fc_wgrad(dy7T, p5T, 4096, 49 * 512, self.grads("fc6/W"))
        train_ops.bias_grad(dy7T, 4096, self.grads("fc6/b"))
        _, _, (dhi, dlo) = self.packed["fc6"]
        gp = self._gact(1, R_cap, 49 * 512, tag=1)
        ops.conv2d(dy6, dhi, dlo, self.zero_bias, 1, False, out=gp)
        # ---- RoI pooling backward -> gradient of conv5_3's output, then the trunk
        if self.roi_ws is None:
            self.roi_ws = torch.empty((ops._lib.load().frcnn_roi_pool_backward_workspace_bytes(self.fh, self.fw, 512),),
                                      dtype=torch.uint8, device=self.device)
        train_ops.roi_pool_backward(self.feat, self.prop.rois, self.prop.count, gp, 7, 7, 1.0 / self.feat_stride, out=self.dfeat,
                                    ws=self.roi_ws)
        gfeat = self._gact(self.fh, self.fw, 512, tag=3)
        train_ops.grad_prepare(self.fh, self.fw, 512, g_f32=self.dfeat, out=gfeat)
        if debug is not None:
            debug["roi"] = dict(dpool5=gp.to_chw_f32().clone(), dfeat=self.dfeat.clone())
        self._backward_chain(self.layers[12::-1], gfeat, debug)

    def step(self, x_chw, gt_boxes, im_info=None, keep_inds=None, masks=None):
        losses = self.forward(x_chw, gt_boxes, im_info, keep_inds, masks)
        self.backward()
        self.update()
        return losses

