"""The whole Faster R-CNN forward detection path as one replayable CUDA graph.

Mirrors FasterRCNN.__call__ (inference branch, /root/reference models/faster_rcnn.py:92-134,175-178):

    input                                    frcnn_pack_image_im2col3x3 (float CHW) or frcnn_preprocess_bgr8 (raw uint8)
    trunk (models/vgg16.py:38-82)            13x frcnn_conv2d(3x3)+ReLU (first layer = K=32 GEMM over the im2col image),
                                             the 4 ceil-mode 2x2 max-pools fused into the producing conv's epilogue
    RPN   (models/region_proposal_network.py:117-124)
                                             frcnn_conv2d(3x3)+ReLU, ONE frcnn_conv2d(1x1) for the twin
                                             heads (18 cls + 36 bbox -> fp32 [H*W, 64])
    ProposalLayer (models/proposal_layer.py:102-198)
                                             frcnn_proposals (18-way softmax fused, NHWC logits in)
    RoI pool + head (models/faster_rcnn.py:123-134)
                                             frcnn_roi_pool, frcnn_conv2d as GEMM for fc6 / fc7 /
                                             (cls_score | bbox_pred merged: 105 -> fp32 [R, 128])
    tail  (models/faster_rcnn.py:175-178)    frcnn_head_decode (softmax + decode + clip)
    caller's per-class NMS (forward.py:48-57) frcnn_detect (optional)

All buffers are allocated once per (image shape, mode); the data-dependent RoI count stays on the
device (rows past it are zero) so the launch sequence is static and is captured into a CUDA graph.
"""
import numpy as np
import torch

from . import ops
from ._lib import FrcnnError

import threading
import time

_CAPTURE_LOCK = threading.Lock()
# diagnostic hook: a list that receives, per ForwardPlan.forward_host call, the host-side milliseconds of
# (pageable->pinned copy, H2D enqueue, graph replay enqueue, D2H enqueue + wait for the result); None = off
HOST_PROFILE = None

VGG16_LAYERS = [
    ("conv1_1", 3, 64), ("conv1_2", 64, 64), "pool",
    ("conv2_1", 64, 128), ("conv2_2", 128, 128), "pool",
    ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256), "pool",
    ("conv4_1", 256, 512), ("conv4_2", 512, 512), ("conv4_3", 512, 512), "pool",
    ("conv5_1", 512, 512), ("conv5_2", 512, 512), ("conv5_3", 512, 512),
]


def _t(a, device):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float32)
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)


class PackedWeights(object):
    """Reference-named fp32 parameters (trunk/conv1_1/W ... bbox_pred/b) repacked once
    into the kernels' layouts: [taps, Cout, Cin] bf16 hi(/lo) + padded fp32 bias."""

    def __init__(self, params, precision="bf16x3", device="cuda", num_classes=21, n_anchors=9,
                 pool_chw=(512, 7, 7)):
        if precision not in ops.PRECISIONS:
            raise FrcnnError("unknown precision %r" % (precision,))
        self.precision, self.device = precision, device
        self.num_classes, self.n_anchors = num_classes, n_anchors
        P = lambda k: _t(params[k], device)
        self.convs = {}
        self._pack_trunk(params, P)
        self.rpn3 = self._pack(P("RPN/rpn_conv_3x3/W"), P("RPN/rpn_conv_3x3/b"))
        # twin 1x1 heads merged along Cout: rows [0,2A) = rpn_cls_score, [2A,6A) = rpn_bbox_pred
        wh = torch.cat([P("RPN/rpn_cls_score/W"), P("RPN/rpn_bbox_pred/W")], dim=0)
        bh = torch.cat([P("RPN/rpn_cls_score/b"), P("RPN/rpn_bbox_pred/b")], dim=0)
        self.rpn_ld = ops.round_up(6 * n_anchors, 32)
        self.rpn_heads = self._pack(wh, bh, n_bias=self.rpn_ld)
        self.fc6 = self._pack(P("fc6/W"), P("fc6/b"), perm_chw=pool_chw)
        self.fc7 = self._pack(P("fc7/W"), P("fc7/b"))
        wc = torch.cat([P("cls_score/W"), P("bbox_pred/W")], dim=0)
        bc = torch.cat([P("cls_score/b"), P("bbox_pred/b")], dim=0)
        self.head_ld = ops.round_up(5 * num_classes, 32)
        self.head = self._pack(wc, bc, n_bias=self.head_ld)
        torch.cuda.synchronize(device)

    def _pack_trunk(self, params, P):
        """VGG16 (models/vgg16.py:38-69); subclasses pack another trunk (resnet_engine.ResNetPackedWeights)."""
        for item in VGG16_LAYERS:
            if item == "pool":
                continue
            name, cin, cout = item
            w = P("trunk/%s/W" % name)
            if cin <= 3:
                # first layer as a K = 3 x 32 GEMM over the compact [H][W+2][8] image (ops.pack_image_c8 / conv3x3_c8)
                hi, lo = ops.pack_conv_weights_c8(w, precision=self.precision)
                b = P("trunk/%s/b" % name)
                self.convs[name] = (hi, lo, ops.pad_bias(b, b.numel()))
            else:
                self.convs[name] = self._pack(w, P("trunk/%s/b" % name), cin_pad=cin)

    def _pack(self, w, b, cin_pad=None, perm_chw=None, n_bias=0):
        hi, lo = ops.pack_conv_weights(w, cin_pad=cin_pad, precision=self.precision, perm_chw=perm_chw)
        return hi, lo, ops.pad_bias(b, max(n_bias, b.numel()))


class ForwardPlan(object):
    """Buffers + launch sequence for one image shape."""

    def __init__(self, weights, H, W, pre_n=6000, post_n=300, nms_thresh=0.7, min_size=16, feat_stride=16,
                 anchors=None, with_detect=False, det_nms_thresh=0.3, det_conf=0.8, use_graph=True, keep_rpn_debug=False,
                 fuse_pool=True, hwc_input=False):
        self.w, self.H, self.W = weights, H, W
        self._ctor = dict(pre_n=pre_n, post_n=post_n, nms_thresh=nms_thresh, min_size=min_size, feat_stride=feat_stride,
                          anchors=anchors, with_detect=with_detect, det_nms_thresh=det_nms_thresh, det_conf=det_conf,
                          use_graph=use_graph, keep_rpn_debug=keep_rpn_debug, fuse_pool=fuse_pool, hwc_input=hwc_input)
        # hwc_input: the static input buffer holds the bytes of a dense (H,W,3) float32 image (what a forward.py-style caller's
        # `img.transpose(2,0,1).astype(np.float32)` really is in memory); the first kernel reads it with HWC strides
        self.hwc_input = bool(hwc_input)
        dev = weights.device
        x3 = weights.precision == "bf16x3"
        self.pre_n, self.post_n, self.nms_thresh, self.min_size, self.feat_stride = pre_n, post_n, nms_thresh, min_size, feat_stride
        self.with_detect, self.det_nms_thresh, self.det_conf = with_detect, det_nms_thresh, det_conf
        A = weights.n_anchors
        if anchors is None:
            raise FrcnnError("ForwardPlan needs the [A,4] float64 anchor table")
        self.anchors = torch.from_numpy(np.ascontiguousarray(anchors, dtype=np.float64)).to(dev)

        def act(h, w, c):
            hi = torch.empty((h, w, c), dtype=torch.bfloat16, device=dev)
            return ops.Act(hi, torch.empty_like(hi) if x3 else None)

        self.x_in = torch.zeros((3, H, W), dtype=torch.float32, device=dev)      # static input (C,H,W)
        self.img_info = torch.tensor([H, W], dtype=torch.int32, device=dev)      # clip bounds (h, w): static in the graph
        self._act = act
        self.fuse_pool = fuse_pool
        h, w_ = self._alloc_trunk(H, W)
        self.fh, self.fw = h, w_
        self.rpn_mid = act(h, w_, weights.rpn3[0].shape[1])
        self.rpn_out = torch.empty((h * w_, weights.rpn_ld), dtype=torch.float32, device=dev)
        # every per-image result lives in ONE device block (4-byte words) so that a host caller gets all of it with a
        # single D2H transfer: [count | prob | boxes | rois | scores | keep_count | conf_count | keep_idx]
        NC = weights.num_classes
        sizes = [("count", 4), ("prob", post_n * NC), ("boxes", post_n * 4 * NC), ("rois", post_n * 4), ("scores", post_n),
                 ("keep_count", NC - 1), ("conf_count", NC - 1), ("keep_idx", (NC - 1) * post_n)]
        self.result_layout, off = {}, 0
        for name, n_ in sizes:
            self.result_layout[name] = (off, n_)
            off += (n_ + 3) // 4 * 4
        self.result_words_nodetect = self.result_layout["keep_count"][0]
        self.result = torch.zeros((off,), dtype=torch.int32, device=dev)
        rf = self.result.view(torch.float32)
        seg = lambda t, name: t[self.result_layout[name][0]: self.result_layout[name][0] + self.result_layout[name][1]]   # noqa: E731
        self.prop = ops.ProposalWorkspace(A, h, w_, pre_n, post_n, dev, debug=keep_rpn_debug,
                                          outputs=(seg(rf, "rois").view(post_n, 4), seg(rf, "scores"), seg(self.result, "count")[:1]))
        C = self.acts[-1].hi.shape[2]
        self.pool5 = act(1, post_n, 49 * C)
        self.fc6 = act(1, post_n, 4096)
        self.fc7 = act(1, post_n, 4096)
        self.head_out = torch.empty((post_n, weights.head_ld), dtype=torch.float32, device=dev)
        # split-K slabs of the swapped-operand head GEMMs (ops.linear); one buffer sized for the largest layer
        nc5 = 5 * weights.num_classes
        self.fc_work = ops.linear_workspace(post_n, 49 * C, 4096, dev)
        for k_, c_ in ((4096, 4096), (4096, nc5)):
            wk = ops.linear_workspace(post_n, k_, c_, dev)
            if wk.numel() > self.fc_work.numel():
                self.fc_work = wk
        self.prob = seg(rf, "prob").view(post_n, NC)
        self.boxes = seg(rf, "boxes").view(post_n, 4 * NC)
        self.det = (seg(self.result, "keep_idx").view(NC - 1, post_n), seg(self.result, "keep_count"), seg(self.result, "conf_count"))
        self._host = None
        self.graph = None
        self.use_graph = use_graph
        self.im_h, self.im_w = H, W
        self.n_launches = 0

    def _alloc_trunk(self, H, W):
        """VGG16 activations; returns the feature-map size.  (resnet_engine.ResNetForwardPlan overrides both trunk hooks.)"""
        act, fuse_pool = self._act, self.fuse_pool
        n_c8 = ops._lib.load().frcnn_image_c8_elems(H, W)          # compact first-layer input: [H][W+2][8] (+ slack), flat planes
        hi0 = torch.empty((n_c8,), dtype=torch.bfloat16, device=self.x_in.device)
        self.acts = [ops.Act(hi0, torch.empty_like(hi0) if self.w.precision == "bf16x3" else None)]
        h, w_ = H, W
        self.trunk_steps = []          # (layer name, fuse the following 2x2 pool into the conv epilogue)
        for i, item in enumerate(VGG16_LAYERS):
            if item == "pool":
                continue
            fuse = fuse_pool and i + 1 < len(VGG16_LAYERS) and VGG16_LAYERS[i + 1] == "pool"
            pooled = i + 1 < len(VGG16_LAYERS) and VGG16_LAYERS[i + 1] == "pool"
            if fuse:
                h, w_ = (h + 1) // 2, (w_ + 1) // 2
                self.acts.append(act(h, w_, item[2]))
                self.trunk_steps.append((item[0], True, False))
            else:
                self.acts.append(act(h, w_, item[2]))
                self.trunk_steps.append((item[0], False, pooled))
                if pooled:
                    h, w_ = (h + 1) // 2, (w_ + 1) // 2
                    self.acts.append(act(h, w_, item[2]))
        return h, w_

    def _run_trunk(self):
        """Launches the trunk; returns (feature map Act, number of launches)."""
        w = self.w
        n = 0
        x = self.acts[0]
        ops.pack_image_c8(self.x_in, out=x, hwc_memory=self.hwc_input)
        n += 1
        i = 0
        for name, fused, pool_after in self.trunk_steps:
            hi, lo, b = w.convs[name]
            if i == 0:                          # conv1_1: sliding-window reads of the compact image, K = 3 x 32
                ops.conv3x3_c8(self.acts[0], self.H, self.W, hi, lo, b, True, out=self.acts[1])
            else:
                ops.conv2d(self.acts[i], hi, lo, b, 3, True, out=self.acts[i + 1], fuse_pool=fused)
            n += 1
            i += 1
            if pool_after:
                ops.maxpool2x2_ceil(self.acts[i], out=self.acts[i + 1])
                n += 1
                i += 1
        return self.acts[-1], n

    # -- the launch sequence (static; safe to capture)
    def _run(self):
        w = self.w
        feat, n = self._run_trunk()
        self.feat = feat
        hi, lo, b = w.rpn3
        ops.conv2d(feat, hi, lo, b, 3, True, out=self.rpn_mid)
        hi, lo, b = w.rpn_heads
        ops.conv2d(self.rpn_mid, hi, lo, b, 1, False, out_act=False, ld_f32=w.rpn_ld, out_f32=self.rpn_out)
        ops.proposals(self.rpn_out, None, self.anchors, w.n_anchors, self.fh, self.fw, self.feat_stride,
                      self.im_h, self.im_w, self.min_size, self.pre_n, self.post_n, self.nms_thresh,
                      layout="nhwc", ld=w.rpn_ld, cls_is_logits=True, work=self.prop,
                      debug=self.prop.dbg_dets is not None)
        n += 2 + 5          # rpn 3x3, rpn heads; decode, select, rank/scatter, IoU mask, mask scan
        ops.roi_pool(feat, self.prop.rois, self.prop.count, 7, 7, 1.0 / self.feat_stride, out=self.pool5)
        hi, lo, b = w.fc6
        ops.linear(self.pool5, hi, lo, b, True, m_valid=self.prop.count, out=self.fc6, work=self.fc_work)
        hi, lo, b = w.fc7
        ops.linear(self.fc6, hi, lo, b, True, m_valid=self.prop.count, out=self.fc7, work=self.fc_work)
        hi, lo, b = w.head
        ops.linear(self.fc7, hi, lo, b, False, m_valid=self.prop.count, out_f32=self.head_out, ld_f32=w.head_ld,
                   work=self.fc_work, want_act=False)
        n += 3          # each head layer = the split-K GEMM + its reduction
        ops.head_decode(self.head_out, w.head_ld, self.prop.rois, self.prop.count, w.num_classes, self.im_h, self.im_w,
                        out_prob=self.prob, out_boxes=self.boxes)
        n += 5
        if self.with_detect:
            ops.detect(self.prob, self.boxes, self.prop.count, self.det_nms_thresh, self.det_conf, out=self.det)
            n += 1
        self.n_launches = n

    def clone(self):
        """A second set of buffers (and its own graph) over the SAME packed weights: lets another image be in flight."""
        p = type(self)(self.w, self.H, self.W, **self._ctor)
        p.set_clip(self.im_h, self.im_w)
        return p

    def set_clip(self, im_h, im_w):
        """img_info as the caller passes it (forward.py:93 passes (H, H)).  Changing it
        invalidates a captured graph (the bounds are kernel arguments)."""
        if (im_h, im_w) != (self.im_h, self.im_w):
            self.im_h, self.im_w = int(im_h), int(im_w)
            self.graph = None

    # -- host-array front end (the reference's model-input interface: a float32 (3,H,W) HOST array in, host results out)
    def result_words(self):
        return int(self.result.numel()) if self.with_detect else int(self.result_words_nodetect)

    def host_io(self):
        """Pinned staging for one synchronous host call at a time: the input image and a mirror of the result block."""
        if self._host is None:
            dev = self.x_in.device
            self._host = dict(x=ops.PinnedBlock(tuple(self.x_in.shape), np.float32),
                              res=ops.PinnedBlock((int(self.result.numel()),), np.int32),
                              stream=torch.cuda.Stream(device=dev))
        return self._host

    def forward_host(self, x_np):
        """x_np: float32 (3,H,W) (or (1,3,H,W)) numpy array.  Copies it into pinned staging (multi-threaded host copy),
        uploads, replays the graph and brings the whole result block back with ONE D2H; blocks until it is there.
        Returns a dict of numpy VIEWS into the pinned mirror (valid until the next forward_host on this plan)."""
        io = self.host_io()
        prof = HOST_PROFILE
        t0 = time.perf_counter() if prof is not None else 0.0
        n = self.result_words()
        st = io["stream"]
        t1 = t0
        io["x"].upload(x_np, self.x_in, st)                        # pageable -> pinned -> device, chunk-pipelined (one C call)
        t2 = time.perf_counter() if prof is not None else 0.0
        with torch.cuda.stream(st):
            self.forward(None)
        t3 = time.perf_counter() if prof is not None else 0.0
        io["res"].d2h(self.result, st, nbytes=4 * n)               # D2H: everything the caller can ask for, one transfer
        ops.stream_synchronize(st)
        if prof is not None:
            t4 = time.perf_counter()
            prof.append((1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3)))
        return self.unpack_result(io["res"].np)

    def unpack_result(self, words):
        """numpy int32 words of a result block -> dict of views (count int, prob [R,NC], boxes [R,4NC], rois, scores, detect lists)."""
        L, post_n, NC = self.result_layout, self.post_n, self.w.num_classes
        f = words.view(np.float32)
        R = int(words[L["count"][0]])
        sl = lambda a, name: a[L[name][0]: L[name][0] + L[name][1]]          # noqa: E731
        out = dict(count=R, prob=sl(f, "prob").reshape(post_n, NC)[:R], boxes=sl(f, "boxes").reshape(post_n, 4 * NC)[:R],
                   rois=sl(f, "rois").reshape(post_n, 4)[:R], scores=sl(f, "scores")[:R])
        if self.with_detect:
            out.update(keep_idx=sl(words, "keep_idx").reshape(NC - 1, post_n), keep_count=sl(words, "keep_count"),
                       conf_count=sl(words, "conf_count"))
        return out

    def forward(self, x_chw=None):
        """Run one image.  x_chw: (3,H,W) float32 CUDA tensor (copied into the static input) or None
        to reuse the current contents of `self.x_in`.  Returns (prob, boxes, count) device tensors."""
        if x_chw is not None:
            self.x_in.copy_(x_chw, non_blocking=True)
        if not self.use_graph:
            self._run()
        else:
            if self.graph is None:
                # one capture at a time in the process (torch's capture stream is shared), and thread-local capture mode so
                # that other host threads (each running its own plan) may keep calling CUDA meanwhile
                with _CAPTURE_LOCK:
                    self._run()                   # warm-up outside capture (func attributes, lazy init)
                    torch.cuda.current_stream().synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, capture_error_mode="thread_local"):
                        self._run()
                    self.graph = g
            self.graph.replay()
        return self.prob, self.boxes, self.prop.count


class LanePool(object):
    """`lanes` independent images in flight on one GPU: one ForwardPlan (buffers + CUDA graph) and one CUDA stream
    per lane over shared weights.  The reference processes one image per call (batch hard-wired to 1, SURVEY.md
    Q10); images are independent, so a second/third image's tensor-core layers fill the SMs that image one's
    single-CTA steps (radix select, NMS scan) and ragged last waves leave idle.  Each image still runs the whole
    path; nothing is batched or skipped, and results are bit-identical to the single-lane run (tests)."""

    def __init__(self, plan, lanes=3):
        self.plans = [plan] + [plan.clone() for _ in range(max(1, int(lanes)) - 1)]
        dev = plan.x_in.device
        self.streams = [torch.cuda.Stream(device=dev) for _ in self.plans]
        for p, st in zip(self.plans, self.streams):
            if p.use_graph and p.graph is None:
                with torch.cuda.stream(st):
                    p.forward(None)
        torch.cuda.synchronize(dev)

    def __len__(self):
        return len(self.plans)

    def fork(self):
        """Lane streams start after everything queued so far on the current stream."""
        ev = torch.cuda.Event()
        ev.record()
        for st in self.streams:
            st.wait_event(ev)

    def submit(self, i, x_chw=None):
        """Queue image number i on lane i % lanes.  Returns that lane's plan (its outputs are valid once the lane's
        stream reaches this point; the next image on the same lane overwrites them)."""
        k = i % len(self.plans)
        with torch.cuda.stream(self.streams[k]):
            self.plans[k].forward(x_chw)
        return self.plans[k]

    def join(self):
        """The current stream waits for every lane."""
        cur = torch.cuda.current_stream()
        for st in self.streams:
            cur.wait_stream(st)


class StreamRunner(object):
    """Host-to-host streaming front end of a ForwardPlan: images come from (pinned) HOST memory and results go
    back to pinned HOST memory, for every image.  A ring of `depth` slots on two CUDA streams: the H2D copy of a
    later image runs on the copy stream while earlier images' graphs run on the compute stream; the D2H of
    (prob, boxes, count) is queued behind each graph.  Nothing is skipped per image -- the copies are overlapped,
    not removed -- so the steady-state rate is max(H2D, graph), not their sum.  The host only blocks when it
    reuses a slot (image i waits for image i-depth), so a stalled driver call (some hosts block cudaMemcpyAsync
    for tens of ms) does not drain the GPU queue."""

    def __init__(self, plan, src_hw=None, pixel_means=None, depth=8, lanes=1):
        """lanes: images in flight on the compute side (LanePool; lane = image index % lanes).
        src_hw=None: host images are the preprocessed (3,H,W) float32 tensors forward.py uploads.
        src_hw=(h0,w0): host images are RAW uint8 (h0,w0,3) BGR images; mean subtraction + bilinear resize run
        on the device (frcnn_preprocess_bgr8) -- 4x+ fewer H2D bytes."""
        self.pool = plan if isinstance(plan, LanePool) else LanePool(plan, lanes)
        self.plan = plan = self.pool.plans[0]
        dev = plan.x_in.device
        self.depth = D = max(2, int(depth), 2 * len(self.pool))
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.src_hw, self.pixel_means = src_hw, pixel_means
        if src_hw is None:
            self.stage = [torch.empty_like(plan.x_in) for _ in range(D)]
        else:
            from . import preprocess
            s, H, W = preprocess.plan_size(src_hw[0], src_hw[1])
            if (H, W) != (plan.H, plan.W):
                raise FrcnnError("source %s resizes to %s, but the plan is for %s" % (src_hw, (H, W), (plan.H, plan.W)))
            self.stage = [torch.empty((src_hw[0], src_hw[1], 3), dtype=torch.uint8, device=dev) for _ in range(D)]
        self.h2d_done = [torch.cuda.Event() for _ in range(D)]
        self.step_done = [torch.cuda.Event() for _ in range(D)]
        self.n_words = plan.result_words()
        self.res = [ops.PinnedBlock((self.n_words,), np.int32) for _ in range(D)]
        self.h2d_bytes = self.stage[0].numel() * self.stage[0].element_size()
        self.d2h_bytes = 4 * self.n_words

    def _deliver(self, i, counts, on_result):
        s = i % self.depth
        self.step_done[s].synchronize()
        counts.append(int(self.res[s].np[self.plan.result_layout["count"][0]]))
        if on_result is not None:
            on_result(i, self.plan.unpack_result(self.res[s].np))

    def run(self, host_images, on_result=None):
        """host_images: sequence of pinned host tensors ((3,H,W) float32, or (h0,w0,3) uint8 with src_hw).
        Calls on_result(i, res) in order with the unpacked result of image i (ForwardPlan.unpack_result: numpy views
        into the pinned block, valid until image i+depth is submitted).  Returns the proposal counts."""
        n, D, L = len(host_images), self.depth, len(self.pool)
        self.pool.fork()
        counts = []
        for i in range(n):
            s = i % D
            plan, cur = self.pool.plans[i % L], self.pool.streams[i % L]
            if i >= D:
                self._deliver(i - D, counts, on_result)         # slot reuse: image i-D must be finished and handed over
            # H2D on the copy stream (overlaps earlier images' graphs): cudaMemcpyAsync straight from the caller's buffer
            # -- asynchronous when that buffer is pinned (ops.PinnedBlock, tensor.pin_memory()), staged by the driver if not
            src = host_images[i]
            src = src.t if isinstance(src, ops.PinnedBlock) else (torch.from_numpy(src) if isinstance(src, np.ndarray) else src)
            if src.numel() * src.element_size() != self.h2d_bytes:
                raise FrcnnError("StreamRunner: host image %d has %d bytes, expected %d" %
                                 (i, src.numel() * src.element_size(), self.h2d_bytes))
            if not src.is_contiguous():
                src = src.contiguous()             # a strided view (e.g. an HWC->CHW transpose): densify on the host first
                host_images = list(host_images)
                host_images[i] = src               # keep it alive until the copy has been consumed
            ops.memcpy_h2d_async(self.stage[s], src.data_ptr(), self.h2d_bytes, self.copy_stream)
            self.h2d_done[s].record(self.copy_stream)
            with torch.cuda.stream(cur):
                cur.wait_event(self.h2d_done[s])
                if self.src_hw is None:
                    plan.x_in.copy_(self.stage[s], non_blocking=True)           # D2D into the graph's static input
                else:
                    from . import preprocess
                    if self.pixel_means is None:
                        preprocess.img_preprocessing(self.stage[s], out=plan.x_in)
                    else:
                        preprocess.img_preprocessing(self.stage[s], self.pixel_means, out=plan.x_in)
                plan.forward(None)
                # ONE D2H: count, class probabilities, boxes, proposals (+ the per-class NMS keep lists and counts)
                self.res[s].d2h(plan.result, cur, nbytes=4 * self.n_words)
                self.step_done[s].record(cur)
        for i in range(max(0, n - D), n):
            self._deliver(i, counts, on_result)
        self.pool.join()
        return counts



class _PlanLease(object):
    """Thread-local handle of a checked-out ForwardPlan: returns the plan to its engine's pool when the owning thread's
    thread-local storage is torn down."""

    def __init__(self, engine, key, plan):
        self.engine, self.key, self.plan = engine, key, plan

    def __del__(self):
        try:
            self.engine._return_plan(self.key, self.plan)
        except Exception:          # noqa: BLE001  (interpreter shutdown)
            pass


class Engine(object):
    """Weights + per-shape plans.  `engine(x)` -> (prob [R,21], boxes [R,84]) device tensors, R synced."""
    supports_hwc_input = True          # the VGG16 plan's first kernel takes source strides (ForwardPlan(hwc_input=True))

    def __init__(self, params, precision="bf16x3", device="cuda", anchors=None, num_classes=21, n_anchors=9,
                 feat_stride=16, **plan_kwargs):
        self.weights = PackedWeights(params, precision, device, num_classes, n_anchors)
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
            self.plans[key] = ForwardPlan(self.weights, H, W, anchors=self.anchors, feat_stride=self.feat_stride, **kw)
        return self.plans[key]

    def thread_plan(self, H, W, **overrides):
        """A ForwardPlan private to the calling thread (buffers, graph, stream over the shared weights): several host
        threads may call the model concurrently, one image each, and their graphs overlap on the GPU.  Plans live in a pool
        per (shape, options): a thread checks one out on its first call and a finalizer on its thread-local state hands it
        back when the thread ends, so the next thread reuses the captured graph and the pinned staging -- nothing is freed
        at thread exit (cudaFree / cudaFreeHost there would stall every other caller)."""
        kw = dict(self.plan_kwargs)
        kw.update(overrides)
        key = (H, W, tuple(sorted(kw.items())))
        held = self._tls.__dict__.setdefault("plans", {})
        lease = held.get(key)
        if lease is None:
            with self._lock:
                pool = self.__dict__.setdefault("_plan_pool", {}).setdefault(key, [])
                p = pool.pop() if pool else None
                if p is None:
                    base = self.plan(H, W, **overrides)
                    if not getattr(base, "_leased", False):
                        p = base
                    else:
                        p = base.clone()
                p._leased = True
            lease = held[key] = _PlanLease(self, key, p)
        return lease.plan

    def _return_plan(self, key, plan):
        with self._lock:
            self.__dict__.setdefault("_plan_pool", {}).setdefault(key, []).append(plan)

    def call_host(self, x_np, img_info=None, **overrides):
        """Host-array call: x_np float32 (3,H,W) numpy -> dict of numpy results (ForwardPlan.forward_host).  A C-contiguous
        array is uploaded as it is; so is the transposed view of a dense (H,W,3) array (forward.py:45 produces exactly that:
        `img.transpose([2, 0, 1]).astype(np.float32)` keeps the HWC memory) -- the first kernel then reads it with HWC
        strides; any other striding is densified on the host first."""
        H, W = int(x_np.shape[-2]), int(x_np.shape[-1])
        if x_np.dtype != np.float32:
            x_np = np.ascontiguousarray(x_np, dtype=np.float32)
        if x_np.flags.c_contiguous:
            dense = x_np
        elif self.supports_hwc_input and x_np.ndim == 3 and x_np.transpose(1, 2, 0).flags.c_contiguous:
            dense = x_np.transpose(1, 2, 0)             # the (H,W,3) memory itself, no copy
            overrides = dict(overrides, hwc_input=True)
        else:
            dense = np.ascontiguousarray(x_np)
        x_np = dense
        p = self.thread_plan(H, W, **overrides)
        if img_info is not None:
            p.set_clip(int(img_info[0]), int(img_info[1]))
        return p.forward_host(x_np), p

    def __call__(self, x_chw, img_info=None, **overrides):
        _, H, W = x_chw.shape
        p = self.plan(H, W, **overrides)
        if img_info is not None:
            p.set_clip(int(img_info[0]), int(img_info[1]))
        prob, boxes, count = p.forward(x_chw)
        R = int(count.item())           # the one documented D2H sync
        return prob[:R], boxes[:R], p


class CForward(object):
    """frcnn_forward_vgg16: the whole forward path behind the single C-ABI entry (include/frcnn_b200.h) -- what a caller
    without Python binds.  Same kernels, same order as ForwardPlan._run; this wrapper only fills the C structs from a
    PackedWeights and owns the workspace / output tensors."""

    def __init__(self, weights, H, W, anchors, pre_n=6000, post_n=300, nms_thresh=0.7, min_size=16, feat_stride=16):
        import ctypes
        from . import _lib
        self._lib, self._ct = _lib.load(), ctypes
        self.w, self.H, self.W, self.post_n = weights, H, W, post_n
        dev = weights.device
        self.anchors = torch.from_numpy(np.ascontiguousarray(anchors, dtype=np.float64)).to(dev)
        cfg = _lib.ForwardConfig(H, W, weights.num_classes, weights.n_anchors, feat_stride, pre_n, post_n, min_size,
                                 float(nms_thresh), 1 if weights.precision == "bf16x3" else 0)
        p = lambda t: None if t is None else t.data_ptr()                     # noqa: E731
        layer = lambda hi, lo, b: _lib.PackedLayer(p(hi), p(lo), p(b))         # noqa: E731
        wt = _lib.Vgg16Weights()
        for i, item in enumerate([it for it in VGG16_LAYERS if it != "pool"]):
            wt.conv[i] = layer(*weights.convs[item[0]])
        wt.rpn3, wt.rpn_heads = layer(*weights.rpn3), layer(*weights.rpn_heads)
        wt.fc6, wt.fc7, wt.head = layer(*weights.fc6), layer(*weights.fc7), layer(*weights.head)
        wt.anchors = self.anchors.data_ptr()
        self.cfg, self.wt = cfg, wt
        nbytes = self._lib.frcnn_forward_workspace_bytes(ctypes.byref(cfg))
        if nbytes == 0:
            raise FrcnnError("frcnn_forward_workspace_bytes: %s" % _lib.last_error())
        self.ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        self.prob = torch.empty((post_n, weights.num_classes), dtype=torch.float32, device=dev)
        self.boxes = torch.empty((post_n, 4 * weights.num_classes), dtype=torch.float32, device=dev)
        self.count = torch.zeros((1,), dtype=torch.int32, device=dev)

    def forward(self, x_chw, im_info=None):
        """x_chw (3,H,W) float32 CUDA.  Returns (prob, boxes, count) device tensors (rows past count are zero)."""
        from . import _lib
        im_h, im_w = (self.H, self.W) if im_info is None else (int(im_info[0]), int(im_info[1]))
        x = x_chw.contiguous().float()
        ct = self._ct
        _lib.check(self._lib.frcnn_forward_vgg16(ct.byref(self.cfg), ct.byref(self.wt), ct.c_void_p(x.data_ptr()), im_h, im_w,
                                                 ct.c_void_p(self.ws.data_ptr()), self.ws.numel(), ct.c_void_p(self.prob.data_ptr()),
                                                 ct.c_void_p(self.boxes.data_ptr()), ct.c_void_p(self.count.data_ptr()),
                                                 ct.c_void_p(torch.cuda.current_stream().cuda_stream)), "frcnn_forward_vgg16")
        return self.prob, self.boxes, self.count
