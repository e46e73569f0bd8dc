"""models.region_proposal_network.RegionProposalNetwork -- same class / constructor / call signature as
/root/reference models/region_proposal_network.py:20-204 for the inference branch (:117-124,158):

    h = relu(rpn_conv_3x3(x)); score = rpn_cls_score(h); prob = softmax(score) over ALL 2A channels (Q1)
    pred = rpn_bbox_pred(h); proposals, probs = proposal_layer(prob, pred, img_info)

On the B200: ONE tcgen05 3x3 conv (+ReLU), ONE 1x1 conv for the merged twin heads (2A+4A outputs,
fp32 NHWC), and frcnn_proposals with the softmax fused.

Training branch (:126-156,160-204; `train` and gt_boxes given): AnchorTargetLayer on device (frcnn_anchor_targets)
and both losses + the accuracy in one pass (frcnn_rpn_loss), which also emits d(rpn_loss)/d(head outputs) in the
merged [H*W, ld] layout (`self.head_grad`) -- the seed of the backward pass.  Returns rpn_loss like the reference and
keeps rpn_loss_cls / rpn_loss_bbox / rpn_cls_accuracy as attributes (the reference reports them to chainer.reporter).
"""
import os

import torch

from frcnn_b200 import arrays, links, ops, train_ops
from models.anchor_target_layer import AnchorTargetLayer
from models.proposal_layer import ProposalLayer


class RegionProposalNetwork(links.Link):
    type_check_enable = int(os.environ.get('CHAINER_TYPE_CHECK', '1')) != 0
    precision = "bf16x3"

    def __init__(self, in_ch=512, mid_ch=512, feat_stride=16, anchor_ratios=(0.5, 1, 2),
                 anchor_scales=(8, 16, 32), num_classes=21, loss_lambda=1., delta=3):
        super(RegionProposalNetwork, self).__init__()
        n_anchors = len(anchor_ratios) * len(anchor_scales)
        self.add_link("rpn_conv_3x3", links.conv_link(in_ch, mid_ch, 3, 0.01))
        self.add_link("rpn_cls_score", links.conv_link(mid_ch, 2 * n_anchors, 1, 0.01))
        self.add_link("rpn_bbox_pred", links.conv_link(mid_ch, 4 * n_anchors, 1, 0.01))
        d = self.__dict__
        d["proposal_layer"] = ProposalLayer(feat_stride, anchor_ratios, anchor_scales)
        d["anchor_target_layer"] = AnchorTargetLayer(feat_stride, anchor_ratios, anchor_scales)
        d["head_out"], d["head_grad"], d["head_ld"] = None, None, 0
        d["rpn_loss"], d["rpn_loss_cls"], d["rpn_loss_bbox"], d["rpn_cls_accuracy"] = None, None, None, None
        d["_loss_lambda"], d["_delta"], d["_feat_stride"] = loss_lambda, delta, feat_stride
        d["_packed"] = (None, -1)
        d["_train"] = True

    @property
    def train(self):
        return self._train

    @train.setter
    def train(self, val):
        self.__dict__["_train"] = val
        self.proposal_layer.train = val

    def _check_data_type_forward(self, x, img_info, gt_boxes):
        from chainer import Variable
        assert isinstance(x, Variable) and isinstance(img_info, Variable)
        assert x.shape[0] == 1 and arrays.dtype_kind(x) == 'f'
        assert img_info.shape == (1, 2) and arrays.dtype_kind(img_info) == 'i'
        if gt_boxes is not None:
            assert isinstance(gt_boxes, Variable)
            assert gt_boxes.shape[0] == 1 and gt_boxes.shape[2] == 5 and arrays.dtype_kind(gt_boxes) == 'f'

    def _weights(self, device):
        packed, ver = self._packed
        if packed is None or ver != self.version_key():
            T = lambda a: torch.from_numpy(a).to(device)
            hi3, lo3 = ops.pack_conv_weights(T(self.rpn_conv_3x3.W.data), precision=self.precision)
            b3 = ops.pad_bias(T(self.rpn_conv_3x3.b.data), self.rpn_conv_3x3.b.data.size)
            wh = torch.cat([T(self.rpn_cls_score.W.data), T(self.rpn_bbox_pred.W.data)], dim=0)
            bh = torch.cat([T(self.rpn_cls_score.b.data), T(self.rpn_bbox_pred.b.data)], dim=0)
            ld = ops.round_up(wh.shape[0], 32)
            hih, loh = ops.pack_conv_weights(wh, precision=self.precision)
            packed = dict(c3=(hi3, lo3, b3), heads=(hih, loh, ops.pad_bias(bh, ld)), ld=ld)
            self.__dict__["_packed"] = (packed, self.version_key())
        return packed

    def forward_device(self, feat, im_h, im_w):
        """feat: ops.Act [H,W,C].  Returns the ProposalWorkspace (rois/scores/count on device, no sync)."""
        pk = self._weights(feat.hi.device)
        pl = self.proposal_layer
        H, W, _ = feat.hi.shape
        self.__dict__["_feat_hw"] = (H, W)
        mid, _ = ops.conv2d(feat, pk["c3"][0], pk["c3"][1], pk["c3"][2], 3, True)
        _, y32 = ops.conv2d(mid, pk["heads"][0], pk["heads"][1], pk["heads"][2], 1, False, out_act=False, ld_f32=pk["ld"])
        self.__dict__["head_out"], self.__dict__["head_ld"] = y32, pk["ld"]
        if pl._anchors_dev is None or pl._anchors_dev.device != y32.device:
            import numpy as np
            pl._anchors_dev = torch.from_numpy(np.ascontiguousarray(pl._anchors, dtype=np.float64)).to(y32.device)
        pl._work = ops.proposals(y32, None, pl._anchors_dev, pl._num_anchors, H, W, self._feat_stride, im_h, im_w,
                                 pl._min_size, pl._pre_nms_top_n, pl._post_nms_top_n, pl._nms_thresh,
                                 layout="nhwc", ld=pk["ld"], cls_is_logits=True, work=pl._work)
        return pl._work

    def _train_losses(self, gt_dev, im_h, im_w, fam):
        """:126-156 -- targets, both losses, accuracy and the head gradient, all on the device."""
        from chainer import Variable
        pl, at = self.proposal_layer, self.anchor_target_layer
        H, W = self._feat_hw
        w = at.run_device(H, W, gt_dev, im_h, im_w)
        losses, dmat, _ = train_ops.rpn_loss(self.head_out, None, pl._anchors_dev, pl._num_anchors, H, W, self._feat_stride,
                                             im_h, im_w, w, delta=float(self._delta), loss_lambda=float(self._loss_lambda),
                                             layout="nhwc", ld=self.head_ld)
        d = self.__dict__
        d["head_grad"] = dmat
        if fam == arrays.NUMPY:
            vals = list(losses.cpu().numpy())             # the caller reads the loss (reporter / trainer): one sync
        else:
            vals = [arrays.from_device(losses[i].clone(), fam) for i in range(4)]
        d["rpn_loss_cls"], d["rpn_loss_bbox"], d["rpn_cls_accuracy"], d["rpn_loss"] = [Variable(v) for v in vals]
        for name in ("rpn_loss_cls", "rpn_loss_bbox", "rpn_loss"):
            d[name].name = name                           # :141,146,149
        return self.rpn_loss

    def __call__(self, x, img_info, gt_boxes=None):
        """x (1,C,H,W) feature map, img_info (1,2) -> (proposals (R,4), probs (R,1))."""
        if self.type_check_enable:
            self._check_data_type_forward(x, img_info, gt_boxes)
        fam = arrays.family(x)
        t = arrays.to_device(x)
        hw = arrays.to_host_ints(img_info)
        feat = ops.pack_image(t[0], c_pad=t.shape[1], precision=self.precision)
        work = self.forward_device(feat, int(hw[0]), int(hw[1]))
        if self.train and gt_boxes is not None:
            return self._train_losses(arrays.to_device(gt_boxes)[0], int(hw[0]), int(hw[1]), fam)
        R = int(work.count.item())
        return arrays.from_device(work.rois[:R].clone(), fam), arrays.from_device(work.scores[:R].reshape(R, 1).clone(), fam)
