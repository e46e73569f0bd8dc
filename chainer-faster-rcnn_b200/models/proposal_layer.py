"""models.proposal_layer.ProposalLayer -- same class, constants, `train` switch and call signature as
/root/reference models/proposal_layer.py:30-221, with the whole body of __call__ (:126-198) replaced by
ONE stream-ordered C-ABI call, frcnn_proposals: anchor grid from the index, decode, clip, min-size
filter, fg-score slice, radix-select + chip-wide rank sort of the top-N, bitmask NMS with a device-side scan, top-N.
The reference's three host<->device round trips (:161-163, :176-177, :195-196) are gone; the only
synchronisation is reading the proposal count to size the returned arrays.
"""
import os

import numpy as np
import torch

from frcnn_b200 import arrays, ops
from models.cpu_nms import cpu_nms  # noqa: F401  (same import surface as the reference, :24)
from models.generate_anchors import generate_anchors
from models.gpu_nms import gpu_nms  # noqa: F401  (reference imports it at :27 without calling it)


class ProposalLayer(object):
    """RPN outputs -> object proposals at input-image scale.

    Args mirror the reference (:60-61): feat_stride, anchor_ratios, anchor_scales.
    """

    RPN_NMS_THRESH = 0.7
    TRAIN_RPN_PRE_NMS_TOP_N = 12000
    TRAIN_RPN_POST_NMS_TOP_N = 2000
    TEST_RPN_PRE_NMS_TOP_N = 6000
    TEST_RPN_POST_NMS_TOP_N = 300
    RPN_MIN_SIZE = 16

    type_check_enable = int(os.environ.get('CHAINER_TYPE_CHECK', '1')) != 0

    def __init__(self, feat_stride=16, anchor_ratios=(0.5, 1, 2), anchor_scales=(8, 16, 32)):
        self._feat_stride = feat_stride
        self._anchors = generate_anchors(ratios=anchor_ratios, scales=anchor_scales)
        self._num_anchors = len(self._anchors)
        self._nms_thresh = float(self.RPN_NMS_THRESH)
        self._min_size = self.RPN_MIN_SIZE
        self._anchors_dev = None
        self._work = None
        self.train = True          # reference default (:68-69): train-mode limits until told otherwise

    @property
    def train(self):
        return self._train

    @train.setter
    def train(self, value):
        self._train = value
        self._pre_nms_top_n = self.TRAIN_RPN_PRE_NMS_TOP_N if value else self.TEST_RPN_PRE_NMS_TOP_N
        self._post_nms_top_n = self.TRAIN_RPN_POST_NMS_TOP_N if value else self.TEST_RPN_POST_NMS_TOP_N

    # -- the reference's input contract (:85-100), same assertion style
    def _check_data_type_forward(self, rpn_cls_prob, rpn_bbox_pred, img_info):
        from chainer import Variable
        for v, mult in ((rpn_cls_prob, 2), (rpn_bbox_pred, 4)):
            assert isinstance(v, Variable)
            assert v.ndim == 4 and v.shape[0] == 1 and v.shape[1] == mult * self._num_anchors
            assert arrays.dtype_kind(v) == 'f'
        assert isinstance(img_info, Variable)
        assert img_info.shape == (1, 2)
        assert arrays.dtype_kind(img_info) == 'i'

    def __call__(self, rpn_cls_prob, rpn_bbox_pred, img_info):
        """rpn_cls_prob (1,2A,H,W), rpn_bbox_pred (1,4A,H,W), img_info (1,2) = (height, width).
        Returns (proposals (R,4), fg_probs (R,1)) in descending score order, R <= post_nms_top_n,
        in the array family of the inputs."""
        if self.type_check_enable:
            self._check_data_type_forward(rpn_cls_prob, rpn_bbox_pred, img_info)
        fam = arrays.family(rpn_cls_prob)
        prob = arrays.to_device(rpn_cls_prob)[0]
        pred = arrays.to_device(rpn_bbox_pred)[0]
        hw = arrays.to_host_ints(img_info)
        work = self.run_device(prob, pred, int(hw[0]), int(hw[1]))
        R = int(work.count.item())
        return arrays.from_device(work.rois[:R].clone(), fam), arrays.from_device(work.scores[:R].reshape(R, 1).clone(), fam)

    def run_device(self, prob_chw, pred_chw, im_h, im_w):
        """Device-only entry: CUDA tensors in, ops.ProposalWorkspace (rois / scores / count) out, no sync."""
        A = self._num_anchors
        _, H, W = pred_chw.shape
        if self._anchors_dev is None or self._anchors_dev.device != prob_chw.device:
            self._anchors_dev = torch.from_numpy(np.ascontiguousarray(self._anchors, dtype=np.float64)).to(prob_chw.device)
        # the reference treats a limit <= 0 as "take all" (models/proposal_layer.py:164-165,189-190); the device pipeline
        # sorts at most 16384 candidates and returns at most 2048 rows, so "all" maps to those limits
        pre_n = self._pre_nms_top_n if self._pre_nms_top_n > 0 else min(A * H * W, 16384)
        post_n = self._post_nms_top_n if self._post_nms_top_n > 0 else 2048
        if A * H * W > 16384 and self._pre_nms_top_n <= 0:
            raise ops.FrcnnError("ProposalLayer: 'no limit' pre_nms_top_n with %d anchors exceeds the device sort's 16384" % (A * H * W))
        self._work = ops.proposals(prob_chw, pred_chw, self._anchors_dev, A, H, W, self._feat_stride, im_h, im_w,
                                   self._min_size, pre_n, post_n, self._nms_thresh, layout="nchw", work=self._work)
        return self._work

    # -- kept for callers that ask for the explicit anchor grid (tests/test_anchor_target_layer.py:38)
    def _generate_all_bbox_use_array_info(self, rpn_bbox_pred):
        fam = arrays.family(rpn_bbox_pred)
        _, feat_h, feat_w = arrays.raw(rpn_bbox_pred).shape
        grid = self._generate_all_bbox(feat_h, feat_w).astype(np.float32)
        return grid if fam == arrays.NUMPY else arrays.from_device(torch.from_numpy(grid).cuda(), fam)

    def _generate_all_bbox(self, feat_h, feat_w):
        """(feat_h*feat_w*A, 4) float64 grid, row (y*feat_w + x)*A + a = anchor a shifted by the cell origin."""
        ys, xs = np.mgrid[0:feat_h, 0:feat_w]
        origin = np.stack([xs, ys, xs, ys], axis=-1).reshape(-1, 1, 4) * self._feat_stride
        return (origin + self._anchors[None, :, :]).reshape(-1, 4)
