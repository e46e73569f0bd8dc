"""models.proposal_target_layer.ProposalTargetLayer -- same class, constants and call signature as
/root/reference models/proposal_target_layer.py:25-150.

    use_gt_boxes, bbox_reg_targets, keep_inds = layer(proposals, gt_boxes)

The IoU of every proposal with the ground truth (the reference copies both to the HOST for this, anchor_target_layer.py
:179-187) and the matched rows / float32 regression targets / class-wise scatter (:138-147) run on the device
(frcnn_roi_overlaps, frcnn_roi_targets); the fg/bg sampling (:99-129) draws from NumPy's global RNG on the host exactly as
the reference does (same pools, same call order), so `np.random.seed(s)` reproduces a reference run bit for bit.
"""
import os

import numpy as np
import torch

from frcnn_b200 import arrays, train_ops
from models.anchor_target_layer import AnchorTargetLayer


class ProposalTargetLayer(AnchorTargetLayer):
    FG_THRESH = 0.5
    BG_THRESH_HI = 0.5
    BG_THRESH_LO = 0.1
    ROIS_PER_IMAGE = 128
    FG_FRACTION = 0.25

    type_check_enable = int(os.environ.get('CHAINER_TYPE_CHECK', '1')) != 0

    def __init__(self, feat_stride=16, anchor_ratios=(0.5, 1, 2), anchor_scales=(8, 16, 32), num_classes=21):
        super(ProposalTargetLayer, self).__init__(feat_stride, anchor_ratios, anchor_scales)
        self._num_classes = num_classes
        self._n_fg_rois = int(self.FG_FRACTION * self.ROIS_PER_IMAGE)

    def _check_data_type_forward(self, proposals, gt_boxes):
        from chainer import Variable
        assert len(proposals) > 0
        assert proposals.ndim == 2
        assert proposals.shape[1] == 4
        assert arrays.dtype_kind(proposals) == 'f'
        assert not isinstance(proposals, Variable)            # plain ndarray (numpy or device), :66
        assert isinstance(gt_boxes, Variable)
        assert gt_boxes.ndim == 3
        assert gt_boxes.shape[0] == 1
        assert gt_boxes.shape[2] == 5
        assert arrays.dtype_kind(gt_boxes) == 'f'

    def __call__(self, proposals, gt_boxes):
        if self.type_check_enable:
            self._check_data_type_forward(proposals, gt_boxes)
        fam = arrays.family(proposals)
        rois = arrays.to_device(proposals)
        gt = arrays.to_device(gt_boxes)[0]
        max_ov, argmax = train_ops.roi_overlaps(rois, None, gt)
        mo = max_ov.cpu().numpy()
        fg_inds = np.where(mo >= self.FG_THRESH)[0]                                         # :99
        n_fg = min(self._n_fg_rois, fg_inds.size)                                           # :103
        if fg_inds.size > 0:
            fg_inds = np.random.choice(fg_inds, size=n_fg, replace=False)                   # :105-110
        bg_inds = np.where((mo < self.BG_THRESH_HI) & (mo >= self.BG_THRESH_LO))[0]         # :113-114
        n_bg = min(self.ROIS_PER_IMAGE - n_fg, bg_inds.size)                                # :116-117
        if bg_inds.size > 0:
            bg_inds = np.random.choice(bg_inds, size=n_bg, replace=False)                   # :119-126
        keep = np.concatenate([fg_inds, bg_inds]).astype(np.int32)                          # :129
        keep_dev = torch.from_numpy(keep).to(rois.device)
        use_gt, ext, _ = train_ops.roi_targets(rois, gt, argmax, keep_dev, self._num_classes)
        return arrays.from_device(use_gt, fam), arrays.from_device(ext, fam), arrays.from_device(keep_dev, fam)
