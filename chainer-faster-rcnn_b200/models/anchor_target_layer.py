"""models.anchor_target_layer.AnchorTargetLayer -- same class, constants and call signature as
/root/reference models/anchor_target_layer.py:28-198, with the body of __call__ replaced by ONE stream-ordered C-ABI
call, frcnn_anchor_targets: float64 anchor grid + inside filter + IoU against the ground truth (the reference copies
anchors and gt to the HOST for this, :179-187) + labelling rules + regression targets.

Subsampling (:148-168).  `subsample = "numpy"` (default) keeps the reference's behaviour bit for bit: the positive /
negative index lists come back to the host and `np.random.choice(..., replace=False)` draws from NumPy's global RNG in
the reference's call order, so `np.random.seed(s)` reproduces a reference run exactly.  `subsample = "device"` selects
by a counter hash of (seed, anchor index) on the GPU with no host round trip (same guarantees on the counts, different
random stream); `seed` advances by one per call.
"""
import os

import numpy as np
import torch

from frcnn_b200 import arrays, train_ops
from models.proposal_layer import ProposalLayer


class AnchorTargetLayer(ProposalLayer):
    RPN_NEGATIVE_OVERLAP = 0.3
    RPN_POSITIVE_OVERLAP = 0.7
    RPN_FG_FRACTION = 0.5
    RPN_BATCHSIZE = 256

    type_check_enable = int(os.environ.get('CHAINER_TYPE_CHECK', '1')) != 0
    subsample = "numpy"
    seed = 0

    def __init__(self, feat_stride=16, anchor_ratios=(0.5, 1, 2), anchor_scales=(8, 16, 32)):
        super(AnchorTargetLayer, self).__init__(feat_stride, anchor_ratios, anchor_scales)
        self._targets = None

    def _check_data_type_forward(self, gt_boxes, img_info):
        from chainer import Variable
        assert isinstance(gt_boxes, Variable)
        assert gt_boxes.shape[0] == 1
        assert gt_boxes.shape[2] == 5
        assert arrays.dtype_kind(gt_boxes) == 'f'
        assert isinstance(img_info, Variable)
        assert img_info.shape == (1, 2)
        assert arrays.dtype_kind(img_info) == 'i'

    def run_device(self, feat_h, feat_w, gt_dev, im_h, im_w):
        """Device-only entry: gt_dev [G,5] float32 CUDA -> train_ops.AnchorTargets with the FINAL labels (after
        subsampling) in labels_full, ready for frcnn_rpn_loss.  "device" mode does not synchronise."""
        if self._anchors_dev is None or self._anchors_dev.device != gt_dev.device:
            self._anchors_dev = torch.from_numpy(np.ascontiguousarray(self._anchors, dtype=np.float64)).to(gt_dev.device)
        num_fg = int(self.RPN_FG_FRACTION * self.RPN_BATCHSIZE)
        kw = dict(work=self._targets, neg_thr=self.RPN_NEGATIVE_OVERLAP, pos_thr=self.RPN_POSITIVE_OVERLAP,
                  batch=self.RPN_BATCHSIZE, num_fg=num_fg)
        args = (self._anchors_dev, self._num_anchors, feat_h, feat_w, self._feat_stride, gt_dev, im_h, im_w)
        if self.subsample == "device":
            self._targets = train_ops.anchor_targets(*args, mode=train_ops.SUBSAMPLE_DEVICE, seed=self.seed, **kw)
            self.seed += 1
            return self._targets
        if self.subsample != "numpy":
            raise ValueError("AnchorTargetLayer.subsample must be 'numpy' or 'device'")
        w = self._targets = train_ops.anchor_targets(*args, mode=train_ops.SUBSAMPLE_NONE, **kw)
        counts = w.counts.cpu().numpy()
        n_inside, fg_before, bg_before = int(counts[0]), int(counts[3]), int(counts[4])
        disable = []
        if fg_before > num_fg or bg_before > self.RPN_BATCHSIZE - min(fg_before, num_fg):
            labels = w.labels_full[w.inds_inside[:n_inside].long()].cpu().numpy()       # inside-compact, like the reference
            fg_inds = np.where(labels == 1)[0]
            if len(fg_inds) > num_fg:                                                     # :151-157
                d = np.random.choice(fg_inds, size=int(len(fg_inds) - num_fg), replace=False)
                labels[d] = -1
                disable.append(d)
            num_bg = self.RPN_BATCHSIZE - np.sum(labels == 1)                             # :160
            bg_inds = np.where(labels == 0)[0]
            if len(bg_inds) > num_bg:                                                     # :162-168
                d = np.random.choice(bg_inds, size=int(len(bg_inds) - num_bg), replace=False)
                disable.append(d)
        if disable:
            pos = torch.from_numpy(np.concatenate(disable).astype(np.int64)).to(gt_dev.device)
            w.labels_full[w.inds_inside[:n_inside].long()[pos]] = -1
            fg_after = min(fg_before, num_fg)
            w.counts[1] = fg_after
            w.counts[2] = min(bg_before, self.RPN_BATCHSIZE - fg_after)
        return w

    def __call__(self, feat_h, feat_w, gt_boxes, img_info):
        """-> (bbox_labels [n_inside] int32 in {-1,0,1}, bbox_reg_targets [n_inside,4] float32,
        inds_inside [n_inside], n_all_bbox) in the array family of gt_boxes (reference :120)."""
        if self.type_check_enable:
            self._check_data_type_forward(gt_boxes, img_info)
        fam = arrays.family(gt_boxes)
        gt = arrays.to_device(gt_boxes)[0]
        hw = arrays.to_host_ints(img_info)
        w = self.run_device(int(feat_h), int(feat_w), gt, int(hw[0]), int(hw[1]))
        labels, targets, inds, n_all = w.compact()
        return (arrays.from_device(labels.clone(), fam), arrays.from_device(targets.clone(), fam),
                arrays.from_device(inds.clone(), fam), n_all)
