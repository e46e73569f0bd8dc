"""models.bbox_transform -- same five public functions as the reference module
(/root/reference models/bbox_transform.py:18,41,79,102,112), array-family polymorphic like the
reference's `xp` code: numpy in -> numpy out, device array in -> device array out.  The three
functions on the forward path (bbox_transform_inv, clip_boxes, filter_boxes) run in
libfrcnn_b200.so (frcnn_bbox_decode) with bit-identical arithmetic to the fused kernels;
bbox_transform / keep_inside are training-side helpers ("next" rows) kept as plain array code.
"""
import torch

from frcnn_b200 import arrays, ops, train_ops


def bbox_transform_inv(boxes, trans):
    """Apply (dx, dy, dw, dh) deltas to boxes (reference :41-76).  boxes (N,4), trans (N,4K) -> (N,4K)."""
    fam = arrays.family(boxes)
    b, t = arrays.to_device(boxes), arrays.to_device(trans)
    if b.shape[0] == 0:
        return arrays.from_device(torch.zeros((0, t.shape[1]), dtype=torch.float32, device=b.device), fam)
    out, _ = ops.bbox_decode(b, t)
    return arrays.from_device(out, fam)


def clip_boxes(boxes, im_shape):
    """Clamp x to [0, W-1] and y to [0, H-1]; im_shape = (height, width) (reference :79-99).
    Like the reference the input array is updated in place and returned."""
    fam = arrays.family(boxes)
    hw = arrays.to_host_ints(im_shape)
    out, _ = ops.bbox_decode(arrays.to_device(boxes), None, clip_to=(int(hw[0]), int(hw[1])))
    res = arrays.from_device(out, fam)
    tgt = arrays.raw(boxes)
    if fam == arrays.NUMPY:
        tgt[...] = res
        return tgt
    (tgt.tensor if fam == arrays.DEVICE else tgt).copy_(out)
    return tgt


def filter_boxes(boxes, min_size):
    """Indices of boxes whose width and height (+1 convention) are both >= min_size (reference :102-109)."""
    fam = arrays.family(boxes)
    b = arrays.to_device(boxes)
    if b.shape[0] == 0:
        return arrays.from_device(torch.zeros((0,), dtype=torch.int64, device=b.device), fam)
    _, flags = ops.bbox_decode(b[:, :4], None, min_size=int(min_size))
    return arrays.from_device(torch.nonzero(flags, as_tuple=False).reshape(-1), fam)


def keep_inside(anchors, img_info):
    """Indices + rows of anchors lying fully inside the image (reference :112-130; training side): the predicate is
    frcnn_keep_inside, the index compaction is plumbing."""
    fam = arrays.family(anchors)
    a = arrays.to_device(anchors)
    hw = arrays.to_host_ints(img_info)
    flags = train_ops.keep_inside_flags(a[:, :4], int(hw[0]), int(hw[1]))
    idx = torch.nonzero(flags, as_tuple=False).reshape(-1)
    if fam == arrays.NUMPY:                     # rows come back in the caller's dtype (the reference keeps float64 anchors)
        idx_np = idx.cpu().numpy()
        return idx_np, arrays.raw(anchors)[idx_np]
    return arrays.from_device(idx, fam), arrays.from_device(a[idx], fam)


def bbox_transform(ex_rois, gt_rois):
    """Regression targets (dx, dy, dw, dh) of gt boxes w.r.t. example boxes (reference :18-38; training side), float32
    rows through frcnn_bbox_transform."""
    fam = arrays.family(ex_rois)
    e, g = arrays.to_device(ex_rois), arrays.to_device(gt_rois)
    return arrays.from_device(train_ops.bbox_transform(e[:, :4], g), fam)
