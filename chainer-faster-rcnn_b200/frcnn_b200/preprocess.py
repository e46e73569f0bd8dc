"""Caller-side pre/post-processing of the reference's forward.py on the device ("next" row).

  img_preprocessing(orig_img, pixel_means, max_size=1000, scale=600)   forward.py:34-45, same signature/returns
  detections(cls_prob, pred_boxes, im_scale, nms_thresh=0.3, conf=0.8)  the numeric part of draw_result, forward.py:48-59

The raw uint8 BGR image is uploaded (4x+ fewer H2D bytes than the float32 tensor forward.py uploads) and mean
subtraction + OpenCV-compatible bilinear resize + HWC->CHW run in frcnn_preprocess_bgr8; per-class NMS + confidence
filter run in frcnn_detect.  Drawing (cv.rectangle / putText) stays out of scope.
"""
import numpy as np
import torch

from . import _lib, ops
from ._lib import check

PIXEL_MEANS = np.array([[[102.9801, 115.9465, 122.7717]]])     # forward.py:22 (BGR)


def plan_size(h0, w0, scale=600, max_size=1000):
    """(im_scale, H, W): forward.py:38-41 + cv.resize's output-size rounding (half to even)."""
    im_scale = float(scale) / float(min(h0, w0))
    if np.round(im_scale * max(h0, w0)) > max_size:
        im_scale = float(max_size) / float(max(h0, w0))
    return im_scale, int(np.rint(h0 * im_scale)), int(np.rint(w0 * im_scale))


def img_preprocessing(orig_img, pixel_means=PIXEL_MEANS, max_size=1000, scale=600, out=None):
    """uint8 (h0,w0,3) BGR image (numpy or CUDA tensor) -> ((3,H,W) float32 CUDA tensor, im_scale)."""
    if isinstance(orig_img, np.ndarray):
        if orig_img.dtype != np.uint8 or orig_img.ndim != 3 or orig_img.shape[2] != 3:
            raise _lib.FrcnnError("img_preprocessing expects a uint8 (h, w, 3) BGR image")
        img = torch.from_numpy(np.ascontiguousarray(orig_img)).cuda(non_blocking=True)
    else:
        img = orig_img.contiguous()
        if img.dtype != torch.uint8 or not img.is_cuda:
            raise _lib.FrcnnError("img_preprocessing expects a uint8 image (numpy, or a CUDA tensor)")
    h0, w0, _ = img.shape
    im_scale, H, W = plan_size(h0, w0, scale, max_size)
    if out is None:
        out = torch.empty((3, H, W), dtype=torch.float32, device=img.device)
    m = np.asarray(pixel_means, dtype=np.float64).reshape(-1)
    check(_lib.load().frcnn_preprocess_bgr8(ops._p(img), h0, w0, float(m[0]), float(m[1]), float(m[2]), float(im_scale),
                                            H, W, ops._p(out), ops._stream()), "frcnn_preprocess_bgr8")
    return out, im_scale


def detections(cls_prob, pred_boxes, im_scale, nms_thresh=0.3, conf=0.8, count=None):
    """forward.py:48-59 without the drawing: for every foreground class, per-class NMS (device), confidence
    filter, boxes divided by im_scale and truncated to int like `map(int, dets[i, :4] / im_scale)`.
    Returns a list of (cls_id, x1, y1, x2, y2, score) in the reference's iteration order."""
    keep_idx, keep_count, conf_count = ops.detect(cls_prob, pred_boxes, count, nms_thresh, conf)
    keep_idx, conf_count = keep_idx.cpu().numpy(), conf_count.cpu().numpy()
    prob, boxes = cls_prob.detach().cpu().numpy(), pred_boxes.detach().cpu().numpy()
    out = []
    for c in range(1, prob.shape[1]):
        for r in keep_idx[c - 1, :conf_count[c - 1]]:
            x1, y1, x2, y2 = map(int, boxes[r, 4 * c:4 * c + 4] / im_scale)
            out.append((c, x1, y1, x2, y2, float(prob[r, c])))
    return out
