"""models.cpu_nms -- same symbol as the reference's Cython extension (models/cpu_nms.pyx:18):

    keep = cpu_nms(dets, thresh)      # dets float32 (N,5) host array -> list[int]

forward.py:14,54 and models/proposal_layer.py:24,178 import exactly this.  The greedy suppression
runs on the GPU with the Cython routine's semantics -- float32 IoU with the +1 convention,
`(double)iou >= thresh`, descending score order (ties: lower index first).  There is no CPU
implementation behind this name any more.  Two device paths:

* standalone (frcnn_cpu_nms_host): the rows go to the GPU, one single-CTA kernel, the keep list comes back;
* in-graph hand-off: forward.py:48-57 calls this function once per class on
  `hstack(bbox_pred[:, 4c:4c+4], cls_score[:, c])` right after `model(x, img_info)`.  A model built with
  `caller_nms_thresh` (default 0.3, forward.py:75) already ran exactly that NMS for all 20 classes inside the
  image's CUDA graph (frcnn_detect) and brought the keep lists back in the same download as the scores.  If `dets`
  equals the corresponding columns of that result BIT FOR BIT (frcnn_match_class_dets) and `thresh` is the graph's,
  the device-computed keep list is returned; anything else -- other rows, other threshold, a model call in between
  on this thread -- takes the standalone path.  Both paths are bit-identical to the oracle's cpu_nms
  (tests/test_parity_hardening_gpu.py); the hand-off saves 20 kernel launches + round trips per image.
  `models.cpu_nms.HANDOFF = False` turns it off process-wide.
"""
import threading

import numpy as np

from frcnn_b200 import _lib, ops

HANDOFF = True
_tls = threading.local()
stats = {"handoff": 0, "standalone": 0}          # not synchronised: a diagnostic, not a counter to rely on


def publish(prob, boxes, keep_idx, keep_count, thresh):
    """Called by models.faster_rcnn.FasterRCNN.__call__ (host path) with VIEWS of the calling thread's result block: prob
    [R,NC] and boxes [R,4NC] float32, keep_idx [NC-1, post_n] int32 and keep_count [NC-1] from the in-graph per-class NMS at
    `thresh`.  The views stay untouched until this thread's next model call, which publishes anew."""
    _tls.rec = (prob, boxes, keep_idx, keep_count, float(thresh), prob.shape[0], prob.shape[1])
    _tls.hint = 1


def withdraw():
    _tls.rec = None


def cpu_nms(dets, thresh):
    if not isinstance(dets, np.ndarray):
        dets = np.asarray(dets)
    if dets.ndim != 2 or dets.shape[1] != 5:
        raise ValueError("Buffer has wrong number of dimensions or columns (expected (N, 5), got %s)" % (dets.shape,))
    if dets.dtype != np.float32:
        raise ValueError("Buffer dtype mismatch, expected 'float32_t' but got '%s'" % dets.dtype)
    rec = getattr(_tls, "rec", None)
    if rec is not None and HANDOFF and thresh == rec[4] and dets.shape[0] == rec[5] and dets.flags.c_contiguous:
        prob, boxes, keep_idx, keep_count, _, R, NC = rec
        c = _lib.load_gil().frcnn_match_class_dets(dets.ctypes.data, R, boxes.ctypes.data, boxes.strides[0] // 4, prob.ctypes.data,
                                                   prob.strides[0] // 4, NC, _tls.hint)
        if c > 0:
            _tls.hint = c + 1 if c + 1 < NC else 1
            stats["handoff"] += 1
            return keep_idx[c - 1, :int(keep_count[c - 1])].tolist()
    stats["standalone"] += 1
    return ops.cpu_nms_host(dets, float(thresh))
