"""models.cpu_nms -- same symbol as the reference's Cython extension (models/cpu_nms.pyx:18):

    keep = cpu_nms(dets, thresh)      # dets float32 (N,5) host array -> list[int]

forward.py:14,54 and models/proposal_layer.py:24,178 import exactly this.  The greedy suppression
runs on the GPU (frcnn_cpu_nms_host: bitmask kernel + device-side scan) with the Cython routine's
semantics -- float32 IoU with the +1 convention, `(double)iou >= thresh`, descending score order
(ties: lower index first).  There is no CPU implementation behind this name any more.
"""
import numpy as np

from frcnn_b200 import ops


def cpu_nms(dets, thresh):
    if not isinstance(dets, np.ndarray):
        dets = np.asarray(dets)
    if dets.ndim != 2 or dets.shape[1] != 5:
        raise ValueError("Buffer has wrong number of dimensions or columns (expected (N, 5), got %s)" % (dets.shape,))
    if dets.dtype != np.float32:
        raise ValueError("Buffer dtype mismatch, expected 'float32_t' but got '%s'" % dets.dtype)
    return ops.cpu_nms_host(dets, float(thresh))
