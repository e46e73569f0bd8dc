"""models.gpu_nms -- same symbol as the reference's Cython wrapper (models/gpu_nms.pyx:16-31) around
`_nms` (models/gpu_nms.hpp:9-10).  Imported by models/proposal_layer.py:27 (never called there, Q2).

    keep = gpu_nms(dets, thresh, device_id=0)   # dets float32 (N,5) host array -> list[int]

Host side mirrors the wrapper: descending sort by score, call the C-ABI `_nms` on the sorted boxes,
map the kept positions back.  `_nms` keeps the CUDA kernel's `>` comparison (nms_kernel.cu:71).
"""
import numpy as np

from frcnn_b200 import ops


def gpu_nms(dets, thresh, device_id=0):
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    if dets.ndim != 2 or dets.shape[1] < 5:
        raise ValueError("gpu_nms expects (N, >=5) float32 dets")
    if dets.shape[0] == 0:
        return []
    order = np.argsort(-dets[:, 4], kind="stable")
    kept = ops.gpu_nms_host(dets[order], float(thresh), int(device_id))
    return [int(v) for v in order[kept]]
