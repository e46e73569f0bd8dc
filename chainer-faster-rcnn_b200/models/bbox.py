"""models.bbox.bbox_overlaps -- the import surface of /root/reference models/bbox.pyx:16-56 (a Cython extension there),
backed by frcnn_bbox_overlaps: host float64 arrays in, host float64 (N,K) matrix out, arithmetic on the GPU in the
reference's operation order (bit-identical results)."""
import numpy as np
import torch

from frcnn_b200 import train_ops


def bbox_overlaps(boxes, query_boxes):
    b = torch.from_numpy(np.ascontiguousarray(boxes, dtype=np.float64)).cuda()
    q = torch.from_numpy(np.ascontiguousarray(query_boxes, dtype=np.float64)).cuda()
    return train_ops.bbox_overlaps(b, q).cpu().numpy()
