"""models.generate_anchors -- same public function as /root/reference models/generate_anchors.py:47.

Host-side constant table (float64, A x 4), computed once per layer construction exactly as the
reference does (it is not on the per-image path); the per-image anchor GRID is never materialised
here -- frcnn_proposals derives each anchor from its index on the device.
"""
import numpy as np


def generate_anchors(base_size=15, ratios=(0.5, 1, 2), scales=(4, 8, 16, 32)):
    """Anchor windows for every (ratio, scale) pair around the [0, 0, base_size, base_size] box.

    Closed form of the reference's enumeration (generate_anchors.py:47-93): for ratio r the window
    keeps the base area with sides (rint(sqrt(area/r)), rint(that*r)); each scale multiplies both
    sides; all boxes share the base window's centre.  Returns float64 [len(ratios)*len(scales), 4].
    """
    ratios = np.asarray(ratios, dtype=np.float64).reshape(-1, 1)
    scales = np.asarray(scales, dtype=np.float64).reshape(1, -1)
    side = float(base_size) + 1.0
    centre = 0.5 * (side - 1.0)
    w0 = np.rint(np.sqrt(side * side / ratios))
    h0 = np.rint(w0 * ratios)
    half_w = 0.5 * ((w0 * scales).reshape(-1) - 1.0)
    half_h = 0.5 * ((h0 * scales).reshape(-1) - 1.0)
    return np.stack([centre - half_w, centre - half_h, centre + half_w, centre + half_h], axis=1)
