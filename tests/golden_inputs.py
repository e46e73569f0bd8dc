"""Seeded input generators shared by tests/golden/make_golden.py (which feeds them to the
REFERENCE to produce the committed golden outputs) and by the tests (which feed the same
inputs to the oracle and to the CUDA path).  Large inputs are regenerated from the seed
instead of being committed; `checksum` guards against generator drift."""
import zlib

import numpy as np

f32 = np.float32


def checksum(*arrays):
    c = 0
    for a in arrays:
        c = zlib.crc32(np.ascontiguousarray(a).tobytes(), c)
    return np.int64(c)


def _unique_f32(rng, draw, n):
    """n float32 values from draw(k), all distinct (tie-free scores)."""
    v = draw(n).astype(f32)
    for _ in range(100):
        _, first = np.unique(v, return_index=True)
        dup = np.setdiff1d(np.arange(n), first)
        if dup.size == 0:
            return v
        v[dup] = draw(dup.size).astype(f32)
    raise RuntimeError("could not make scores unique")


# --------------------------------------------------------------------------- bbox_transform
def box_transform_case(n, k, seed):
    rng = np.random.default_rng(seed)
    xy = rng.uniform(-50, 900, size=(n, 2))
    wh = rng.uniform(4, 400, size=(n, 2))
    boxes = np.hstack([xy, xy + wh]).astype(f32)
    trans = (rng.standard_normal((n, 4 * k)) * 0.6).astype(f32)
    return boxes, trans


# --------------------------------------------------------------------------- cpu_nms
def _clustered_dets(n, seed, ncl=40, W=1000.0, H=600.0):
    rng = np.random.default_rng(seed)
    c = rng.uniform([0, 0], [W, H], size=(ncl, 2))
    s = rng.uniform(30, 300, size=(ncl, 2))
    idx = rng.integers(0, ncl, size=n)
    ctr = c[idx] + rng.standard_normal((n, 2)) * 12
    wh = s[idx] * np.exp(rng.standard_normal((n, 2)) * 0.15)
    x1y1 = np.clip(ctr - wh / 2, 0, [W - 1, H - 1])
    x2y2 = np.clip(ctr + wh / 2, 0, [W - 1, H - 1])
    sc = _unique_f32(rng, lambda m: rng.uniform(0, 1, size=m), n)
    return np.hstack([x1y1, x2y2, sc[:, None]]).astype(f32)


def _integer_dets(n, seed, grid=24):
    """Integer-coordinate boxes on a small grid: exact rational IoUs such as 7/10, 3/10, 1/2
    are frequent, which is what separates `(double)ovr >= thr` (cpu_nms.pyx:66) from
    `ovr >= (float)thr` and from nms_kernel.cu's `>`."""
    rng = np.random.default_rng(seed)
    x1 = rng.integers(0, grid, size=(n, 2))
    wh = rng.integers(1, grid, size=(n, 2))
    sc = _unique_f32(rng, lambda m: rng.uniform(0, 1, size=m), n)
    d = np.hstack([x1, x1 + wh, sc[:, None]]).astype(f32)
    # hand-made exact cases first (highest scores so they are evaluated against each other):
    #   A=(0,0,9,9) area 100;  B=(0,0,9,6) area 70 -> IoU 7/10;  C=(0,0,9,2) area 30 -> 3/10 with A
    #   D=(0,0,9,4) area 50 -> 1/2 with A
    d[0] = [100, 100, 109, 109, 0.99]
    d[1] = [100, 100, 109, 106, 0.98]
    d[2] = [100, 100, 109, 102, 0.97]
    d[3] = [100, 100, 109, 104, 0.96]
    return d


NMS_CASES = {
    "n0": (0, 0.7), "n1": (1, 0.7), "n63": (63, 0.7), "n64": (64, 0.7), "n65": (65, 0.7),
    "n300_t03": (300, 0.3), "n2000_t07": (2000, 0.7), "n6000_t07": (6000, 0.7),
    "identical_t07": (100, 0.7), "disjoint_t03": (128, 0.3),
    "int_t07": (600, 0.7), "int_t03": (600, 0.3), "int_t05": (600, 0.5),
}


def nms_case(name):
    n, thr = NMS_CASES[name]
    seed = zlib.crc32(name.encode()) & 0xFFFF
    if name.startswith("identical"):
        rng = np.random.default_rng(seed)
        sc = _unique_f32(rng, lambda m: rng.uniform(0, 1, size=m), n)
        d = np.tile(np.array([[10, 20, 110, 220, 0]], dtype=f32), (n, 1))
        d[:, 4] = sc
        return d, thr
    if name.startswith("disjoint"):
        rng = np.random.default_rng(seed)
        i = np.arange(n)
        x1 = (i % 16) * 50.0
        y1 = (i // 16) * 50.0
        sc = _unique_f32(rng, lambda m: rng.uniform(0, 1, size=m), n)
        return np.stack([x1, y1, x1 + 30, y1 + 30, sc], axis=1).astype(f32), thr
    if name.startswith("int_"):
        return _integer_dets(n, seed), thr
    if n == 0:
        return np.zeros((0, 5), dtype=f32), thr
    return _clustered_dets(n, seed, ncl=max(1, min(40, n // 8 + 1))), thr


# --------------------------------------------------------------------------- ProposalLayer
PROPOSAL_CASES = {
    # name: (feat_h, feat_w, (img_h, img_w), train_mode, kind, seed)
    "t14_train": (14, 14, (224, 224), True, "uniform", 1),     # tests/test_proposal_layer.py:20-32
    "t14_test": (14, 14, (224, 224), False, "uniform", 2),
    "c0_test": (38, 50, (600, 600), False, "softmax", 3),      # forward.py:93 passes (H, H) (Q7)
    "c0w_test": (38, 50, (600, 800), False, "softmax", 4),
    "c1_test": (38, 63, (600, 1000), False, "softmax", 5),     # headline config
    "c1_train": (38, 63, (600, 1000), True, "softmax", 6),
    "c1_wide_test": (38, 63, (600, 1000), False, "wide", 7),   # large deltas: min-size filter bites
}


def proposal_case(name, A=9):
    fh, fw, info, train, kind, seed = PROPOSAL_CASES[name]
    rng = np.random.default_rng(1000 + seed)
    n = A * fh * fw
    if kind == "uniform":
        prob = rng.uniform(0, 1, size=(1, 2 * A, fh, fw)).astype(f32)
        fg = _unique_f32(rng, lambda m: rng.uniform(0, 1, size=m), n)
        pred = rng.uniform(0, 1, size=(1, 4 * A, fh, fw)).astype(f32)
    else:
        logits = rng.standard_normal((1, 2 * A, fh, fw)) * 1.5
        e = np.exp(logits - logits.max(axis=1, keepdims=True))
        prob = (e / e.sum(axis=1, keepdims=True)).astype(f32)
        fg0 = prob[0, A:].reshape(-1).copy()
        fg = _unique_f32(rng, lambda m: rng.uniform(1e-4, 0.6, size=m), n)
        # keep the softmax values where they are already unique, patch duplicates only
        _, first = np.unique(fg0, return_index=True)
        uniq_mask = np.zeros(n, dtype=bool)
        uniq_mask[first] = True
        fg = np.where(uniq_mask, fg0, fg).astype(f32)
        assert np.unique(fg).size == n
        std = 1.2 if kind == "wide" else 0.35
        pred = (rng.standard_normal((1, 4 * A, fh, fw)) * std).astype(f32)
    prob[0, A:] = fg.reshape(A, fh, fw)
    assert np.unique(prob[0, A:]).size == n
    return prob, pred, np.array([info], dtype=np.int32), train
