"""Diagnostic (not a pytest file): per-phase clock64 deltas of topk_sort_kernel + event timings of the 4 proposal kernels."""
import ctypes, sys
import numpy as np, torch
sys.path[:0] = ["chainer-faster-rcnn_b200", "oracle", "tests"]
import frcnn_oracle as orc, golden_inputs as gi
from frcnn_b200 import ops, _lib
prob, pred, info, train = gi.proposal_case("c1_test")
anchors = torch.from_numpy(orc.generate_anchors(ratios=(0.5, 1, 2), scales=(8, 16, 32))).cuda()
p, d = torch.from_numpy(prob[0]).cuda(), torch.from_numpy(pred[0]).cuda()
clocks = torch.zeros(8, dtype=torch.int64, device="cuda")
lib = _lib.load()
lib.frcnn_debug_sort_clocks(ctypes.c_void_p(clocks.data_ptr()))
work = None
for _ in range(3):
    work = ops.proposals(p, d, anchors, 9, 38, 63, 16, 600, 1000, 16, 6000, 300, 0.7, work=work)
torch.cuda.synchronize()
c = clocks.cpu().numpy()
names = ["count+cache", "radix select", "compaction", "bitonic", "gather"]
print("sort phases (cycles):", {n: int(c[i + 1] - c[i]) for i, n in enumerate(names)}, "total", int(c[5] - c[0]))
lib.frcnn_debug_sort_clocks(None)
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
e[0].record()
for _ in range(20):
    ops.proposals(p, d, anchors, 9, 38, 63, 16, 600, 1000, 16, 6000, 300, 0.7, work=work)
e[1].record(); torch.cuda.synchronize()
print("whole frcnn_proposals: %.1f us" % (e[0].elapsed_time(e[1]) / 20 * 1e3), "R =", int(work.count.item()))
