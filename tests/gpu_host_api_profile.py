"""Diagnostic (not a pytest): where does the time of the reference-interface call go?
    python tests/gpu_host_api_profile.py
Times, on the GPU box: the host copy pageable -> pinned (torch, multi-threaded), the H2D, the graph, the D2H + sync, the whole
FasterRCNN.__call__ on a host array (dense CHW and forward.py's HWC-strided view), and models.cpu_nms.cpu_nms on 300 host rows."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "chainer-faster-rcnn_b200"), os.path.join(ROOT, "oracle"), ROOT):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import frcnn_oracle as orc  # noqa: E402
import bench  # noqa: E402
from frcnn_b200 import ops  # noqa: E402


def med(fn, n=30, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(1e3 * (time.perf_counter() - t0))
    ts.sort()
    return ts[len(ts) // 2], ts[0], ts[-1]


params = orc.make_params(seed=1234)
model = bench.build_reference_api_model(params)
from chainer import Variable  # noqa: E402
from models.cpu_nms import cpu_nms  # noqa: E402
x = orc.make_image(600, 1000, seed=0)                    # forward.py:45's layout: CHW view of HWC memory
xc = np.ascontiguousarray(x)
info = Variable(np.array([[600, 1000]], np.int32))
print("torch threads", torch.get_num_threads(), "| x c_contiguous", x[0].flags.c_contiguous)
print("model(x) strided-HWC host array  : median %.3f ms (min %.3f max %.3f)" % med(lambda: model(Variable(x), info)))
print("model(x) dense CHW host array    : median %.3f ms (min %.3f max %.3f)" % med(lambda: model(Variable(xc), info)))
eng = model.engine()
plan = eng.thread_plan(600, 1000, pre_n=6000, post_n=300, nms_thresh=0.7, min_size=16)
io = plan.host_io()
src = torch.from_numpy(xc[0]).reshape(-1)
print("host copy pageable->pinned 7.2MB : median %.3f ms (min %.3f max %.3f)" % med(lambda: io["x"].t.view(-1).copy_(src)))
for nt in (1, 4, 8, 16):
    torch.set_num_threads(nt)
    print("  ... with %2d torch threads        : median %.3f ms (min %.3f max %.3f)" % ((nt,) + med(lambda: io["x"].t.view(-1).copy_(src))))
torch.set_num_threads(8)
dst_np = io["x"].np.reshape(-1)
src_np = xc[0].reshape(-1)
print("numpy copyto pageable->pinned    : median %.3f ms (min %.3f max %.3f)" % med(lambda: np.copyto(dst_np, src_np)))
st = io["stream"]


def h2d():
    io["x"].h2d(plan.x_in, st)
    ops.stream_synchronize(st)


def graph():
    with torch.cuda.stream(st):
        plan.forward(None)
    ops.stream_synchronize(st)


def d2h():
    io["res"].d2h(plan.result, st, nbytes=4 * plan.result_words())
    ops.stream_synchronize(st)
print("H2D + sync                       : median %.3f ms (min %.3f max %.3f)" % med(h2d))
print("graph replay + sync              : median %.3f ms (min %.3f max %.3f)" % med(graph))
print("D2H + sync                       : median %.3f ms (min %.3f max %.3f)" % med(d2h))
print("forward_host (all of the above)  : median %.3f ms (min %.3f max %.3f)" % med(lambda: plan.forward_host(xc[0])))
print("model(x) dense, 8 torch threads  : median %.3f ms (min %.3f max %.3f)" % med(lambda: model(Variable(xc), info)))
cls, box = model(Variable(xc), info)
dets = np.hstack((box[:, 4:8], cls.data[:, 1][:, np.newaxis]))
print("cpu_nms(300 host rows)           : median %.3f ms (min %.3f max %.3f)" % med(lambda: cpu_nms(dets, 0.3), n=200))
print("np.hstack of the caller          : median %.4f ms" % med(lambda: np.hstack((box[:, 4:8], cls.data[:, 1][:, np.newaxis])), n=200)[0])
d32 = np.ascontiguousarray(dets, np.float32)
print("ops.cpu_nms_host direct          : median %.3f ms (min %.3f max %.3f)" % med(lambda: ops.cpu_nms_host(d32, 0.3), n=200))
print("whole image, reference interface : median %.3f ms (min %.3f max %.3f)" %
      med(lambda: bench.reference_api_image(model, Variable(xc), info, cpu_nms, np)))
# where the per-call time of the host-array NMS goes: the round-trip floor (n = 2 rows: launch + mapped read + flag) and
# the kernel's own device time at n = 300 (CUPTI)
d2 = np.ascontiguousarray(d32[:2])
print("ops.cpu_nms_host, 2 rows (floor) : median %.3f ms (min %.3f max %.3f)" % med(lambda: ops.cpu_nms_host(d2, 0.3), n=200))
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(50):
        ops.cpu_nms_host(d32, 0.3)
    torch.cuda.synchronize()
for ev in prof.key_averages():
    if "nms" in ev.key:
        print("device time of %-40s: %.1f us avg over %d launches" % (ev.key[:40], ev.device_time_total / max(ev.count, 1), ev.count))
import ctypes  # noqa: E402
from frcnn_b200 import _lib  # noqa: E402

acc = np.zeros(5)
for _ in range(50):
    ops.cpu_nms_host(d32, 0.3)
    c = (ctypes.c_longlong * 8)()
    if _lib.load().frcnn_host_nms_phase_cycles(ctypes.cast(c, ctypes.c_void_p)):
        acc += np.diff(np.array(list(c)[:6], np.float64))
print("nms_small_fast_kernel phases, SM cycles avg (read rows | rank | bitmask | chain | write-out): " +
      " | ".join("%.0f" % v for v in acc / 50))
