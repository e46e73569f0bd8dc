"""Diagnostic (not a pytest): throughput and fairness of T caller threads driving the reference's interface
(models.faster_rcnn.FasterRCNN.__call__ on host arrays [+ the caller's 20 cpu_nms calls]) -- per-thread wall seconds."""
import os
import sys
import threading
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "chainer-faster-rcnn_b200"), os.path.join(ROOT, "oracle"), ROOT):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import frcnn_oracle as orc  # noqa: E402
import bench  # noqa: E402

params = orc.make_params(seed=1234)
model = bench.build_reference_api_model(params)
from chainer import Variable  # noqa: E402
from models.cpu_nms import cpu_nms  # noqa: E402
xs = [Variable(orc.make_image(600, 1000, seed=i)) for i in range(4)]
info = Variable(np.array([[600, 1000]], np.int32))
sys.setswitchinterval(float(os.environ.get("DIAG_SWITCH_INTERVAL", "1e-4")))
N = int(os.environ.get("DIAG_IMAGES", "48"))


def run(T, with_nms):
    secs = [0.0] * T
    bar = threading.Barrier(T + 1)

    def work(k):
        torch.cuda.set_device(0)
        for i in range(2):
            model(xs[(k + i) % 4], info)
        bar.wait()
        t0 = time.perf_counter()
        for i in range(N // T):
            if with_nms:
                bench.reference_api_image(model, xs[(k + i) % 4], info, cpu_nms, np)
            else:
                model(xs[(k + i) % 4], info)
        secs[k] = time.perf_counter() - t0
    ths = [threading.Thread(target=work, args=(k,)) for k in range(T)]
    for t in ths:
        t.start()
    bar.wait()
    t0 = time.perf_counter()
    for t in ths:
        t.join()
    tot = time.perf_counter() - t0
    return (N // T) * T / max(secs), (N // T) * T / tot, secs


if os.environ.get("DIAG_PHASES"):
    # where a caller thread's time goes under load: phases of forward_host (engine.HOST_PROFILE) + the 20 cpu_nms calls
    from frcnn_b200 import engine as _eng
    _raw_nms = cpu_nms
    _nms_ms = []
    _nms_host = []
    import ctypes
    from frcnn_b200 import _lib as _libmod
    _L = _libmod.load()

    def cpu_nms(dets, thresh):  # noqa: F811
        t0 = time.perf_counter()
        r = _raw_nms(dets, thresh)
        _nms_ms.append(1e3 * (time.perf_counter() - t0))
        c = (ctypes.c_longlong * 8)()
        if _L.frcnn_host_nms_phase_cycles(ctypes.cast(c, ctypes.c_void_p)):
            _nms_host.append((c[6] * 1e-3, c[7] * 1e-3, (c[5] - c[0])))
        return r

    for T in [int(v) for v in sys.argv[1].split(",")]:
        run(T, True)
        _eng.HOST_PROFILE = []
        del _nms_ms[:]
        del _nms_host[:]
        t0 = time.perf_counter()
        ips, _, secs = run(T, True)
        ph = np.array(_eng.HOST_PROFILE)
        _eng.HOST_PROFILE = None
        nm = np.array(_nms_ms)
        per_img = 1e3 * np.mean(secs) / (N // T)
        print("T=%d: %.1f img/s | per image per thread %.3f ms = upload %.3f + graph enqueue %.3f + D2H/wait %.3f + 20 x cpu_nms %.3f "
              "(mean call %.1f us, p50 %.1f, p90 %.1f, max %.1f) + python rest %.3f" %
              (T, ips, per_img, ph[:, 1].mean(), ph[:, 2].mean(), ph[:, 3].mean(), 20 * nm.mean(), 1e3 * nm.mean(),
               1e3 * np.median(nm), 1e3 * np.percentile(nm, 90), 1e3 * nm.max(),
               per_img - ph[:, 1:4].sum(1).mean() - 20 * nm.mean()), flush=True)
        hh = np.array(_nms_host)
        print("      inside cpu_nms: launch call mean %.1f us (p90 %.1f), flag poll mean %.1f us (p90 %.1f), kernel entry->done %.0f SM cycles" %
              (hh[:, 0].mean(), np.percentile(hh[:, 0], 90), hh[:, 1].mean(), np.percentile(hh[:, 1], 90), hh[:, 2].mean()), flush=True)
    sys.exit(0)

T_LIST = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 2, 3, 4, 6]
NMS_LIST = [bool(int(v)) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [False, True]
for with_nms in NMS_LIST:
    for T in T_LIST:
        for rep in range(3):
            ips, ips_wall, secs = run(T, with_nms)
            print("switch %g nms=%d T=%d rep %d: %.1f img/s over the threads' loops (%.1f incl. thread start/stop)  thread seconds %s" %
                  (sys.getswitchinterval(), with_nms, T, rep, ips, ips_wall, [round(s, 3) for s in secs]), flush=True)
