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
sys.setswitchinterval(1e-4)
N = 48


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


T_LIST = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 2, 3, 4, 6]
NMS_LIST = [bool(int(v)) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [False, True]
for with_nms in NMS_LIST:
    for T in T_LIST:
        for rep in range(3):
            ips, ips_wall, secs = run(T, with_nms)
            print("nms=%d T=%d rep %d: %.1f img/s over the threads' loops (%.1f incl. thread start/stop)  thread seconds %s" %
                  (with_nms, T, rep, ips, ips_wall, [round(s, 3) for s in secs]), flush=True)
