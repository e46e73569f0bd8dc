#!/usr/bin/env python
"""bench.py -- images/sec of the end-to-end VGG16 Faster R-CNN forward path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16x3|bf16] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config #2): synthetic 600x1000 image (uniform(0,255) - BGR means), random-init weights
(heads N(0,0.01), He-normal trunk), test-mode ProposalLayer (6000 -> NMS 0.7 -> 300), 21 classes.
One step = one image through the whole graph (trunk, RPN, ProposalLayer, RoI pool, fc6/fc7, heads,
softmax/decode/clip) on each GPU; images shard one per GPU with no collective ("scaling": "weak").

Prints ONE JSON line (rank 0).  `value`   : device-timed (CUDA events) images/s, inputs resident in HBM.
                                `e2e`     : same metric through the public call with a pinned HOST image copied
                                            H2D and the result copied D2H inside the timed region, every step.
                                `roofline`: the conv/GEMM tensor-core kernel, timed live per launch.
                                `cpu_baseline`: the CPU oracle pipeline on this box's host cores (rank 0, N=1).
`--impl reference` times the reference-equivalent CPU pipeline instead (torch-CPU fp32 dense ops standing in for
Chainer-NumPy -- Chainer is not installable offline -- plus the reference's own compiled cpu_nms.pyx when
oracle/_ref holds it, else the C restatement).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "chainer-faster-rcnn_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

H_IMG, W_IMG = 600, 1000
CONV_STACK_GFLOP = 379.17       # trunk 367.74 + RPN 3x3 11.30 + RPN 1x1 0.13
WHOLE_GFLOP = 451.15


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1421.6), d.get("hbm_gbs", 6571.9), "measured"
    return 1590.0, 6650.0, "fallback"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------- CPU (reference-equivalent) arm
def cpu_pipeline_once(orc, params, x, info, ref_nms=None, raw=None):
    """One image through the reference-equivalent CPU path.  Returns (seconds, nms_seconds).
    raw: optional uint8 (h0,w0,3) BGR image -- then forward.py's img_preprocessing (34-45) is part of the timed work
    and replaces `x`, like the B200 arm's e2e."""
    t0 = time.perf_counter()
    if raw is not None:
        x = orc.img_preprocessing(raw)[0][None]
    if ref_nms is not None:
        saved = orc.cpu_nms
        tn = [0.0]

        def timed_nms(dets, thr):
            s = time.perf_counter()
            k = [int(v) for v in ref_nms.cpu_nms(np.ascontiguousarray(dets, dtype=np.float32), thr)]
            tn[0] += time.perf_counter() - s
            return k
        orc.cpu_nms = timed_nms
    else:
        tn = [0.0]
        saved = orc.cpu_nms

        def timed_nms2(dets, thr):
            s = time.perf_counter()
            k = saved(dets, thr)
            tn[0] += time.perf_counter() - s
            return k
        orc.cpu_nms = timed_nms2
    try:
        cls_prob, pred_boxes, _ = orc.faster_rcnn_forward(x, params, info)
        orc.detect(cls_prob, pred_boxes, 0.3, 0.8)          # forward.py:48-57
    finally:
        orc.cpu_nms = saved
    return time.perf_counter() - t0, tn[0]


def pick_cpu_threads(torch):
    """Give the CPU arm its best case: time one mid-trunk 3x3 convolution at a few intra-op thread counts and keep the
    fastest (all hardware threads is often NOT the fastest for torch-CPU on a many-core host)."""
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (cores, cores // 2, cores // 4, 32, 16) if 1 <= c <= cores}, reverse=True)
    x = torch.randn(1, 256, 150, 250)
    w = torch.randn(256, 256, 3, 3)
    best, best_t = cores, None
    for c in cands:
        torch.set_num_threads(c)
        torch.nn.functional.conv2d(x, w, padding=1)
        t = None
        for _ in range(5):
            t0 = time.perf_counter()
            torch.nn.functional.conv2d(x, w, padding=1)
            dt = time.perf_counter() - t0
            t = dt if t is None else min(t, dt)
        if best_t is None or t < 0.9 * best_t:          # fewer threads only when clearly (>10%) faster
            best, best_t = c, t
    torch.set_num_threads(best)
    return best


def run_reference_arm(args, rank):
    if rank != 0:
        return
    import torch
    import frcnn_oracle as orc
    import build_ref
    cores = pick_cpu_threads(torch)
    ref_nms = build_ref.load()
    params = orc.make_params(seed=1234)
    x = None
    raw = np.random.default_rng(7).integers(0, 256, (375, 625, 3), dtype=np.uint8)     # resizes to 600x1000
    info = np.array([[H_IMG, W_IMG]], np.int32)
    for _ in range(max(1, min(args.warmup, 1))):
        t_probe, _ = cpu_pipeline_once(orc, params, x, info, ref_nms, raw=raw)
    steps = args.steps
    if t_probe * steps > 240.0:                 # keep the whole run within a few minutes
        steps = max(1, int(240.0 / t_probe))
    ts, tn = [], []
    for _ in range(steps):
        a, b = cpu_pipeline_once(orc, params, x, info, ref_nms, raw=raw)
        ts.append(a)
        tn.append(b)
    total = sum(ts)
    val = steps / total
    kind = "reference" if ref_nms is not None else "port"
    line = {
        "impl": "reference", "metric": "images/sec end-to-end VGG16 Faster R-CNN forward @600x1000",
        "value": val, "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": 1,
        "ms_per_step": 1e3 * total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "VGG16 Faster R-CNN forward, synthetic 600x1000, 300 proposals (config #2)",
                   "device": "host CPU", "requested_steps": args.steps},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": kind,
                         "sample": "%d whole image(s): raw uint8 375x625 -> preprocessing -> 600x1000 forward -> per-class "
                                   "NMS; dense ops torch-CPU fp32 (Chainer not installable offline), NMS = %s; NMS share %.1f%%" %
                                   (steps, "reference cpu_nms.pyx (oracle/_ref)" if ref_nms is not None else "C port",
                                    100.0 * sum(tn) / total)},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- B200 arm
def conv_layer_table(plan, torch, reps=3, spin=True):
    """Per-launch device time of the tensor-core kernel over one forward: eager re-run with CUDA events
    around each frcnn_conv2d call (same stream, same buffers).  Returns [(name, ms, gflop)]."""
    from frcnn_b200 import ops
    rows = []
    orig = ops.conv2d

    def timed(x, w_hi, w_lo, bias, ksize, relu, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # a ~40 us spin kernel first: while the GPU spins, the host enqueues e0 + the conv launch + e1, so the interval
        # between the events is the kernel's execution alone (no host launch latency inside it, even on a slow host)
        if spin:
            torch.cuda._sleep(80000)
        e0.record()
        r = orig(x, w_hi, w_lo, bias, ksize, relu, **kw)
        e1.record()
        Hh, Ww, Cin = x.hi.shape
        taps, Cout, _ = w_hi.shape
        k_true = 27 if (Cin == 32 and taps == 1 and Hh > 1) else taps * Cin     # conv1_1 = im2col GEMM, 27 real K
        rows.append([(e0, e1), 2.0 * Hh * Ww * Cout * k_true / 1e9, "%dx%dx%d->%d k%d" % (Hh, Ww, Cin, Cout, ksize)])
        return r
    ops.conv2d = timed
    acc = {}
    try:
        for _ in range(reps):
            rows.clear()
            plan._run()
            torch.cuda.synchronize()
            for i, (ev, gf, name) in enumerate(rows):
                acc.setdefault(i, [name, gf, []])[2].append(ev[0].elapsed_time(ev[1]))
    finally:
        ops.conv2d = orig
    return [(v[0], min(v[2]), v[1]) for _, v in sorted(acc.items())]


def run_b200_arm(args, rank, local_rank, world):
    import torch
    import frcnn_oracle as orc              # synthetic weights / image generators + cpu_baseline only
    from frcnn_b200.engine import Engine

    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    anchors = orc.generate_anchors(ratios=(0.5, 1, 2), scales=(8, 16, 32))
    params = orc.make_params(seed=1234)
    eng = Engine(params, precision=args.precision, anchors=anchors, use_graph=True)
    plan = eng.plan(H_IMG, W_IMG)
    n_img = 4                                # rotate distinct images: no step sees the previous step's input
    from frcnn_b200.shard import image_seed
    imgs_host = [torch.from_numpy(orc.make_image(H_IMG, W_IMG, seed=image_seed(rank, i))[0]).pin_memory() for i in range(n_img)]
    imgs_dev = [t.cuda() for t in imgs_host]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- device-timed value: inputs resident in HBM
    from frcnn_b200.engine import LanePool
    pool = LanePool(plan, lanes=args.in_flight)      # args.in_flight independent images in flight (one stream + graph each)
    # clocks / throttle reasons are sampled (nvidia-smi, every 100 ms) from the warm-up through the timed region and the
    # one-image-in-flight repeat of it: the GPU is under the same load throughout, and a 25 ms timed region alone would
    # yield a single sample
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    pool.fork()
    for i in range(max(args.warmup, 3)):
        pool.submit(i, imgs_dev[i % n_img])
    pool.join()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    pool.fork()
    for i in range(args.steps):
        pool.submit(i, imgs_dev[i % n_img])          # every step = one whole image through the whole path
    pool.join()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    R_last = int(plan.prop.count.item())
    # the same K steps with ONE image in flight (no overlap between images): reported beside the headline
    barrier()
    e0.record()
    for i in range(args.steps):
        plan.forward(imgs_dev[i % n_img])
    e1.record()
    barrier()
    ms_single = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    # per-launch times of the tensor-core kernel, taken right here: same thermal / power state as the timed region above
    table = conv_layer_table(plan, torch) if rank == 0 else None
    from frcnn_b200 import shard
    ms_max = shard.max_over_ranks(ms, device="cuda")          # slowest rank decides
    value = world * args.steps / (ms_max / 1e3)

    # ---------------- e2e: host image in, host result out, every step (public call)
    res_prob = torch.empty((plan.post_n, 21), dtype=torch.float32).pin_memory()
    res_box = torch.empty((plan.post_n, 84), dtype=torch.float32).pin_memory()
    res_cnt = torch.empty((1,), dtype=torch.int32).pin_memory()

    dbg = os.environ.get("FRCNN_BENCH_DEBUG") == "1"
    dbg_rows = []

    def e2e_step(i):
        t = [time.perf_counter()]
        plan.x_in.copy_(imgs_host[i % n_img], non_blocking=True)        # H2D from pinned memory
        t.append(time.perf_counter())
        prob, boxes, count = plan.forward(None)
        t.append(time.perf_counter())
        res_prob.copy_(prob, non_blocking=True)                         # D2H
        res_box.copy_(boxes, non_blocking=True)
        res_cnt.copy_(count, non_blocking=True)
        t.append(time.perf_counter())
        torch.cuda.synchronize()                                        # the caller reads the result
        t.append(time.perf_counter())
        if dbg:
            dbg_rows.append([1e3 * (b - a) for a, b in zip(t, t[1:])])
        return int(res_cnt[0])
    for i in range(3):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    t_serial = time.perf_counter() - t0                      # one image at a time: latency, not throughput
    if dbg:
        print("e2e serial phases ms [h2d-call, replay-call, d2h-calls, sync] median:",
              np.round(np.median(np.array(dbg_rows[3:]), 0), 3), file=sys.stderr)
    # the public streaming call: host image in, host result out for EVERY step; the H2D of image i+1
    # overlaps the graph of image i (frcnn_b200.engine.StreamRunner)
    from frcnn_b200.engine import StreamRunner
    runner = StreamRunner(pool)
    seq = [imgs_host[i % n_img] for i in range(args.steps)]
    runner.run(seq[: min(4, len(seq))])                      # warm-up
    barrier()
    t0 = time.perf_counter()
    counts = runner.run(seq)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    assert len(counts) == args.steps and all(c > 0 for c in counts)
    e2e_val = world * args.steps / shard.max_over_ranks(t_e2e, device="cuda")
    e2e_serial_ms = 1e3 * shard.max_over_ranks(t_serial, device="cuda") / args.steps
    h2d = runner.h2d_bytes
    d2h = runner.d2h_bytes
    # same streaming call fed with RAW uint8 375x625 BGR images: mean-subtract + OpenCV-compatible bilinear resize to
    # 600x1000 run on the device (forward.py:34-45 moved onto the GPU); reported beside the float32-input number
    rng8 = np.random.default_rng(7 + rank)
    raw = [torch.from_numpy(rng8.integers(0, 256, (375, 625, 3), dtype=np.uint8)).pin_memory() for _ in range(n_img)]
    runner8 = StreamRunner(pool, src_hw=(375, 625))
    seq8 = [raw[i % n_img] for i in range(args.steps)]
    runner8.run(seq8[: min(4, len(seq8))])
    barrier()
    t0 = time.perf_counter()
    runner8.run(seq8)
    torch.cuda.synchronize()
    e2e8_val = world * args.steps / shard.max_over_ranks(time.perf_counter() - t0, device="cuda")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the tensor-core kernel (live, CUDA events, rank 0)
    peak_tf, peak_hbm, peak_src = load_peaks()
    conv_rows = [r for r in table if " k3" in r[0] or "->64 k1" in r[0]]   # trunk + RPN convs (3x3 and the twin 1x1)
    conv_ms = sum(r[1] for r in conv_rows)
    all_ms = sum(r[1] for r in table)
    exec_mult = 3.0 if args.precision == "bf16x3" else 1.0
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_conv_stack_dram.json")
    if args.precision == "bf16x3" and os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_src = tj["conv_stack_dram_bytes_per_step"], tj["source"]
    achieved = CONV_STACK_GFLOP / conv_ms          # GFLOP/ms == TFLOP/s (algorithmic flops)
    roofline = {
        "bound": "tensor", "kernel": "conv_gemm_kernel (tcgen05 implicit-GEMM, %d launches/step)" % len(table),
        "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
        "peak_source": "%s bf16_tflops_sustained (kernel timed inside a step)" % peak_src,
        "algorithmic_gflop_per_step": CONV_STACK_GFLOP, "conv_stack_ms": conv_ms, "all_gemm_ms": all_ms,
        "executed_mma_flop_multiplier": exec_mult, "executed_tflops": achieved * exec_mult,
        "traffic": traffic,     # DRAM read+write bytes of the conv-stack launches of one step, from the committed ncu pass
        "traffic_source": traffic_src,
        # every conv-stack activation written once + read once (hi+lo planes, 4 B/elem), the im2col input read
        # once, weights (17.1 M params, hi+lo) read once: 1.02 GB; measured DRAM traffic below that
        # means part of the producer->consumer traffic stays in the 126 MB L2, above it would mean re-reads
        "algorithmic_bytes_per_step": 1.02e9 if args.precision == "bf16x3" else 0.51e9,
        "layers": [{"shape": n, "ms": round(m, 4), "tflops_algorithmic": round(g / m, 1)} for n, m, g in table],
    }

    # ---------------- CPU baseline (rank 0, N == 1 only): the oracle pipeline on this box's cores
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        import build_ref
        cores = pick_cpu_threads(torch)
        ref_nms = build_ref.load()
        raw0 = np.random.default_rng(7).integers(0, 256, (375, 625, 3), dtype=np.uint8)
        info = np.array([[H_IMG, W_IMG]], np.int32)
        cpu_pipeline_once(orc, params, None, info, ref_nms, raw=raw0)          # warm-up
        tt, tn = cpu_pipeline_once(orc, params, None, info, ref_nms, raw=raw0)
        cpu_baseline = {"value": 1.0 / tt, "unit": "images/s", "cores": cores,
                        "kind": "reference" if ref_nms is not None else "port",
                        "sample": "1 whole image (raw 375x625 -> 600x1000) after 1 warm-up (%.2f s, NMS %.2f s); dense ops torch-CPU fp32 "
                                  "stand-in for Chainer, NMS = %s" %
                                  (tt, tn, "reference cpu_nms.pyx" if ref_nms is not None else "C port of cpu_nms.pyx")}

    line = {
        "metric": "images/sec end-to-end VGG16 Faster R-CNN forward @600x1000",
        "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x3 (bf16 hi/lo split operands, 3 tcgen05 MMAs per product, fp32 accumulate)"
                 if args.precision == "bf16x3" else "bf16 (fp32 accumulate)",
        "data": "synthetic",
        "config": {"workload": "VGG16 Faster R-CNN forward, synthetic 600x1000, 300 proposals (config #2), one image per GPU",
                   "precision": args.precision, "proposals_last_step": R_last,
                   "l2": "per-step working set (activations+weights ~1.5 GB) exceeds the 126 MB L2; 4 input images rotated",
                   "images_in_flight_per_gpu": len(pool),
                   "one_image_in_flight": {"images_per_s_this_rank": args.steps / (ms_single / 1e3),
                                           "ms_per_image": ms_single / args.steps},
                   "cuda_graph": True, "frac_of_conv_roofline": (value / world) * CONV_STACK_GFLOP / 1e3 / peak_tf},
        "clocks": clocks,
        # headline e2e = the public streaming call fed with what a caller actually has: the decoded RAW image.
        # Every step: H2D of a pinned uint8 375x625 BGR image, device preprocessing (forward.py:34-45: mean-sub +
        # OpenCV-compatible bilinear resize to 600x1000), the whole graph, D2H of (prob, boxes, count).
        "e2e": {"value": e2e8_val, "unit": "images/s", "h2d_bytes_per_step": runner8.h2d_bytes, "d2h_bytes_per_step": d2h,
                "mode": "StreamRunner(src_hw=(375,625)): pinned RAW uint8 image H2D + device preprocessing + graph + D2H "
                        "of (prob, boxes, count) every step; ring of %d slots, copy stream + %d compute lanes" % (runner8.depth, len(pool)),
                "float32_chw_input": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": h2d,
                                      "note": "the reference's model-input interface: the already preprocessed "
                                              "(3,600,1000) float32 tensor is uploaded every step (7.2 MB)"},
                "latency_ms_one_image_serial_float32_input": e2e_serial_ms},
        "gpu_launches": (plan.n_launches + 1) * args.steps,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_train_arm(args, rank, local_rank, world):
    """Secondary workload (BASELINE config #5, a "next" row): one train_rpn.py step per image -- forward, AnchorTargetLayer,
    RPN losses, backward through 15 convs, gradient all-reduce over the ranks, WeightDecay + MomentumSGD.  Not the headline."""
    import torch
    import frcnn_oracle as orc
    from frcnn_b200 import shard
    from frcnn_b200.train_engine import RpnTrainer
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    anchors = orc.generate_anchors(ratios=(0.5, 1, 2), scales=(8, 16, 32))
    params = orc.make_params(seed=1234)
    tr = RpnTrainer(params, H_IMG, W_IMG, anchors, precision=args.precision, subsample="device")
    imgs = [torch.from_numpy(orc.make_image(H_IMG, W_IMG, seed=shard.image_seed(rank, i))[0]).cuda() for i in range(4)]
    gt = torch.tensor([[100, 120, 400, 380, 3], [500, 200, 900, 560, 7], [50, 50, 200, 180, 1], [600, 30, 780, 150, 5]],
                      dtype=torch.float32).cuda()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    for i in range(max(args.warmup, 3)):
        tr.step(imgs[i % 4], gt)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        tr.step(imgs[i % 4], gt)
    e1.record()
    barrier()
    ms = shard.max_over_ranks(e0.elapsed_time(e1), device="cuda")
    clocks = sampler.stop() if rank == 0 else None
    losses = [float(v) for v in tr.last_losses.cpu().numpy()]
    if rank == 0:
        cpu_baseline = None
        if not args.no_cpu_baseline and world == 1:
            cores = pick_cpu_threads(torch)
            x = orc.make_image(H_IMG, W_IMG, seed=0)
            info = np.array([[H_IMG, W_IMG]], np.int32)
            t0 = time.perf_counter()
            r = orc.anchor_target_layer(38, 63, gt.cpu().numpy()[None], info, choice=lambda a, n: np.asarray(a)[:n])
            orc.rpn_train_step(params, x, r["labels"], r["targets"], r["inds_inside"], dtype="float32")
            tt = time.perf_counter() - t0
            cpu_baseline = {"value": 1.0 / tt, "unit": "images/s", "cores": cores, "kind": "port",
                            "sample": "1 step, 600x1000: AnchorTargetLayer restatement + torch-CPU fp32 autograd standing in for "
                                      "Chainer's backward (%.1f s)" % tt}
        print(json.dumps({
            "metric": "images/sec through the train_rpn.py step (forward + AnchorTargetLayer + RPN loss + backward + MomentumSGD) @600x1000",
            "value": world * args.steps / (ms / 1e3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": "train_rpn.py step, VGG16 trunk + RPN trainable, one 600x1000 image per GPU per step, "
                                   "all-reduce(SUM) of the flat fp32 gradient bucket (config #5; secondary workload)",
                       "last_losses_cls_bbox_acc_total": losses},
            "clocks": clocks, "cpu_baseline": cpu_baseline}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_train_rcnn_arm(args, rank, local_rank, world):
    """Secondary workload: one train_rcnn.py step per image (frozen RPN -> ProposalTargetLayer -> RoI pool -> fc6/fc7 with
    dropout -> losses -> backward through head, RoI pool and trunk -> MomentumSGD), device-side sampling disabled: the
    reference-faithful NumPy sampling (one small D2H per step) is what is timed.  Not the headline."""
    import torch
    import frcnn_oracle as orc
    from frcnn_b200 import shard
    from frcnn_b200.train_engine import RcnnTrainer
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    anchors = orc.generate_anchors(ratios=(0.5, 1, 2), scales=(8, 16, 32))
    params = orc.make_params(seed=1234)
    tr = RcnnTrainer(params, H_IMG, W_IMG, anchors, precision=args.precision, lr=1e-5)
    imgs = [torch.from_numpy(orc.make_image(H_IMG, W_IMG, seed=shard.image_seed(rank, i))[0]).cuda() for i in range(4)]
    gt = torch.tensor([[100, 120, 400, 380, 3], [500, 200, 900, 560, 7], [50, 50, 200, 180, 1], [600, 30, 780, 150, 5]],
                      dtype=torch.float32).cuda()
    np.random.seed(1234 + rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    for i in range(max(args.warmup, 3)):
        tr.step(imgs[i % 4], gt)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        tr.step(imgs[i % 4], gt)
    e1.record()
    barrier()
    ms = shard.max_over_ranks(e0.elapsed_time(e1), device="cuda")
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        print(json.dumps({
            "metric": "images/sec through the train_rcnn.py step @600x1000 (300 proposals, 128 RoIs per image)",
            "value": world * args.steps / (ms / 1e3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": "train_rcnn.py step, trunk + fc6/fc7/cls_score/bbox_pred trainable, one 600x1000 image per GPU "
                                   "per step (secondary workload)", "kept_rois_last_step": int(tr.keep.numel()),
                       "last_losses_cls_bbox_acc_total": [float(v) for v in tr.last_losses.cpu().numpy()]},
            "clocks": clocks}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_resnet_arm(args, rank, local_rank, world):
    """Secondary workload (BASELINE config #4, a "next" row): ResNet-101 trunk Faster R-CNN forward, 800x1333, 1000 proposals,
    one image per GPU.  Device-timed with inputs resident in HBM, args.in_flight images in flight.  Not the headline."""
    import torch
    import frcnn_oracle as orc
    from frcnn_b200 import shard
    from frcnn_b200.engine import LanePool
    from frcnn_b200.resnet_engine import ResNetEngine
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    H, W = 800, 1333
    anchors = orc.generate_anchors(ratios=(0.5, 1, 2), scales=(8, 16, 32))
    params = orc.make_resnet_params(101, seed=4321)
    eng = ResNetEngine(params, 101, precision=args.precision, anchors=anchors, use_graph=True, post_n=1000)
    plan = eng.plan(H, W)
    imgs = [torch.from_numpy(orc.make_image(H, W, seed=shard.image_seed(rank, i))[0]).cuda() for i in range(4)]
    pool = LanePool(plan, lanes=args.in_flight)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    pool.fork()
    for i in range(max(args.warmup, 3)):
        pool.submit(i, imgs[i % 4])
    pool.join()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    pool.fork()
    for i in range(args.steps):
        pool.submit(i, imgs[i % 4])
    pool.join()
    e1.record()
    barrier()
    ms = shard.max_over_ranks(e0.elapsed_time(e1), device="cuda")
    clocks = sampler.stop() if rank == 0 else None
    R_last = int(plan.prop.count.item())
    barrier()
    e0.record()
    for i in range(args.steps):
        plan.forward(imgs[i % 4])
    e1.record()
    barrier()
    ms1 = e0.elapsed_time(e1)
    if rank == 0:
        print(json.dumps({
            "metric": "images/sec end-to-end ResNet-101 Faster R-CNN forward @800x1333, 1000 proposals",
            "value": world * args.steps / (ms / 1e3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": "ResNet-101 trunk Faster R-CNN forward, synthetic 800x1333, 1000 proposals, one image per GPU "
                                   "(config #4; secondary workload)", "images_in_flight_per_gpu": len(pool),
                       "proposals_last_step": R_last, "one_image_in_flight_ms": ms1 / args.steps,
                       "launches_per_image": plan.n_launches},
            "clocks": clocks}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--in-flight", type=int, default=3, help="independent images in flight per GPU (streams/graphs)")
    ap.add_argument("--workload", default="forward", choices=["forward", "train_rpn", "train_rcnn", "resnet101"],
                    help="forward = the headline metric (default); train_rpn / resnet101 = secondary workloads (configs #5 / #4)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.workload == "train_rpn":
        run_train_arm(args, rank, local_rank, world)
        return
    if args.workload == "resnet101":
        run_resnet_arm(args, rank, local_rank, world)
        return
    if args.workload == "train_rcnn":
        run_train_rcnn_arm(args, rank, local_rank, world)
        return
    run_b200_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
