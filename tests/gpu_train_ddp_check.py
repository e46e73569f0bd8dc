"""2-rank check of the training path's gradient all-reduce (run under torchrun, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/gpu_train_ddp_check.py
Each rank back-propagates ITS OWN image (train_rpn.py:169-174 ParallelUpdater semantics: gradients are ADDED), then both
apply the same update.  Rank 0 recomputes both gradients alone and checks: bucket == g(image0) + g(image1) exactly (the
kernels are deterministic, NCCL's 2-rank sum is a single fp32 add), and the weights of the two ranks are bit-identical."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "chainer-faster-rcnn_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import frcnn_oracle as orc  # noqa: E402
from frcnn_b200.train_engine import RpnTrainer  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    H, W = 296, 392
    anchors = orc.generate_anchors(ratios=(0.5, 1, 2), scales=(8, 16, 32))
    params = orc.make_params(seed=77)
    imgs = [torch.from_numpy(orc.make_image(H, W, seed=10 + r)[0]).cuda() for r in range(world)]
    gts = [torch.tensor([[20 + 30 * r, 30, 150 + 30 * r, 170, 3], [100, 20 + 10 * r, 330, 240, 7]], dtype=torch.float32).cuda()
           for r in range(world)]
    tr = RpnTrainer(params, H, W, anchors, subsample="none")
    tr.set_grad_exchange("fp32", overlap=False)       # the exact reference point: ONE all-reduce after backward
    tr.forward(imgs[rank], gts[rank])
    tr.backward()
    g_own = tr.g_flat.clone()
    tr.update()                                   # all-reduce(SUM) + WeightDecay + MomentumSGD + repack
    g_sum = tr.g_flat.clone()
    w_after = tr.w_flat.clone()
    # every rank's weights identical
    ws = [torch.empty_like(w_after) for _ in range(world)]
    dist.all_gather(ws, w_after)
    same = all(torch.equal(ws[0], w) for w in ws)
    ok = same
    if rank == 0:
        ref = RpnTrainer(params, H, W, anchors, subsample="none")
        ref.set_grad_exchange("fp32", overlap=False)      # rank 0 alone: backward must not start a collective
        total = torch.zeros_like(g_own)
        for r in range(world):
            ref.forward(imgs[r], gts[r])
            ref.backward()
            if r == 0:
                assert torch.equal(ref.g_flat, g_own), "kernels are not deterministic"
            total += ref.g_flat
        exact = torch.equal(total, g_sum)
        print("ranks identical:", same, "| bucket == sum of per-image gradients (exact):", exact,
              "| max |g|", float(g_sum.abs().max()), flush=True)
        ok = ok and exact
    # ---- BASELINE config #5's exchange: bf16 bucket, deep layers reduced on a side stream while the shallow layers still
    # back-propagate.  Replicas must stay bit-identical; the update must agree with the exact fp32 exchange to bf16 rounding.
    tb = RpnTrainer(params, H, W, anchors, subsample="none")
    tb.set_grad_exchange("bf16", overlap=True)
    tb.forward(imgs[rank], gts[rank])
    tb.backward()
    tb.update()
    torch.cuda.synchronize()
    wb = tb.w_flat.clone()
    ws = [torch.empty_like(wb) for _ in range(world)]
    dist.all_gather(ws, wb)
    same_b = all(torch.equal(ws[0], w) for w in ws)
    # v = -lr * (g + wd * w): compare the momentum buffers (= the applied gradients) of the two exchanges
    num = float((tb.v_flat - tr.v_flat).abs().max())
    den = float(tr.v_flat.abs().max())
    close = num <= 2.0 ** -7 * den                   # each rank's bf16 rounding (2^-9) + the bf16 sum (2^-9), relative to max |g|
    exposed = tb.last_exposed_exchange_ms()
    if rank == 0:
        print("bf16 bucket: ranks identical:", same_b, "| max |dv| / max |v| = %.2e (bound 2^-7)" % (num / max(den, 1e-30)),
              "| exposed exchange %.3f ms" % exposed, flush=True)
    ok = ok and same_b and close
    # ---- fp32 exchange with the two overlapped buckets must reproduce the single all-reduce exactly
    to = RpnTrainer(params, H, W, anchors, subsample="none")
    to.set_grad_exchange("fp32", overlap=True)
    to.forward(imgs[rank], gts[rank])
    to.backward()
    to.update()
    torch.cuda.synchronize()
    ok = ok and torch.equal(to.g_flat, g_sum) and torch.equal(to.w_flat, w_after)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        sys.exit(1)
    if rank == 0:
        print("DDP_CHECK_OK")


if __name__ == "__main__":
    main()
