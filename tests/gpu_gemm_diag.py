"""Diagnostic (not a pytest): split-K GEMM mode feature isolation, one subprocess per configuration."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys
sys.path.insert(0, %r)
import torch
from frcnn_b200 import train_ops as t
M, N, K, groups, splits, rs = [int(v) for v in sys.argv[1:7]]
A = torch.randint(-3, 4, (M, K), device="cuda").float()
B = torch.randint(-3, 4, (N, K), device="cuda").float()
p = t.gemm_nt_splitk(A.to(torch.bfloat16), None, B.to(torch.bfloat16), None, groups=groups, row_stride=rs, splits=splits)
torch.cuda.synchronize()
S = p.shape[1]
per = -(-(K // 64) // S) * 64
ok = True
for g in range(groups):
    off = ((g // 3 - 1) * rs + (g %% 3 - 1)) if groups == 9 else 0
    Bs = torch.zeros_like(B)
    lo, hi = max(0, -off), min(K, K - off)
    Bs[:, lo:hi] = B[:, lo + off:hi + off]
    for s in range(S):
        k0, k1 = s * per, min(K, (s + 1) * per)
        want = A[:, k0:k1].double() @ Bs[:, k0:k1].double().T
        if not torch.equal(p[g, s, :, :N].double(), want):
            ok = False
            print("  mismatch g", g, "s", s, "max", (p[g, s, :, :N].double() - want).abs().max().item())
print("OK" if ok else "WRONG", p.shape)
''' % os.path.join(ROOT, "chainer-faster-rcnn_b200")
for cfg in [(128, 64, 640, 1, 1, 0), (64, 64, 640, 1, 1, 0), (128, 64, 640, 1, 3, 0), (128, 64, 640, 9, 1, 8), (128, 64, 640, 9, 1, 37),
            (64, 64, 2368, 9, 5, 37), (512, 512, 2560, 9, 3, 37)]:
    r = subprocess.run([sys.executable, "-c", CHILD] + [str(v) for v in cfg], capture_output=True, text=True, timeout=300)
    print(cfg, "->", (r.stdout.strip().splitlines() or ["?"])[-3:], (r.stderr.strip().splitlines() or [""])[-1][:200], flush=True)
