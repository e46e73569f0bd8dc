"""Diagnostic runner (not a pytest file): each conv/GEMM configuration in its own process so a
trapped kernel (bounded mbarrier wait) cannot poison the others.  Prints one line per config."""
import json
import subprocess
import sys

CONFIGS = [
    # H, W, Cin, Cout, k, precision, tile(bn,th,tw) or None, cta_group
    (32, 16, 64, 128, 3, "bf16", None, 2),
    (32, 16, 64, 128, 1, "bf16", None, 2),
    (33, 41, 128, 256, 3, "bf16", None, 2),
    (33, 41, 128, 256, 3, "bf16x3", None, 2),
    (33, 41, 128, 256, 3, "bf16x3", None, 1),
    (38, 63, 512, 512, 3, "bf16x3", None, 2),
    (1, 300, 1024, 256, 1, "bf16x3", None, 2),
    (75, 125, 256, 512, 3, "bf16", (256, 16, 8), 2),
]

CHILD = r'''
import sys, json, numpy as np, torch
sys.path[:0] = ["chainer-faster-rcnn_b200", "oracle", "tests"]
from frcnn_b200 import ops
H, W, Cin, Cout, k, prec, tile, cg = json.loads(sys.argv[1])
ops.set_conv_cta_group(cg)
rng = np.random.default_rng(1)
x = rng.standard_normal((Cin, H, W)).astype(np.float32)
w = (rng.standard_normal((Cout, Cin, k, k)) * (2.0 / (Cin * k * k)) ** 0.5).astype(np.float32)
def q(a):
    t = torch.from_numpy(a); hi = t.to(torch.bfloat16).float()
    if prec == "bf16": return hi.numpy()
    lo = (t - hi).to(torch.bfloat16).float(); return (hi + lo).numpy()
ref = torch.nn.functional.conv2d(torch.from_numpy(q(x))[None].double(), torch.from_numpy(q(w)).double(), padding=(k - 1) // 2)[0].numpy()
act = ops.pack_image(torch.from_numpy(x).cuda(), c_pad=Cin, precision=prec)
wh, wl = ops.pack_conv_weights(torch.from_numpy(w).cuda(), cin_pad=Cin, precision=prec)
bias = torch.zeros(max(Cout, 32) + 224, device="cuda")
if tile: ops.set_conv_tile(*tile)
ld = (Cout + 31) // 32 * 32
y, y32 = ops.conv2d(act, wh, wl, bias, k, False, out_act=True, ld_f32=ld)
torch.cuda.synchronize()
got = y32.cpu().numpy().reshape(H, W, ld)[:, :, :Cout].transpose(2, 0, 1)
err = np.abs(got - ref)
scale = np.abs(ref).max()
bad = err > 1e-3 * scale
print(json.dumps(dict(max_rel=float(err.max() / scale), frac_bad=float(bad.mean()),
      bad_by_chan=[int(v) for v in bad.reshape(Cout, -1).any(1).nonzero()[0][:8]],
      bad_rows=[int(v) for v in bad.any(0).any(1).nonzero()[0][:8]],
      got0=[float(v) for v in got.ravel()[:4]], ref0=[float(v) for v in ref.ravel()[:4]])))
'''

for cfg in CONFIGS:
    try:
        r = subprocess.run([sys.executable, "-c", CHILD, json.dumps(cfg)], capture_output=True, text=True, timeout=180)
        out = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ("rc=%d " % r.returncode) + r.stderr.strip()[-400:]
    except subprocess.TimeoutExpired:
        out = "TIMEOUT"
    print(cfg, "->", out, flush=True)
