"""Diagnostic (not a pytest file): host<->device copy behaviour on the GPU box."""
import time, torch
x = torch.empty((3, 600, 1000), dtype=torch.float32)
xp = x.pin_memory()
print("is_pinned:", x.is_pinned(), xp.is_pinned())
d = torch.empty_like(x, device="cuda")
s2 = torch.cuda.Stream()
def t(fn, n=10):
    fn(); torch.cuda.synchronize(); a = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - a) / n * 1e3
print("pageable H2D   %.3f ms" % t(lambda: d.copy_(x, non_blocking=True)))
print("pinned   H2D   %.3f ms" % t(lambda: d.copy_(xp, non_blocking=True)))
def on_s2():
    with torch.cuda.stream(s2): d.copy_(xp, non_blocking=True)
print("pinned H2D on side stream %.3f ms" % t(on_s2))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); d.copy_(xp, non_blocking=True); e1.record(); torch.cuda.synchronize()
print("pinned H2D gpu-timeline %.3f ms -> %.2f GB/s" % (e0.elapsed_time(e1), 7.2e-3 / e0.elapsed_time(e1) * 1e3 / 1e0))
big = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32).pin_memory(); dbig = torch.empty_like(big, device="cuda")
e0.record(); dbig.copy_(big, non_blocking=True); e1.record(); torch.cuda.synchronize()
print("pinned H2D 256MB: %.2f GB/s" % (0.268 / e0.elapsed_time(e1) * 1e3))
h = torch.empty((300, 84), dtype=torch.float32).pin_memory(); dd = torch.empty((300, 84), device="cuda")
print("pinned D2H 100KB + sync %.3f ms" % t(lambda: (h.copy_(dd, non_blocking=True), torch.cuda.synchronize())))
# does an H2D on the side stream overlap a long kernel on the main stream?
a = torch.randn(8192, 8192, device="cuda"); 
def overlap():
    y = a @ a
    with torch.cuda.stream(s2): d.copy_(xp, non_blocking=True)
    return y
tk = t(lambda: a @ a, 5); to = t(overlap, 5)
print("matmul alone %.3f ms, matmul + side-stream H2D %.3f ms" % (tk, to))
import subprocess
print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max,pcie.link.width.max --format=csv", shell=True, capture_output=True, text=True).stdout)
