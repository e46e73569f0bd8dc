"""Diagnostic (not a pytest file): does an nvidia-smi poller stall cudaMemcpyAsync? does terminate() kill it?"""
import os, signal, sys, time, subprocess
import numpy as np, torch
x = torch.empty((3, 600, 1000), dtype=torch.float32).pin_memory(); d = torch.empty_like(x, device="cuda")
def pattern(label, n=300):
    ts = []
    for _ in range(n):
        a = time.perf_counter(); d.copy_(x, non_blocking=True); torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - a))
    ts = np.array(ts)
    print("%-44s median %.2f ms, max %.1f ms, stalls>5ms: %d of %d, total %.0f ms" % (label, np.median(ts), ts.max(), (ts > 5).sum(), n, ts.sum()))
pattern("before any nvidia-smi")
cmd = ["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits", "-lms", "100"]
p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
time.sleep(0.5); pattern("while nvidia-smi -lms 100 runs")
p.terminate(); time.sleep(0.5); pattern("after Popen.terminate() (poll=%s)" % p.poll())
print(subprocess.run("ps -eo pid,ppid,pgid,cmd | grep -i nvidia-smi | grep -v grep", shell=True, capture_output=True, text=True).stdout)
p2 = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, start_new_session=True)
time.sleep(0.5)
os.killpg(os.getpgid(p2.pid), signal.SIGKILL); time.sleep(0.5); pattern("after killpg of a new-session sampler")
print(subprocess.run("ps -eo pid,ppid,pgid,cmd | grep -i nvidia-smi | grep -v grep; file $(which nvidia-smi)", shell=True, capture_output=True, text=True).stdout)
