import os, sys, subprocess, threading, time
sys.path.insert(0, os.getcwd())
import torch
from cosnet_b200 import _lib
from cosnet_b200.coattention import workspace_bytes
import numpy as np
lib = _lib.load()
dev = torch.device("cuda:0")
n, c, h, w = 32, 256, 60, 60
L = h * w
va = torch.randn(n, c, h, w, device=dev); vb = torch.randn(n, c, h, w, device=dev)
va = torch.where(va > 0, va, 0.25 * va) * 0.66; vb = torch.where(vb > 0, vb, 0.25 * vb) * 0.66
W = (torch.rand(c, c, device=dev) * 2 - 1) / 16; g = torch.randn(c, device=dev) * 0.01
nbytes = workspace_bytes(n, c, h, w)
ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev); wsp = (ws.data_ptr() + 1023) // 1024 * 1024
ca = torch.empty(n, 2 * c, h, w, device=dev); cb = torch.empty_like(ca); lse = torch.empty(2, n, L, device=dev)
st = torch.cuda.current_stream().cuda_stream
lib.coattn_stage_prep(va.data_ptr(), vb.data_ptr(), W.data_ptr(), wsp, nbytes, n, c, h, w, 0, st)
lib.coattn_stage_project(wsp, nbytes, n, c, h, w, 0, st)
def attend():
    lib.coattn_stage_attend_gate(va.data_ptr(), vb.data_ptr(), ca.data_ptr(), cb.data_ptr(), None, lse.data_ptr(), None, g.data_ptr(), None, wsp, nbytes, n, c, h, w, 0, st)
for _ in range(3): attend()
torch.cuda.synchronize()
samples = []
stop = False
def sampler():
    import pynvml
    pynvml.nvmlInit(); hd = pynvml.nvmlDeviceGetHandleByIndex(0)
    while not stop:
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(hd, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(hd) / 1000.0))
        time.sleep(0.005)
th = threading.Thread(target=sampler); th.start()
for burst in (1, 10, 100, 1000):
    torch.cuda.synchronize(); time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(burst): attend()
    e1.record(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    ss = [s for s in samples if t0 <= s[0] <= t1]
    clk = np.median([s[1] for s in ss]) if ss else -1
    pw = np.max([s[2] for s in ss]) if ss else -1
    print(f"burst {burst:5d}: {e0.elapsed_time(e1)/burst:.4f} ms/launch  sm clock median {clk} MHz  max power {pw:.0f} W  ({len(ss)} samples)", flush=True)
stop = True; th.join()
