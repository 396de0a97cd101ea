#!/usr/bin/env python
"""Host cost of one eager modality call at batch 1 (60x60): wall time per call with the GPU kept busy-free (sync every call is NOT
done: we measure issue rate), split by cProfile."""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cosnet_b200 import coattention
from cosnet_b200.coattention import coattention_forward_raw
from cosnet_b200 import _lib
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(0)
va, vb = (torch.randn(1, 256, 60, 60, device=dev, generator=g) for _ in range(2))
W = torch.randn(256, 256, device=dev, generator=g) / 16; gw = torch.randn(256, device=dev, generator=g) * 0.01
for _ in range(20):
    coattention_forward_raw(va, vb, W, gw, None, want_z=False, want_lse=False)
torch.cuda.synchronize()
N = 400
t0 = time.perf_counter()
for _ in range(N):
    coattention_forward_raw(va, vb, W, gw, None, want_z=False, want_lse=False)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"issue: {(t1 - t0) / N * 1e6:.1f} us per modality call; with final sync {(t2 - t0) / N * 1e6:.1f} us")
# library call alone
lib = _lib.load()
from cosnet_b200.coattention import workspace_bytes
nb = workspace_bytes(1, 256, 60, 60)
ws = torch.empty(nb + 1024, dtype=torch.uint8, device=dev); wp = (ws.data_ptr() + 1023) // 1024 * 1024
ca = torch.empty(1, 512, 60, 60, device=dev); cb = torch.empty_like(ca); lse = torch.empty(2, 1, 3600, device=dev)
st = torch.cuda.current_stream().cuda_stream
args = (va.data_ptr(), vb.data_ptr(), W.data_ptr(), gw.data_ptr(), None, ca.data_ptr(), cb.data_ptr(), None, lse.data_ptr(), None, wp, nb, 1, 256, 60, 60, 0, st)
for _ in range(20):
    lib.coattn_forward(*args)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(N):
    lib.coattn_forward(*args)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"library call alone (ctypes, 3 launches): {(t1 - t0) / N * 1e6:.1f} us")
pr = cProfile.Profile(); pr.enable()
for _ in range(N):
    coattention_forward_raw(va, vb, W, gw, None, want_z=False, want_lse=False)
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18); print(s.getvalue()[:3500])
# what the pieces of the Python path cost on this host (us per call, no profiler)
def per(fn, n=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    d = (time.perf_counter() - t) / n * 1e6
    torch.cuda.synchronize()
    return d
host = torch.zeros(4, dtype=torch.int32).pin_memory(); status = torch.zeros(4, dtype=torch.int32, device=dev)
ev = torch.cuda.Event(); cur = torch.cuda.current_stream(dev)
pieces = {
    "torch.empty [1,512,60,60]": lambda: torch.empty((1, 512, 60, 60), dtype=torch.float32, device=dev),
    "pinned.copy_(status, non_blocking)": lambda: host.copy_(status, non_blocking=True),
    "event.record(stream)": lambda: ev.record(cur),
    "torch.cuda.current_stream(dev)": lambda: torch.cuda.current_stream(dev),
    "_cuda_getCurrentRawStream": lambda: torch._C._cuda_getCurrentRawStream(0),
    "is_current_stream_capturing": lambda: torch.cuda.is_current_stream_capturing(),
    "with torch.cuda.device(dev)": lambda: torch.cuda.device(dev).__enter__(),
    "_check_inputs": lambda: sys.modules["cosnet_b200.coattention"]._check_inputs(va, vb, W, gw, None),
    "coattention() no-grad": lambda: sys.modules["cosnet_b200.coattention"].coattention(va, vb, W, gw, None),
}
with torch.no_grad():
    for k, f in pieces.items():
        print(f"  {k}: {per(f, 400 if 'coattention()' in k else 2000):.2f} us")
