#!/usr/bin/env python
"""One cfg-5 style forward + backward (C ABI) of both modalities; run under `ncu --metrics gpu__time_duration.sum` to list
the kernels of the backward with their durations, or plain for CUDA-event timings of the backward alone."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from cosnet_b200 import _lib
from cosnet_b200.coattention import backward_workspace_bytes, workspace_bytes

n = int(os.environ.get("PAIRS", "8")); h = int(os.environ.get("FH", "60")); w = int(os.environ.get("FW", "60")); C = 256
reps = int(os.environ.get("REPS", "1"))
bflags = int(os.environ.get("BWD_FLAGS", "0"))      # 512 = COATTN_FLAG_PLANES_READY: the backward reuses the forward's planes
dev = torch.device("cuda:0")
lib = _lib.load()
g = torch.Generator(device=dev); g.manual_seed(1)
feats = lambda: F.prelu(torch.randn((n, C, h, w), generator=g, device=dev), torch.tensor([0.25], device=dev)) * 0.66
va, vb = feats(), feats()
W = (torch.rand((C, C), generator=g, device=dev) * 2 - 1) / 16
gw = torch.randn((C,), generator=g, device=dev) * 0.01
gb = torch.zeros(1, device=dev)
L = h * w
ca, cb = torch.empty((n, 2 * C, h, w), device=dev), torch.empty((n, 2 * C, h, w), device=dev)
z, lse, mask = torch.empty((2, n, C, L), device=dev), torch.empty((2, n, L), device=dev), torch.empty((2, n, L), device=dev)
ra, rb = torch.randn((n, 2 * C, h, w), generator=g, device=dev) * 1e-3, torch.randn((n, 2 * C, h, w), generator=g, device=dev) * 1e-3
dva, dw, dgw, dgb = torch.empty((n, C, h, w), device=dev), torch.empty((C, C), device=dev), torch.empty(C, device=dev), torch.empty(1, device=dev)
nbf, nbb = workspace_bytes(n, C, h, w), backward_workspace_bytes(n, C, h, w, False)
ws = torch.empty(max(nbf, nbb) + 1024, dtype=torch.uint8, device=dev)
wsp = (ws.data_ptr() + 1023) // 1024 * 1024
st = torch.cuda.current_stream(dev).cuda_stream
P = lambda t: None if t is None else t.data_ptr()
_lib.check(lib.coattn_forward(P(va), P(vb), P(W), P(gw), P(gb), P(ca), P(cb), P(z), P(lse), P(mask), wsp, nbf, n, C, h, w, bflags & 1, st), "fwd")      # bit 0 = COATTN_FLAG_BF16: forward and backward in one format
for has_b in (True, False):
    for r in range(reps + 2):
        if r == 2:
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
        _lib.check(lib.coattn_backward(P(va), P(vb), P(W), P(gw), P(z), P(lse), P(mask), P(ra), P(rb) if has_b else None, P(dva), None,
                                       P(dw), P(dgw), P(dgb), wsp, nbb, n, C, h, w, bflags, st), "bwd")
    e1.record(); torch.cuda.synchronize()
    print(f"backward has_b={has_b}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us per call ({n} pairs {h}x{w})", flush=True)
