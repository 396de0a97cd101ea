#!/usr/bin/env python
"""A/B of two builds of the library on one box: attend-stage time at the headline shape, alternating, CUDA events.
    python tools/ab_attend.py libA.so libB.so"""
import ctypes, os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 2:
    for rnd in range(3):
        for lib in sys.argv[1:]:
            env = dict(os.environ, COATTN_B200_LIB=os.path.abspath(lib))
            out = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True).stdout.strip().splitlines()[-1]
            print(os.path.basename(lib), out, flush=True)
    sys.exit(0)
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
from cosnet_b200 import _lib
from cosnet_b200.coattention import workspace_bytes
lib = _lib.load()
dev = torch.device("cuda:0")
n, C, H, W = 32, 256, 60, 60
g = torch.Generator(device=dev); g.manual_seed(1)
f = lambda: F.prelu(torch.randn(n, C, H, W, generator=g, device=dev), torch.tensor([0.25], device=dev)) * 0.66
va, vb = f(), f()
wt = (torch.rand(C, C, generator=g, device=dev) * 2 - 1) / 16; gw = torch.randn(C, generator=g, device=dev) * 0.01
ca, cb = torch.empty(n, 2 * C, H, W, device=dev), torch.empty(n, 2 * C, H, W, device=dev)
lse = torch.empty(2, n, H * W, device=dev)
nb = workspace_bytes(n, C, H, W); ws = torch.empty(nb + 1024, dtype=torch.uint8, device=dev); wp = (ws.data_ptr() + 1023) // 1024 * 1024
st = torch.cuda.current_stream().cuda_stream
NOPASS = os.environ.get("AB_NOPASS") == "1"      # attend without the passthrough copy warp's traffic (v_a / v_b = NULL)
def step():
    _lib.check(lib.coattn_stage_prep_project(va.data_ptr(), vb.data_ptr(), wt.data_ptr(), wp, nb, n, C, H, W, 0, st), "p")
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
    _lib.check(lib.coattn_stage_attend_gate(None if NOPASS else va.data_ptr(), None if NOPASS else vb.data_ptr(), ca.data_ptr(), cb.data_ptr(), None, lse.data_ptr(), None, gw.data_ptr(), None, wp, nb, n, C, H, W, 0, st), "a")
    e1.record(); return e0, e1
for _ in range(5): step()
torch.cuda.synchronize()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True); t0.record()
ev = [step() for _ in range(20)]
t1.record(); torch.cuda.synchronize()
print(json.dumps({"attend_ms": sum(a.elapsed_time(b) for a, b in ev) / len(ev), "call_ms": t0.elapsed_time(t1) / 20}))
