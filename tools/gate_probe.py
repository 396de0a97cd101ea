#!/usr/bin/env python
"""Stand-alone gate / sigmoid / concat epilogue (`coattn_stage_gate`) at the benchmark's shape: GB/s against the measured
copy bandwidth of MEASURED_PEAKS.json, next to a plain device-to-device copy of the same byte count timed the same way.
LIB=<path to another build of the library> for an A/B."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cosnet_b200 import _lib
if os.environ.get("LIB"):
    _lib.LIB_PATH = os.environ["LIB"]
lib = _lib.load()
dev = torch.device("cuda:0")
n, C, H, W = int(os.environ.get("N", 32)), 256, 60, 60
L = H * W
_junk = []
def pad():      # SHIFT=<MB>: shifts the relative placement of the five tensors (address-mapping sensitivity)
    if os.environ.get("SHIFT"):
        _junk.append(torch.empty(int(float(os.environ["SHIFT"]) * (1 << 20)), dtype=torch.uint8, device=dev))
z = torch.randn(2, n, C, L, device=dev) * float(os.environ.get("ZSIGMA", 1.0)); pad(); va = torch.randn(n, C, H, W, device=dev); pad(); vb = torch.randn(n, C, H, W, device=dev)
g = torch.randn(C, device=dev) * 0.1
pad(); ca = torch.empty(n, 2 * C, H, W, device=dev); pad(); cb = torch.empty_like(ca)
st = torch.cuda.current_stream().cuda_stream
def run():
    _lib.check(lib.coattn_stage_gate(z.data_ptr(), va.data_ptr(), vb.data_ptr(), g.data_ptr(), None, ca.data_ptr(), cb.data_ptr(), n, C, H, W, st), "gate")
src = torch.empty(2 * n * 2 * C * L, device=dev); dst = torch.empty_like(src)
def timed(fn, reps=20):
    for _ in range(3):
        fn()
    ev = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2], t[0]
bytes_ = 2 * n * 16.0 * L * C
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
if os.environ.get("SOAK"):      # what bench.py's epilogue leg sees: seconds of tensor-core load right before it
    import time
    a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16); b = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    t0 = time.time()
    while time.time() - t0 < float(os.environ["SOAK"]):
        for _ in range(20):
            a @ b
        torch.cuda.synchronize()
if os.environ.get("PRE"):       # a full forward (attend2: 227 KB of shared memory per CTA) on the same stream right before the gate calls
    from cosnet_b200.coattention import workspace_bytes
    nb = workspace_bytes(n, C, H, W)
    ws = torch.empty(nb + 1024, dtype=torch.uint8, device=dev); wp = (ws.data_ptr() + 1023) // 1024 * 1024
    Wt = torch.randn(C, C, device=dev) / 16; lse = torch.empty(2, n, L, device=dev)
    def fwd():
        _lib.check(lib.coattn_forward(va.data_ptr(), vb.data_ptr(), Wt.data_ptr(), g.data_ptr(), None, ca.data_ptr(), cb.data_ptr(), z.data_ptr() if os.environ["PRE"] == "z" else None, lse.data_ptr(), None, wp, nb, n, C, H, W, 0, st), "fwd")
    fwd()
    if os.environ["PRE"] == "each":
        _run = run
        def run():
            fwd(); _run()
gm, gb = timed(run)
cm, cbest = timed(lambda: dst.copy_(src))
print(json.dumps({"kernel": "coattn_stage_gate", "n": n, "ms_median": gm, "ms_best": gb, "GBps_median": bytes_ / gm / 1e6, "GBps_best": bytes_ / gb / 1e6,
                  "frac_of_measured_copy_peak": bytes_ / gm / 1e6 / peaks["hbm_gbs"],
                  "torch_copy_same_bytes": {"ms_median": cm, "GBps_median": bytes_ / cm / 1e6}, "lib": os.environ.get("LIB", "in-tree"), "shift_mb": os.environ.get("SHIFT"), "ptrs_mod_2MB": [t.data_ptr() % (1 << 21) for t in (z, va, vb, ca, cb)]}))
