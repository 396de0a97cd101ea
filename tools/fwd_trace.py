"""Runs one batch-32 forward through a -DCOATTN_TRACE build (clock64 stamps printed by CTA 0 of the attend kernel)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosnet_b200.coattention import coattention_forward_raw
dev = torch.device("cuda:0")
n, h, w = 32, 60, 60
g = torch.Generator(device=dev); g.manual_seed(1)
import torch.nn.functional as F
SIGMA = float(os.environ.get("SIGMA", "0.66"))     # bench.py's feature distribution: prelu(randn, 0.25) * sigma
va = F.prelu(torch.randn(n, 256, h, w, generator=g, device=dev), torch.tensor([0.25], device=dev)) * SIGMA
vb = F.prelu(torch.randn(n, 256, h, w, generator=g, device=dev), torch.tensor([0.25], device=dev)) * SIGMA
wt = (torch.rand(256, 256, generator=g, device=dev) * 2 - 1) / 16
gw = torch.randn(256, generator=g, device=dev) * 0.01
for _ in range(3):     # the trace build prints on every launch; the last one is warm
    coattention_forward_raw(va, vb, wt, gw, None, want_z=False)  # like bench.py: no raw-Z output
    torch.cuda.synchronize()
print("done")
