#!/usr/bin/env python
"""Smallest end-to-end exercise of every product kernel (forward with the in-kernel projection, fused tail, 16-bit
interface, backward with counterpart gradients) for `compute-sanitizer --tool racecheck|memcheck python tools/sanitize_probe.py`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from cosnet_b200 import coattention
from cosnet_b200.coattention import coattention_forward16_raw
from oracle import coattn_oracle as orc
dev = torch.device("cuda:0")
n, h, w = 2, 12, 11          # L = 132: two key tiles, ragged tail, one query tile per (sample, pass)
v_a, v_b = (torch.from_numpy(x).to(dev) for x in orc.synthetic_features(7, n, h, w, 0.66))
W, g, b = (torch.from_numpy(x).to(dev) for x in orc.synthetic_weights(8, bias=True))
va = v_a.clone().requires_grad_(True); vb = v_b.clone().requires_grad_(True)
wt = W.clone().requires_grad_(True); gw = g.clone().view(1, -1, 1, 1).requires_grad_(True); gb = b.clone().requires_grad_(True)
ca, cb = coattention(va, vb, wt, gw, gb)
(ca.sum() + cb.square().sum()).backward()
c16 = coattention_forward16_raw(v_a.half(), v_b.half(), W, g, b)
torch.cuda.synchronize()
ref = orc.coattention(v_a.cpu().numpy(), v_b.cpu().numpy(), W.cpu().numpy(), g.cpu().numpy(), b.cpu().numpy())
err = float(np.linalg.norm(ca.detach().cpu().numpy() - ref["cat_a"]) / np.linalg.norm(ref["cat_a"]))
print("forward rel-L2", err, "grad finite", bool(torch.isfinite(va.grad).all() and torch.isfinite(vb.grad).all() and torch.isfinite(wt.grad).all()))
assert err < 1e-3
