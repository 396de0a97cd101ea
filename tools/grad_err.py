#!/usr/bin/env python
"""Gradient rel-L2 of the CUDA backward against the fp64 oracle for a few shapes / feature scales (prints a table)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from cosnet_b200 import coattention
from oracle import coattn_oracle as orc
from tests.helpers import rel_l2
dev = torch.device("cuda:0")
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
for (n, h, w, sigma, gscale, bf16) in ((2, 12, 11, 0.66, 1.0, False), (1, 31, 41, 0.66, 1.0, False), (1, 60, 60, 0.66, 1e-4, False),
                                      (1, 40, 40, 1.0, 1.0, False), (1, 40, 40, 1.0, 1e3, False), (1, 31, 41, 0.66, 1.0, True)):
    v_a, v_b = orc.synthetic_features(300 + h * w, n, h, w, sigma)
    W, g, b = orc.synthetic_weights(301 + h * w, bias=True)
    rng = np.random.default_rng(5)
    r_a = (rng.standard_normal((n, 512, h, w), dtype=np.float32) * gscale).astype(np.float32)
    r_b = (rng.standard_normal((n, 512, h, w), dtype=np.float32) * gscale).astype(np.float32)
    va = t(v_a).requires_grad_(True); wt = t(W).requires_grad_(True); gw = t(g).view(1, -1, 1, 1).requires_grad_(True); gb = t(b).requires_grad_(True)
    ca, cb = coattention(va, t(v_b), wt, gw, gb, bf16)
    ((ca * t(r_a)).sum() + (cb * t(r_b)).sum()).backward()
    torch.cuda.synchronize()
    ref = orc.coattention_grads(v_a, v_b, W, g, b, r_a, r_b)
    print(f"n={n} {h}x{w} sigma={sigma} cotangent x{gscale:g} bf16={bf16}: d_v_a {rel_l2(va.grad.cpu().numpy(), ref['d_v_a']):.2e}  "
          f"d_w {rel_l2(wt.grad.cpu().numpy(), ref['d_w']):.2e}  d_gate_w {rel_l2(gw.grad.view(-1).cpu().numpy(), ref['d_gate_w']):.2e}", flush=True)
