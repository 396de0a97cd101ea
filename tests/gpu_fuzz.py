"""Randomised shape / scale fuzzing of the CUDA forward and backward against the oracle (diagnostic, B200 box).
    python tests/gpu_fuzz.py [cases] [seed]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cosnet_b200 import coattention
from oracle import coattn_oracle as orc

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dev = torch.device("cuda:0")
rel = lambda x, r: float(np.linalg.norm(np.asarray(x, np.float64) - r) / max(np.linalg.norm(r), 1e-30))
worst = {"fwd": 0.0, "d_v_a": 0.0, "d_v_b": 0.0, "d_w": 0.0}
bad = 0
for k in range(cases):
    n = int(rng.integers(1, 4)); h = int(rng.integers(2, 45)); w = int(rng.integers(2, 45))
    sigma = float(rng.choice([0.25, 0.66, 1.0])); bias = bool(rng.integers(0, 2)); both = bool(rng.integers(0, 2))
    v_a, v_b = orc.synthetic_features(int(rng.integers(1 << 30)), n, h, w, sigma)
    W, g, b = orc.synthetic_weights(int(rng.integers(1 << 30)), bias=bias)
    r_a = rng.standard_normal((n, 512, h, w), dtype=np.float32); r_b = rng.standard_normal((n, 512, h, w), dtype=np.float32)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    va = t(v_a).requires_grad_(True); vb = t(v_b).requires_grad_(both); wt = t(W).requires_grad_(True)
    gw = t(g).requires_grad_(True); gb = None if b is None else t(b).requires_grad_(True)
    ca, cb = coattention(va, vb, wt, gw, gb)
    ((ca * t(r_a)).sum() + (cb * t(r_b)).sum()).backward()
    torch.cuda.synchronize()
    ref = orc.coattention(v_a, v_b, W, g, b)
    gr = orc.coattention_grads(v_a, v_b, W, g, b, r_a, r_b, counterpart_grad=both)
    e = {"fwd": max(rel(ca.detach().cpu().numpy(), ref["cat_a"]), rel(cb.detach().cpu().numpy(), ref["cat_b"])),
         "d_v_a": rel(va.grad.cpu().numpy(), gr["d_v_a"]), "d_w": rel(wt.grad.cpu().numpy(), gr["d_w"]),
         "d_v_b": rel(vb.grad.cpu().numpy(), gr["d_v_b"]) if both else 0.0}
    flag = e["fwd"] > 1e-3 or max(e["d_v_a"], e["d_v_b"], e["d_w"]) > 1.5e-2
    bad += flag
    for kk in worst: worst[kk] = max(worst[kk], e[kk])
    print(f"case {k:3d} n={n} {h}x{w} L={h*w} sigma={sigma} bias={bias} both={both}: " + " ".join(f"{kk}={vv:.2e}" for kk, vv in e.items()) + ("  <-- CHECK" if flag else ""), flush=True)
print("worst:", worst, "flagged:", bad)
