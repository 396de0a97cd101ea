"""Backward diagnostics on the B200 box: runs coattention backward on a small case and prints errors vs the oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cosnet_b200 import coattention
from oracle import coattn_oracle as orc

n, h, w = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (1, 4, 5)))
bias = True
dev = torch.device("cuda:0")
v_a, v_b = orc.synthetic_features(300, n, h, w, 0.66)
W, g, b = orc.synthetic_weights(301, bias=bias)
rng = np.random.default_rng(5)
r_a = rng.standard_normal((n, 512, h, w), dtype=np.float32)
r_b = rng.standard_normal((n, 512, h, w), dtype=np.float32)
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
va = t(v_a).requires_grad_(True); vb = t(v_b); wt = t(W).requires_grad_(True)
gw = t(g).view(1, -1, 1, 1).requires_grad_(True); gb = t(b).requires_grad_(True)
cat_a, cat_b = coattention(va, vb, wt, gw, gb, bool(int(os.environ.get("BF16", "0"))))
loss = (cat_a * t(r_a)).sum() + (cat_b * t(r_b)).sum()
torch.cuda.synchronize(); print("forward ok", flush=True)
loss.backward()
torch.cuda.synchronize(); print("backward ok", flush=True)
ref = orc.coattention_grads(v_a, v_b, W, g, b, r_a, r_b)
def rel(x, r):
    x = np.asarray(x, np.float64); r = np.asarray(r, np.float64)
    return float(np.linalg.norm(x - r) / max(np.linalg.norm(r), 1e-30))
print("d_v_a", rel(va.grad.cpu().numpy(), ref["d_v_a"]))
print("d_w", rel(wt.grad.cpu().numpy(), ref["d_w"]))
print("d_gate_w", rel(gw.grad.view(-1).cpu().numpy(), ref["d_gate_w"]))
print("d_gate_b", float(gb.grad[0]), float(ref["d_gate_b"]))
