"""Determinism / batch-invariance stress of the CUDA path (diagnostic, run on the B200 box)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cosnet_b200 import coattention_forward_raw as op
from oracle import coattn_oracle as orc

dev = torch.device("cuda:0")
n, h, w = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 60, 60
gen = torch.Generator(device=dev); gen.manual_seed(1)
x = torch.randn((2, n, 256, h, w), generator=gen, device=dev)
f = torch.where(x >= 0, x, 0.25 * x) * 0.66
W, g, b = (torch.from_numpy(t).to(dev) for t in orc.synthetic_weights(5, bias=True))
ref = [t.clone() for t in op(f[0], f[1], W, g, b)]
torch.cuda.synchronize()
names = ["cat_a", "cat_b", "z", "lse"]
for it in range(10):
    out = op(f[0], f[1], W, g, b)
    torch.cuda.synchronize()
    for nm, a, r in zip(names, out, ref):
        if not torch.equal(a, r):
            d = (a != r)
            idx = d.nonzero()
            print(f"iter {it}: {nm} differs in {int(d.sum())} elements; first idx {idx[0].tolist()} last {idx[-1].tolist()} maxabs {float((a-r).abs().max())}")
# chunked vs full
for lo in range(0, n, 4):
    out = op(f[0][lo:lo+4].contiguous(), f[1][lo:lo+4].contiguous(), W, g, b)
    torch.cuda.synchronize()
    for nm, a, r in zip(names[:2], out[:2], ref[:2]):
        if not torch.equal(a, r[lo:lo+4]):
            d = (a != r[lo:lo+4]); idx = d.nonzero()
            print(f"chunk {lo}: {nm} differs in {int(d.sum())} elements; first {idx[0].tolist()} last {idx[-1].tolist()} maxabs {float((a-r[lo:lo+4]).abs().max())}")
print("stress done")
