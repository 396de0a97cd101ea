"""Runs forward + backward (8 pairs, 60x60x256, RGB-style: both concat gradients) through a trace build of the library:

    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --shared -Xcompiler -fPIC -DCOATTN_TRACE_BWD \
         -o build_trace/libcoattn_trace.so cosnet_b200/csrc/coattn_api.cu
    COATTN_B200_LIB=build_trace/libcoattn_trace.so python tools/bwd_trace.py

CTA 0 of `bwd_tile_kernel` prints where its producer, MMA issuer and epilogue warp 0 spent their cycles."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosnet_b200 import coattention
dev = torch.device("cuda:0")
n, h, w = 8, 60, 60
g = torch.Generator(device=dev); g.manual_seed(1)
va = (torch.randn(n,256,h,w, generator=g, device=dev)*0.66).requires_grad_(True)
vb = torch.randn(n,256,h,w, generator=g, device=dev)*0.66
wt = ((torch.rand(256,256, generator=g, device=dev)*2-1)/16).requires_grad_(True)
gw = (torch.randn(1,256,1,1, generator=g, device=dev)*0.01).requires_grad_(True)
for it in range(2):
    ca, cb = coattention(va, vb, wt, gw, None)
    loss = (ca.sum() + cb.sum()) * 1e-3
    loss.backward()
    torch.cuda.synchronize()
print("done")
