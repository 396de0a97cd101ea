"""HostPipeline sweep with STREAMED calls (join=False), gated-only contract, headline shape: pairs per chunk x streams -> frame-pairs/s."""
import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from cosnet_b200.coattention import HostPipeline
dev = torch.device("cuda:0")
n, C, H, W = 32, 256, 60, 60
g = torch.Generator(device=dev); g.manual_seed(1)
hin = [(torch.randn(n, C, H, W, generator=g, device=dev) * 0.66).cpu().pin_memory() for _ in range(4)]
w = (torch.rand(C, C, generator=g, device=dev) * 2 - 1) / 16
gw = torch.randn(C, generator=g, device=dev) * 0.01
hout = [torch.empty(n, C, H, W).pin_memory() for _ in range(4)]
for chunk in (2, 4, 8, 16, 32):
    for slots in (2, 3, 4, 6):
        pipe = HostPipeline(n, C, H, W, chunk=chunk, slots=slots, device=dev, gated_only=True)
        def step():
            pipe(hin[0], hin[1], w, gw, None, hout[0], hout[1], join=False)
            pipe(hin[2], hin[3], w, gw, None, hout[2], hout[3], join=False)
        step(); pipe.join(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(8): step()
        pipe.join(); torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 8
        print(f"chunk={chunk} slots={slots}: {dt*1e3:.2f} ms per step -> {n/dt:.0f} pairs/s", flush=True)
        del pipe
