"""Sweep of the HostPipeline chunking (pairs per chunk x streams) at the headline shape; prints frame-pairs/s."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosnet_b200.coattention import HostPipeline
dev = torch.device("cuda:0")
n, C, H, W = 32, 256, 60, 60
g = torch.Generator(device=dev); g.manual_seed(1)
hin = [(torch.randn(n, C, H, W, generator=g, device=dev) * 0.66).cpu().pin_memory() for _ in range(2)]
w = (torch.rand(C, C, generator=g, device=dev) * 2 - 1) / 16
gw = torch.randn(C, generator=g, device=dev) * 0.01
for gated in (False, True):
    oc = C if gated else 2 * C
    hout = [torch.empty(n, oc, H, W).pin_memory() for _ in range(2)]
    for chunk in (2, 4, 8, 16):
        for slots in (2, 3, 4):
            pipe = HostPipeline(n, C, H, W, chunk=chunk, slots=slots, device=dev, gated_only=gated)
            pipe(hin[0], hin[1], w, gw, None, hout[0], hout[1]); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(6):
                pipe(hin[0], hin[1], w, gw, None, hout[0], hout[1])
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 6
            print(f"gated_only={gated} chunk={chunk} slots={slots}: {dt*1e3:.2f} ms per modality call -> {n/(2*dt):.0f} pairs/s", flush=True)
            del pipe
