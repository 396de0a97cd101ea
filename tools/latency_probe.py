#!/usr/bin/env python
"""GPU time of ONE frame pair (60x60x256, RGB + depth modality call) through the C ABI, replayed from a CUDA graph so that
host launch overhead is out of the picture: default path vs COATTN_FLAG_SPLIT_KEYS (and the 16-bit interface).

    python tools/latency_probe.py [--pairs 1] [--hw 60 60] [--replays 200]
Prints one JSON line per variant (CUDA events around `replays` back-to-back graph launches).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.nn.functional as F


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1)
    ap.add_argument("--hw", type=int, nargs=2, default=[60, 60])
    ap.add_argument("--replays", type=int, default=200)
    args = ap.parse_args()
    from cosnet_b200 import _lib
    from cosnet_b200.coattention import workspace_bytes
    lib = _lib.load()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    n, (h, w), C = args.pairs, args.hw, 256
    g = torch.Generator(device=dev); g.manual_seed(1234)
    feats = lambda: F.prelu(torch.randn((n, C, h, w), generator=g, device=dev), torch.tensor([0.25], device=dev)) * 0.66
    va, vb, da, db = feats(), feats(), feats(), feats()
    k = 1.0 / 16
    W = [((torch.rand((C, C), generator=g, device=dev) * 2 - 1) * k) for _ in range(2)]
    G = [torch.randn((C,), generator=g, device=dev) * 0.01 for _ in range(2)]
    Bd = (torch.rand((1,), generator=g, device=dev) * 2 - 1) * k
    nbytes = workspace_bytes(n, C, h, w)
    ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
    wsp = (ws.data_ptr() + 1023) // 1024 * 1024
    out32 = [torch.empty((n, 2 * C, h, w), device=dev) for _ in range(4)]
    out16 = [torch.empty((n, 2 * C, h, w), device=dev, dtype=torch.float16) for _ in range(4)]
    f16 = [t.half() for t in (va, vb, da, db)]

    def fwd32(flags):
        def call(st):
            _lib.check(lib.coattn_forward(va.data_ptr(), vb.data_ptr(), W[0].data_ptr(), G[0].data_ptr(), None, out32[0].data_ptr(),
                                          out32[1].data_ptr(), None, None, None, wsp, nbytes, n, C, h, w, flags, st), "fwd")
            _lib.check(lib.coattn_forward(da.data_ptr(), db.data_ptr(), W[1].data_ptr(), G[1].data_ptr(), Bd.data_ptr(),
                                          out32[2].data_ptr(), out32[3].data_ptr(), None, None, None, wsp, nbytes, n, C, h, w, flags,
                                          st), "fwd")
        return call

    side = torch.cuda.Stream(dev)
    ws2 = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)      # the modalities run concurrently: own workspaces
    wsp2 = (ws2.data_ptr() + 1023) // 1024 * 1024

    def fwd32_two_streams(flags):
        def call(st):
            cur = torch.cuda.current_stream(dev)
            side.wait_stream(cur)
            _lib.check(lib.coattn_forward(va.data_ptr(), vb.data_ptr(), W[0].data_ptr(), G[0].data_ptr(), None, out32[0].data_ptr(),
                                          out32[1].data_ptr(), None, None, None, wsp, nbytes, n, C, h, w, flags, st), "fwd")
            _lib.check(lib.coattn_forward(da.data_ptr(), db.data_ptr(), W[1].data_ptr(), G[1].data_ptr(), Bd.data_ptr(),
                                          out32[2].data_ptr(), out32[3].data_ptr(), None, None, None, wsp2, nbytes, n, C, h, w, flags,
                                          side.cuda_stream), "fwd")
            cur.wait_stream(side)
        return call

    def fwd16(st):
        _lib.check(lib.coattn_forward16(f16[0].data_ptr(), f16[1].data_ptr(), W[0].data_ptr(), G[0].data_ptr(), None,
                                        out16[0].data_ptr(), out16[1].data_ptr(), None, None, wsp, nbytes, n, 1, C, h, w, 0, st), "fwd16")
        _lib.check(lib.coattn_forward16(f16[2].data_ptr(), f16[3].data_ptr(), W[1].data_ptr(), G[1].data_ptr(), Bd.data_ptr(),
                                        out16[2].data_ptr(), out16[3].data_ptr(), None, None, wsp, nbytes, n, 1, C, h, w, 0, st), "fwd16")

    # host cost of issuing one frame pair through the C ABI (no graph, no Python operator): the GPU is allowed to fall behind
    import time
    call = fwd32(0)
    st0 = torch.cuda.current_stream(dev).cuda_stream
    for _ in range(20):
        call(st0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        call(st0)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(json.dumps({"workload": "latency_probe", "variant": "host time to issue the two coattn_forward calls (ctypes)",
                      "pairs": n, "feat_hw": [h, w], "us_per_frame_pair_host": (t1 - t0) / 200 * 1e6}), flush=True)

    for name, call in (("default path (coattn_forward)", fwd32(0)),
                       ("COATTN_FLAG_SPLIT_KEYS", fwd32(_lib.FLAG_SPLIT_KEYS)),
                       ("16-bit interface (coattn_forward16)", fwd16),
                       ("default path, RGB and depth calls on two streams", fwd32_two_streams(0)),
                       ("COATTN_FLAG_SPLIT_KEYS, RGB and depth calls on two streams", fwd32_two_streams(_lib.FLAG_SPLIT_KEYS))):
        warm = torch.cuda.Stream(dev)
        warm.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(warm):
            call(warm.cuda_stream)
        torch.cuda.current_stream(dev).wait_stream(warm)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            call(torch.cuda.current_stream(dev).cuda_stream)
        for _ in range(10):
            graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.replays):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.replays
        print(json.dumps({"workload": "latency_probe", "variant": name, "pairs": n, "feat_hw": [h, w],
                          "ms_per_step": ms, "us_per_frame_pair": ms * 1e3 / n, "replays": args.replays,
                          "how": "CUDA graph of the RGB + depth modality calls, back-to-back replays, CUDA events"}), flush=True)


if __name__ == "__main__":
    main()
