#!/usr/bin/env python
"""Secondary workloads of BASELINE.json (configs 3-5) on the co-attention hot path, synthetic features.

    python tools/bench_extra.py --workload {hd|inference|train} [--steps K --warmup W]
    torchrun --nproc-per-node N tools/bench_extra.py --workload ...        (pair / query sharded, weak scaling)

  hd         cfg 3: 480x854 input -> 61x107x256 features (what the reference really produces), batch 16 per GPU
  inference  cfg 4: test.py-style, each query co-attended with 5 reference frames (frame-A outputs only),
             480x640 input -> 61x81x256 features, 8 queries (40 pairs) per GPU
  latency / latency_split   one pair per step: default path vs COATTN_FLAG_SPLIT_KEYS
  io16 / io16_bf16   cfg 2 through the 16-bit feature interface (coattn_forward16); inference16: cfg 4 likewise
  eager / eager_bf16 / sdpa   secondary comparators: the reference's op sequence and PyTorch's fused attention on the same GPU
  train_abi  the same through the C ABI only (no autograd, allocator or stand-in loss inside the timed region)
  train      cfg 5: forward + hand-written backward of both modalities, 8 pairs per GPU, NCCL all-reduce of the
             hot-path gradients (W, gate: 131 585 floats per step)
Prints one JSON line (rank 0).  Times are CUDA events, max over ranks.
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
    ap.add_argument("--workload", required=True, choices=["hd", "inference", "gated", "gated16", "latency", "latency_split", "latency_graph", "io16", "io16_bf16", "inference16", "train", "train_abi", "eager", "eager_bf16", "sdpa"])
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--pairs", type=int, default=0, help="hd: pairs per GPU instead of 16 (cfg 3 strong-scaled over 8 GPUs = 2)")
    ap.add_argument("--pair-call", action="store_true", help="latency + --two-streams: both modality calls through coattention_pair")
    ap.add_argument("--two-streams", action="store_true", help="hd / inference: the depth modality call runs on a second "
                    "stream, so its CTA pairs back-fill the tail wave of the RGB attend kernel")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from cosnet_b200 import coattention
    from cosnet_b200.coattention import coattention_forward_raw
    C = 256
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)

    def feats(n, h, w, grad=False):
        x = torch.randn((n, C, h, w), generator=g, device=dev)
        x = F.prelu(x, torch.tensor([0.25], device=dev)) * 0.66
        return x.requires_grad_(grad)
    k = 1.0 / 16
    W = [((torch.rand((C, C), generator=g, device=dev) * 2 - 1) * k) for _ in range(2)]
    G = [torch.randn((C,), generator=g, device=dev) * 0.01 for _ in range(2)]
    Bd = (torch.rand((1,), generator=g, device=dev) * 2 - 1) * k

    if args.workload == "hd":
        n, h, w = (args.pairs or 16), 61, 107
        va, vb, da, db = (feats(n, h, w) for _ in range(4))
        side = torch.cuda.Stream(dev)
        def step():
            if args.two_streams:
                cur = torch.cuda.current_stream(dev)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    coattention_forward_raw(da, db, W[1], G[1], Bd, want_z=False)
                coattention_forward_raw(va, vb, W[0], G[0], None, want_z=False)
                cur.wait_stream(side)
                return
            coattention_forward_raw(va, vb, W[0], G[0], None, want_z=False)
            coattention_forward_raw(da, db, W[1], G[1], Bd, want_z=False)
        pairs = n
        desc = f"co-attention at 480x854 input (61x107x256 features), batch {n} per GPU" + (", modalities on two streams" if args.two_streams else "")
    elif args.workload == "inference":
        qn, r, h, w = 8, 5, 61, 81
        va, da = feats(qn, h, w), feats(qn, h, w)
        vb, db = feats(qn * r, h, w), feats(qn * r, h, w)
        from cosnet_b200.coattention import coattention_queries_raw
        def step():     # the query side (cast, Q = W V_a) is prepared once per query: coattn_forward_queries
            coattention_queries_raw(va, vb, W[0], G[0], None, refs=r)
            coattention_queries_raw(da, db, W[1], G[1], Bd, refs=r)
        pairs = qn * r
        desc = "test.py-style inference: 8 queries x 5 references per GPU, 61x81x256 features, frame-A outputs only"
    elif args.workload in ("latency", "latency_split"):
        # cfg 1 regime: ONE frame pair (60x60x256), RGB + depth call, default path vs COATTN_FLAG_SPLIT_KEYS
        n, h, w = 1, 60, 60
        va, vb, da, db = (feats(n, h, w) for _ in range(4))
        sk = args.workload == "latency_split"
        from cosnet_b200.coattention import run_modalities
        def step():
            if args.two_streams and args.pair_call:      # both calls through coattention_pair (streams by handle)
                from cosnet_b200 import coattention_pair
                return coattention_pair((va, vb, W[0], G[0], None), (da, db, W[1], G[1], Bd))
            if args.two_streams:      # what the drop-in module's eval forward does at this size
                return run_modalities(lambda: coattention_forward_raw(va, vb, W[0], G[0], None, want_z=False, want_lse=False, split_keys=sk),
                                      lambda: coattention_forward_raw(da, db, W[1], G[1], Bd, want_z=False, want_lse=False, split_keys=sk),
                                      (da, db), True)
            coattention_forward_raw(va, vb, W[0], G[0], None, want_z=False, want_lse=False, split_keys=sk)
            coattention_forward_raw(da, db, W[1], G[1], Bd, want_z=False, want_lse=False, split_keys=sk)
        pairs = n
        desc = ("one frame pair per step (60x60x256, RGB + depth call)" + (", COATTN_FLAG_SPLIT_KEYS" if sk else ", default path")
                + (", eager calls on two streams" if args.two_streams else "") + (" through coattention_pair" if args.pair_call else ""))
    elif args.workload == "latency_graph":
        # the same single pair through GraphedCoAttention: one graph launch per step (host time included, like `latency`)
        from cosnet_b200.graphed import GraphedCoAttention
        n, h, w = 1, 60, 60
        gr = GraphedCoAttention(n, h, w, (W[0], G[0], None), (W[1], G[1], Bd), device=dev)
        for buf in (gr.v_a, gr.v_b, gr.d_a, gr.d_b):
            buf.copy_(feats(n, h, w))
        step = gr.replay
        pairs = n
        desc = "one frame pair per step (60x60x256, RGB + depth call), GraphedCoAttention.replay() (CUDA graph, modalities on two streams)"
    elif args.workload in ("gated", "gated16"):
        # cfg 2 with the gated-only contract (the consumer splits the reduce conv, the concat and its passthrough
        # copy never exist): fp32 features, or fp16 features through coattn_forward16
        from cosnet_b200.coattention import coattention_forward16_raw
        n, h, w = 32, 60, 60
        half = args.workload == "gated16"
        va, vb, da, db = ((feats(n, h, w).half() if half else feats(n, h, w)) for _ in range(4))
        def step():
            if half:
                coattention_forward16_raw(va, vb, W[0], G[0], None, gated_only=True)
                coattention_forward16_raw(da, db, W[1], G[1], Bd, gated_only=True)
            else:
                coattention_forward_raw(va, vb, W[0], G[0], None, want_z=False, gated_only=True)
                coattention_forward_raw(da, db, W[1], G[1], Bd, want_z=False, gated_only=True)
        pairs = n
        desc = ("co-attention module alone, 60x60x256, batch 32 per GPU, gated half only (no concat / passthrough copy), "
                + ("fp16 features in and out" if half else "fp32 features"))
    elif args.workload in ("io16", "io16_bf16"):
        # the headline shape through the 16-bit feature interface (coattn_forward16): fp16 (or bf16) features in and out,
        # read in place by TMA -- no cast pass, half the concat bytes
        from cosnet_b200.coattention import coattention_forward16_raw
        n, h, w = 32, 60, 60
        dt = torch.float16 if args.workload == "io16" else torch.bfloat16
        va, vb, da, db = (feats(n, h, w).to(dt) for _ in range(4))
        def step():
            coattention_forward16_raw(va, vb, W[0], G[0], None)
            coattention_forward16_raw(da, db, W[1], G[1], Bd)
        pairs = n
        desc = f"co-attention module alone, 60x60x256 feature pairs, batch 32 per GPU, {dt} features in and out (coattn_forward16)"
    elif args.workload == "inference16":
        from cosnet_b200.coattention import coattention_forward16_raw
        qn, r, h, w = 8, 5, 61, 81
        va, da = feats(qn, h, w).half(), feats(qn, h, w).half()
        vb, db = feats(qn * r, h, w).half(), feats(qn * r, h, w).half()
        def step():
            coattention_forward16_raw(va, vb, W[0], G[0], None, refs=r)
            coattention_forward16_raw(da, db, W[1], G[1], Bd, refs=r)
        pairs = qn * r
        desc = "test.py-style inference, fp16 features in and out: 8 queries x 5 references per GPU, 61x81x256 (L = 4941, copy path), frame-A outputs only"
    elif args.workload in ("eager", "eager_bf16", "sdpa"):
        # secondary GPU comparators at the headline shape: the reference's own op sequence run eagerly on the
        # B200 (fp32 as written, or bf16), and two scaled_dot_product_attention calls (scale 1) + the gate epilogue
        n, h, w = 32, 60, 60
        dt = torch.float32 if args.workload == "eager" else torch.bfloat16
        va, vb, da, db = (feats(n, h, w).to(dt) for _ in range(4))
        Wd = [x.to(dt) for x in W]; Gd = [x.to(dt).view(1, C, 1, 1) for x in G]; Bdd = Bd.to(dt)

        def epilogue(z_a, z_b, v_a, v_b, gw, gb):
            m_a = torch.sigmoid(F.conv2d(z_a, gw, gb)); m_b = torch.sigmoid(F.conv2d(z_b, gw, gb))
            return torch.cat([z_a * m_a, v_a], 1), torch.cat([z_b * m_b, v_b], 1)

        def eager(v_a, v_b, wt, gw, gb):
            a, b = v_a.view(n, C, h * w), v_b.view(n, C, h * w)
            q = F.linear(a.transpose(1, 2).contiguous(), wt)
            s_ = torch.bmm(q, b)
            s_row = F.softmax(s_.clone(), dim=1)
            s_col = F.softmax(s_.transpose(1, 2), dim=1)
            z_b = torch.bmm(a, s_row).view(n, C, h, w)
            z_a = torch.bmm(b, s_col).view(n, C, h, w)
            return epilogue(z_a, z_b, v_a, v_b, gw, gb)

        def sdpa(v_a, v_b, wt, gw, gb):
            a, b = v_a.view(n, C, h * w).transpose(1, 2), v_b.view(n, C, h * w).transpose(1, 2)      # [n, L, C]
            q = F.linear(a, wt)
            z_a = F.scaled_dot_product_attention(q.unsqueeze(1), b.unsqueeze(1), b.unsqueeze(1), scale=1.0).squeeze(1)
            z_b = F.scaled_dot_product_attention(b.unsqueeze(1), q.unsqueeze(1), a.unsqueeze(1), scale=1.0).squeeze(1)
            z_a = z_a.transpose(1, 2).reshape(n, C, h, w); z_b = z_b.transpose(1, 2).reshape(n, C, h, w)
            return epilogue(z_a, z_b, v_a, v_b, gw, gb)

        fn = sdpa if args.workload == "sdpa" else eager
        def step():
            with torch.no_grad():
                fn(va, vb, Wd[0], Gd[0], None)
                fn(da, db, Wd[1], Gd[1], Bdd)
        pairs = n
        desc = {"eager": "reference op sequence (rgbd_segmentation_RAA.py:154-187) eagerly on the GPU, fp32, batch 32, 60x60x256",
                "eager_bf16": "reference op sequence eagerly on the GPU in bf16, batch 32, 60x60x256",
                "sdpa": "2 x F.scaled_dot_product_attention(scale=1) in bf16 + eager projection / gate / concat, batch 32, 60x60x256"}[args.workload]
    elif args.workload == "train_abi":
        # the same forward + backward as `train`, but through the C ABI with preallocated buffers: only the library's own
        # kernels are inside the timed region (no autograd, no allocator, no stand-in loss)
        from cosnet_b200 import _lib
        from cosnet_b200.coattention import backward_workspace_bytes, workspace_bytes
        lib = _lib.load()
        n, h, w = 8, 60, 60
        L = h * w
        va, da, vb, db = feats(n, h, w), feats(n, h, w), feats(n, h, w), feats(n, h, w)
        gws = [G[0].reshape(-1).contiguous(), G[1].reshape(-1).contiguous()]
        ra = torch.randn((n, 2 * C, h, w), generator=g, device=dev) * 1e-3
        rb = torch.randn((n, 2 * C, h, w), generator=g, device=dev) * 1e-3
        nb_f, nb_b = workspace_bytes(n, C, h, w), backward_workspace_bytes(n, C, h, w, False)
        ws = torch.empty(max(nb_f, nb_b) + 1024, dtype=torch.uint8, device=dev)
        wsp = (ws.data_ptr() + 1023) // 1024 * 1024
        st = torch.cuda.current_stream(dev).cuda_stream
        mods = []
        for (a, b, wt, gw, gb, has_b) in ((va, vb, W[0], gws[0], None, True), (da, db, W[1], gws[1], Bd, False)):
            mods.append(dict(a=a, b=b, w=wt, gw=gw, gb=gb, has_b=has_b,
                             ca=torch.empty((n, 2 * C, h, w), device=dev), cb=torch.empty((n, 2 * C, h, w), device=dev),
                             z=torch.empty((2, n, C, L), device=dev), lse=torch.empty((2, n, L), device=dev),
                             mask=torch.empty((2, n, L), device=dev), dva=torch.empty((n, C, h, w), device=dev),
                             dw=torch.empty((C, C), device=dev), dgw=torch.empty((C,), device=dev), dgb=torch.empty((1,), device=dev)))
        P = lambda t: None if t is None else t.data_ptr()
        ws2 = torch.empty(max(nb_f, nb_b) + 1024, dtype=torch.uint8, device=dev)
        wsp2 = (ws2.data_ptr() + 1023) // 1024 * 1024
        side2 = torch.cuda.Stream(dev)

        def modality(m, wp, s_):      # forward + backward of one modality on one stream
            _lib.check(lib.coattn_forward(P(m["a"]), P(m["b"]), P(m["w"]), P(m["gw"]), P(m["gb"]), P(m["ca"]), P(m["cb"]),
                                          P(m["z"]), P(m["lse"]), P(m["mask"]), wp, nb_f, n, C, h, w, 0, s_), "coattn_forward")
            _lib.check(lib.coattn_backward(P(m["a"]), P(m["b"]), P(m["w"]), P(m["gw"]), P(m["z"]), P(m["lse"]), P(m["mask"]),
                                           P(ra), P(rb) if m["has_b"] else None, P(m["dva"]), None, P(m["dw"]), P(m["dgw"]),
                                           P(m["dgb"]) if m["gb"] is not None else None, wp, nb_b, n, C, h, w, 0, s_),
                       "coattn_backward")

        def step_two_streams():       # the modalities are independent until the gradient all-reduce
            cur = torch.cuda.current_stream(dev)
            side2.wait_stream(cur)
            modality(mods[1], wsp2, side2.cuda_stream)
            modality(mods[0], wsp, st)
            cur.wait_stream(side2)
            if world > 1:
                dist.all_reduce(torch.cat([mods[0]["dw"].reshape(-1), mods[0]["dgw"], mods[1]["dw"].reshape(-1), mods[1]["dgw"], mods[1]["dgb"]]))

        def step():
            if args.two_streams:
                return step_two_streams()
            for m in mods:
                _lib.check(lib.coattn_forward(P(m["a"]), P(m["b"]), P(m["w"]), P(m["gw"]), P(m["gb"]), P(m["ca"]), P(m["cb"]),
                                              P(m["z"]), P(m["lse"]), P(m["mask"]), wsp, nb_f, n, C, h, w, 0, st), "coattn_forward")
            for m in mods:
                _lib.check(lib.coattn_backward(P(m["a"]), P(m["b"]), P(m["w"]), P(m["gw"]), P(m["z"]), P(m["lse"]), P(m["mask"]),
                                               P(ra), P(rb) if m["has_b"] else None, P(m["dva"]), None, P(m["dw"]), P(m["dgw"]),
                                               P(m["dgb"]) if m["gb"] is not None else None, wsp, nb_b, n, C, h, w, 0, st),
                           "coattn_backward")
            if world > 1:
                dist.all_reduce(torch.cat([mods[0]["dw"].reshape(-1), mods[0]["dgw"], mods[1]["dw"].reshape(-1), mods[1]["dgw"], mods[1]["dgb"]]))
        pairs = n
        desc = ("train step on the hot path through the C ABI: coattn_forward + coattn_backward of both modalities (RGB full, "
                "depth A-branch), 8 pairs per GPU, preallocated buffers, NCCL all-reduce of hot-path grads"
                + (", modalities on two streams" if args.two_streams else ""))
    else:
        n, h, w = 8, 60, 60
        va, da = feats(n, h, w, True), feats(n, h, w, True)
        vb, db = feats(n, h, w), feats(n, h, w)
        params = [W[0].requires_grad_(True), G[0].requires_grad_(True), W[1].requires_grad_(True), G[1].requires_grad_(True), Bd.requires_grad_(True)]
        ra = torch.randn((n, 2 * C, h, w), generator=g, device=dev); rb = torch.randn((n, 2 * C, h, w), generator=g, device=dev)
        from cosnet_b200.coattention import run_modalities
        def step():
            for p in params + [va, da]:
                p.grad = None
            if args.two_streams:      # autograd replays each branch's backward on the stream its forward ran on
                (ca, cb), (dca, dcb) = run_modalities(lambda: coattention(va, vb, params[0], params[1], None),
                                                      lambda: coattention(da, db, params[2], params[3], params[4]),
                                                      (da, db), True)
            else:
                ca, cb = coattention(va, vb, params[0], params[1], None)
                dca, dcb = coattention(da, db, params[2], params[3], params[4])
            # depth: the B branch is gradient dead in the reference (:240-247) -> only cat_a carries gradient
            loss = (ca * ra).sum() + (cb * rb).sum() + (dca * ra).sum()
            loss.backward()
            if world > 1:
                flat = torch.cat([p.grad.reshape(-1) for p in params])
                dist.all_reduce(flat)
        pairs = n
        desc = ("train step on the hot path: forward + backward (RGB full, depth A-branch), 8 pairs per GPU, NCCL all-reduce of hot-path grads"
                + (", modalities on two streams" if args.two_streams else ""))

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    if rank == 0:
        print(json.dumps({"metric": "co-attn frame-pairs/sec", "workload": args.workload, "config": desc, "value": pairs * world * args.steps / (ms * 1e-3),
                          "unit": "frame-pairs/s", "n_gpus": world, "steps": args.steps, "ms_per_step": ms / args.steps, "scaling": "weak",
                          "includes": ("only the library's kernels (C ABI, preallocated buffers)" if args.workload == "train_abi" else
                                       "PyTorch kernels only (none of this repository's)" if args.workload in ("eager", "eager_bf16", "sdpa") else
                                       "tensor allocations of the Python operator (caching allocator), all kernels of both modalities")}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
