#!/usr/bin/env python
"""What switching buys for the WHOLE raa model on one B200 (BASELINE cfg 1 shape on the GPU): the drop-in
RGBDSegmentation_RAA (ResNet-101 RGB + ResNet-50 depth encoders on cuDNN, random init) at 473x473 input, timed with
  (a) the sm_100a co-attention of this repo, and
  (b) the reference's op sequence (rgbd_segmentation_RAA.py:154-187: transpose, Linear, bmm, two softmaxes, two bmms,
      1x1 gate conv, sigmoid, concat) run eagerly by PyTorch on the same GPU, injected as `coattention_impl`.
Reports ms per forward (eval) and per training step (forward + backward, counterpart frozen) and the peak memory.

    python tools/model_probe.py [--batch 1 4] [--train-batch 4]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.nn.functional as F


def eager_coattention(v_a, v_b, weight, gate_weight, gate_bias, gated_only=False):
    n, c, h, w = v_a.shape
    a, b = v_a.view(n, c, h * w), v_b.view(n, c, h * w)
    q = F.linear(a.transpose(1, 2).contiguous(), weight)                 # :158-159
    s = torch.bmm(q, b)                                                  # :160
    s_row = F.softmax(s.clone(), dim=1)                                  # :164
    s_col = F.softmax(s.transpose(1, 2), dim=1)                          # :165
    z_b = torch.bmm(a, s_row).view(n, c, h, w)                           # :169
    z_a = torch.bmm(b, s_col).view(n, c, h, w)                           # :170
    gw = gate_weight.view(1, c, 1, 1)
    m_a = torch.sigmoid(F.conv2d(z_a, gw, gate_bias))                    # :175-177
    with torch.no_grad():
        m_b = torch.sigmoid(F.conv2d(z_b, gw, gate_bias))                # :178-182
    if gated_only:
        return z_a * m_a, z_b * m_b
    return torch.cat([z_a * m_a, v_a], 1), torch.cat([z_b * m_b, v_b], 1)    # :183-187


def timed(fn, warmup, steps):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, torch.cuda.max_memory_allocated() / 2 ** 30


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, nargs="+", default=[1, 4])
    ap.add_argument("--train-batch", type=int, default=4)      # the reference trains at batch 4
    ap.add_argument("--size", type=int, default=473)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    from cosnet_b200.backbone import Bottleneck
    from cosnet_b200.coattention import coattention
    from cosnet_b200.rgbd_segmentation_raa import RGBDSegmentation_RAA
    dev = torch.device("cuda:0")
    torch.manual_seed(1234)
    model = RGBDSegmentation_RAA(Bottleneck, [3, 4, 23, 3], [3, 4, 6, 3], num_classes=1).to(dev)
    s = args.size

    def inputs(n):
        return (torch.randn(n, 3, s, s, device=dev), torch.randn(n, 3, s, s, device=dev),
                torch.randn(n, 1, s, s, device=dev), torch.randn(n, 1, s, s, device=dev))

    for n in args.batch:
        x = inputs(n)
        model.eval()
        row = {"workload": "model_probe", "mode": "eval forward", "pairs": n, "input": [s, s]}
        for name, impl in (("b200", coattention), ("eager_reference_ops", eager_coattention)):
            model.coattention_impl = impl

            def fwd():
                with torch.no_grad():
                    return model(*x)
            ms, gib = timed(fwd, 3, args.steps)
            row[name] = {"ms_per_forward": ms, "ms_per_frame_pair": ms / n, "peak_gib": gib}
        print(json.dumps(row), flush=True)

    # the whole eval forward as one CUDA graph (cosnet_b200.graphed.GraphedEvalModel): what the host-bound batch-1 case costs
    # once the ~700 launches are replayed instead of issued
    from cosnet_b200.graphed import GraphedEvalModel
    model.coattention_impl = coattention
    for n in args.batch:
        x = inputs(n)
        row = {"workload": "model_probe", "mode": "eval forward, whole-model CUDA graph", "pairs": n, "input": [s, s]}
        for fuse in (True, False):
            model.fuse_eval_path = fuse
            gm = GraphedEvalModel(model, *x)
            with torch.no_grad():
                want = model(*x)
            got = gm(*x)
            torch.cuda.synchronize()
            err = max(float((a - b).abs().max()) for a, b in zip(got, want))
            ms, gib = timed(lambda: gm(*x), 3, args.steps)
            row["fused_eval_path" if fuse else "plain_eval_path"] = {"ms_per_forward": ms, "ms_per_frame_pair": ms / n,
                                                                      "max_abs_diff_vs_eager": err}
            del gm
        model.fuse_eval_path = True
        print(json.dumps(row), flush=True)

    # the same model in half precision (`model.half()`): fp16 encoders on cuDNN hand fp16 features to coattn_forward16
    import copy
    half = copy.deepcopy(model).half().eval()
    half.coattention_impl = coattention
    for n in args.batch:
        x = tuple(t.half() for t in inputs(n))

        def fwd16():
            with torch.no_grad():
                return half(*x)
        ms, gib = timed(fwd16, 3, args.steps)
        print(json.dumps({"workload": "model_probe", "mode": "eval forward, model.half()", "pairs": n, "input": [s, s],
                          "b200": {"ms_per_forward": ms, "ms_per_frame_pair": ms / n, "peak_gib": gib}}), flush=True)
    del half

    n = args.train_batch
    x = inputs(n)
    gt = (torch.rand(n, 1, s, s, device=dev) > 0.5).float()
    model.train()
    row = {"workload": "model_probe", "mode": "train step (forward + backward, no optimiser)", "pairs": n, "input": [s, s]}
    for name, impl in (("b200", coattention), ("eager_reference_ops", eager_coattention)):
        model.coattention_impl = impl

        def step():
            model.zero_grad(set_to_none=True)
            x1, x2, _ = model(*x)
            loss = F.binary_cross_entropy(x1, gt) + F.binary_cross_entropy(x2, gt)
            loss.backward()
        ms, gib = timed(step, 2, max(3, args.steps // 2))
        row[name] = {"ms_per_step": ms, "ms_per_frame_pair": ms / n, "peak_gib": gib}
    print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
