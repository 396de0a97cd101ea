#!/usr/bin/env python
"""Full raa model (ResNet-101 RGB + ResNet-50 depth, 142 M parameters, random init) training step at 473x473 with
`cosnet_b200.train_step.TrainStep` on N ranks (one process per GPU, frame pairs sharded, bucketed NCCL all-reduce launched from
backward hooks -- the replacement of nn.DataParallel, train.py:491-496).  Prints ms per step (max over ranks), with the
overlapped buckets and with ONE flat all-reduce after the backward (round 1) for comparison.

    python tools/train_scale_probe.py                       # 1 GPU
    torchrun --nproc-per-node 8 tools/train_scale_probe.py  # 8 GPUs, same per-GPU batch
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from cosnet_b200.backbone import Bottleneck
from cosnet_b200.rgbd_segmentation_raa import RGBDSegmentation_RAA
from cosnet_b200 import train_step as ts

B = int(os.environ.get("PAIRS", "2")); S = int(os.environ.get("SIZE", "473")); STEPS = int(os.environ.get("STEPS", "8"))
torch.manual_seed(1234)
model = RGBDSegmentation_RAA(Bottleneck, [3, 4, 23, 3], [3, 4, 6, 3], num_classes=1).to(dev).train()
g = torch.Generator(device=dev); g.manual_seed(rank)
x = [torch.randn(B, c, S, S, device=dev, generator=g) for c in (3, 3, 1, 1)]
gt = (torch.rand(B, 1, S, S, device=dev, generator=g) > 0.5).float()


def timed(step):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(STEPS):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / STEPS
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    return ms


out = {"workload": "train_scale_probe", "n_gpus": world, "pairs_per_gpu": B, "input": [S, S], "params": sum(p.numel() for p in model.parameters())}
stepper = ts.TrainStep(model, learning_rate=1e-4, max_iter=1000)
out["bucketed_overlapped_ms"] = timed(lambda: stepper(x[0], x[1], x[2], x[3], gt, gt))
if stepper.buckets is not None:
    out["buckets"] = len(stepper.buckets.buckets); out["buckets_launched_during_backward"] = stepper.buckets.launched_in_backward
    stepper.buckets.remove()
    # round-1 arrangement: one flat all-reduce after the whole backward
    opt = stepper.optimizer
    for gr in opt.param_groups:
        for p in gr["params"]:
            p.grad = None

    def flat_step():
        opt.zero_grad(set_to_none=True)
        p1, p2, _ = model(x[0], x[1], x[2], x[3])
        ts.segmentation_loss(p1, gt, p2, gt).backward()
        ts.allreduce_gradients(p for gr in opt.param_groups for p in gr["params"])
        opt.step()
    out["flat_allreduce_after_backward_ms"] = timed(flat_step)
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
