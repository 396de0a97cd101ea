"""Host logic of the train-step harness (reference train.py:161-216, :538-602), CPU only, incl. a 2-rank gloo run."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from cosnet_b200 import train_step as ts
from cosnet_b200.backbone import Bottleneck
from cosnet_b200.rgbd_segmentation_raa import RGBDSegmentation_RAA


def reference_bce(pred, label):
    """train.py:176-204 written out literally (weight tensor filled with total/num_pos)."""
    num_pos = int((label >= 0.5).int().sum().item())
    if num_pos == 0:
        return torch.nn.BCELoss()(pred, label)
    total = label.shape[0] * label.shape[2] * label.shape[3]
    weight = torch.full_like(label, total / num_pos)
    return torch.nn.BCELoss(weight=weight)(pred, label)


def test_loss_matches_reference_formula():
    g = torch.Generator().manual_seed(0)
    pred = torch.rand(2, 1, 9, 7, generator=g).clamp(1e-3, 1 - 1e-3)
    gt = (torch.rand(2, 1, 9, 7, generator=g) > 0.7).float()
    want = reference_bce(pred, gt) + 0.8 * F.l1_loss(pred, gt)
    assert torch.allclose(ts.segmentation_loss(pred, gt), want, rtol=1e-6)
    both = ts.segmentation_loss(pred, gt, pred.flip(0), gt.flip(0))
    assert torch.allclose(both, 2 * want, rtol=1e-6)
    empty = torch.zeros_like(gt)     # "empty GT" branch: plain BCE
    assert torch.allclose(ts.weighted_bce(pred, empty), torch.nn.BCELoss()(pred, empty), rtol=1e-6)


def test_poly_schedule_and_group_rates():
    assert ts.lr_poly(1e-3, 0, 100, 0.9, 0) == pytest.approx(1e-3)
    assert ts.lr_poly(1e-3, 50, 100, 0.9, 3) == pytest.approx(1e-3 * 0.5 ** 0.9)
    assert ts.lr_poly(1e-3, 50, 100, 0.9, 6) == pytest.approx(0.5e-3 * 0.5 ** 0.9)
    model = RGBDSegmentation_RAA(Bottleneck, [1, 1, 1, 1], [1, 1, 1, 1], num_classes=1)
    opt = ts.make_optimizer(model, 1e-3)
    lr = ts.adjust_learning_rate(opt, 1e-3, 10, 0, 100)
    assert opt.param_groups[0]["lr"] == pytest.approx(0.01 * lr) and opt.param_groups[1]["lr"] == pytest.approx(10 * lr)
    # the two groups are disjoint and together cover every trainable parameter except encoder.main_classifier,
    # which the reference's group builder also leaves out (get_params("encoder") returns the whole encoder: included)
    ids0 = {id(p) for p in opt.param_groups[0]["params"]}
    ids1 = {id(p) for p in opt.param_groups[1]["params"]}
    assert not (ids0 & ids1)
    trainable = {id(p) for p in model.parameters() if p.requires_grad}
    assert ids0 | ids1 == trainable
    hot = {id(model.rgb_similarity_weights.weight), id(model.gate.weight), id(model.depth_similarity_weights.weight),
           id(model.depth_gate.weight), id(model.depth_gate.bias)}
    assert hot <= ids1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
        x = torch.randn(8, 6)
        y = torch.randn(8, 1)
        # single-process gradient on the whole batch (mean loss)
        ref = [g.clone() for g in torch.autograd.grad(F.mse_loss(net(x), y), list(net.parameters()))]
        # sharded: each rank takes its contiguous half; mean of per-shard mean losses == full mean (equal shard sizes)
        from cosnet_b200.pair_batcher import shard_range
        s, c = shard_range(8, world, rank)
        F.mse_loss(net(x[s:s + c]), y[s:s + c]).backward()
        ts.allreduce_gradients(net.parameters())
        for p, r in zip(net.parameters(), ref):
            assert torch.allclose(p.grad, r, atol=1e-6), (p.grad - r).abs().max()
    finally:
        dist.destroy_process_group()


def _bucket_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(6, 32), torch.nn.Tanh(), torch.nn.Linear(32, 16), torch.nn.Tanh(),
                                  torch.nn.Linear(16, 1))
        unused = torch.nn.Parameter(torch.ones(7))          # never reaches the loss: its bucket is flushed by finish()
        x, y = torch.randn(8, 6), torch.randn(8, 1)
        ref = [g.clone() for g in torch.autograd.grad(F.mse_loss(net(x), y), list(net.parameters()))]
        params = list(net.parameters()) + [unused]
        gb = ts.GradientBuckets(params, bucket_bytes=600, first_bucket_bytes=100)     # several buckets
        assert len(gb.buckets) >= 3
        from cosnet_b200.pair_batcher import shard_range
        s, c = shard_range(8, world, rank)
        for _ in range(2):                                  # second step: zero() keeps the views, results identical
            gb.zero()
            F.mse_loss(net(x[s:s + c]), y[s:s + c]).backward()
            assert gb.launched_in_backward >= len(gb.buckets) - 1     # all but the unused parameter's bucket overlapped
            gb.finish()
            for p, r in zip(net.parameters(), ref):
                assert torch.allclose(p.grad, r, atol=1e-6), (p.grad - r).abs().max()
            assert unused.grad is None          # no gradient on any rank: stays None (the optimiser skips it)
            # the bucket holds the same averages that were unpacked into the gradients
            b0 = gb.buckets[gb._bucket_of[id(net[0].weight)]]
            v0 = b0["views"][[id(q) for q in b0["params"]].index(id(net[0].weight))]
            assert torch.equal(v0, net[0].weight.grad)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_bucketed_overlapped_allreduce_equals_single_process():
    mp.spawn(_bucket_worker, args=(2, _free_port()), nprocs=2, join=True)


def test_two_rank_gloo_gradient_allreduce_equals_single_process():
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)
