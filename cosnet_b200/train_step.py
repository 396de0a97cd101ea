"""Training step around the drop-in model: the loop body of the reference's train.py:554-602, re-stated.

  loss       weighted BCE + 0.8 * L1 on both predictions               (train.py:176-216, :595-597)
  optimiser  SGD, two parameter groups (encoder | attention + depth + decoder)  (train.py:538-540, :220-346)
  schedule   poly LR; group 0 runs at 0.01 * lr, group 1 at 10 * lr     (train.py:161-174, :348-355)
  parallel   one process per GPU, frame pairs sharded across ranks, gradients averaged with ONE flat NCCL all-reduce
             per step (replaces nn.DataParallel's broadcast + reduce-to-GPU0, train.py:493)

The reference's per-iteration `gc.collect()` / `torch.cuda.empty_cache()` (train.py:619-620) and the host sync on the
positive-label count (`.item()`, train.py:184) are not reproduced: the weight is computed on the device.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.nn.functional as F


def weighted_bce(pred: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    """BCE whose per-element weight is (#pixels / #positive pixels) when the ground truth has positives, plain BCE
    otherwise (train.py:176-204).  #pixels = N*H*W (the channel dim is 1)."""
    positives = (label >= 0.5).sum()
    total = label.shape[0] * label.shape[2] * label.shape[3]
    weight = torch.where(positives > 0, total / positives.clamp(min=1).to(pred.dtype), torch.ones((), device=pred.device, dtype=pred.dtype))
    return F.binary_cross_entropy(pred, label) * weight     # a constant weight factors out of the mean


def segmentation_loss(pred1, gt1, pred2=None, gt2=None) -> torch.Tensor:
    """train.py:595-597: BCE + 0.8 * L1 for the current frame, plus the same for the counterpart frame."""
    loss = weighted_bce(pred1, gt1) + 0.8 * F.l1_loss(pred1, gt1)
    if pred2 is not None:
        loss = loss + weighted_bce(pred2, gt2) + 0.8 * F.l1_loss(pred2, gt2)
    return loss


def lr_poly(base_lr: float, it: int, max_iter: int, power: float, epoch: int) -> float:
    """train.py:348-355."""
    factor = 1.0 if epoch < 6 else 0.5
    return base_lr * factor * ((1.0 - float(it) / max_iter) ** power)


def parameter_groups(model) -> List[dict]:
    """Two SGD groups as train.py builds them for `resnet_aspp_add` (:241-243, :293-303): the RGB encoder, and
    rgb_attention + depth + decoder.  Frozen parameters (projection-shortcut BN affines) are left out."""
    def params(mods: Iterable[torch.nn.Module]):
        return [p for m in mods for p in m.parameters() if p.requires_grad]
    slow = params(model.get_params("encoder"))
    fast = params(model.get_params("rgb_attention") + model.get_params("depth") + model.get_params("decoder"))
    return [{"params": slow}, {"params": fast}]


def make_optimizer(model, learning_rate: float, momentum: float = 0.9, weight_decay: float = 0.0005):
    groups = parameter_groups(model)
    groups[0]["lr"] = learning_rate
    groups[1]["lr"] = 10 * learning_rate
    return torch.optim.SGD(groups, lr=learning_rate, momentum=momentum, weight_decay=weight_decay)


def adjust_learning_rate(optimizer, base_lr: float, it: int, epoch: int, max_iter: int, power: float = 0.9) -> float:
    """train.py:161-174: group 0 at 0.01 * lr, group 1 at 10 * lr."""
    lr = lr_poly(base_lr, it, max_iter, power, epoch)
    optimizer.param_groups[0]["lr"] = 0.01 * lr
    optimizer.param_groups[1]["lr"] = 10 * lr
    return lr


def allreduce_gradients(params: Iterable[torch.nn.Parameter], world_size: Optional[int] = None) -> None:
    """Average gradients over all ranks with one flat all-reduce (NCCL over NVLink on GPUs, gloo in CPU tests).
    Parameters without a gradient on this rank contribute zeros."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return
    world = world_size or dist.get_world_size()
    if world == 1:
        return
    plist = [p for p in params if p.requires_grad]
    if not plist:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in plist])
    dist.all_reduce(flat)
    flat /= world
    off = 0
    for p in plist:
        n = p.numel()
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n


class TrainStep:
    """One optimiser step on a rank-local shard of frame pairs (train.py:582-602)."""

    def __init__(self, model, learning_rate: float = 2.5e-4, momentum: float = 0.9, weight_decay: float = 0.0005,
                 power: float = 0.9, max_iter: int = 30000):
        self.model = model
        self.base_lr, self.power, self.max_iter = learning_rate, power, max_iter
        self.optimizer = make_optimizer(model, learning_rate, momentum, weight_decay)
        self.iteration = 0

    def __call__(self, rgb_a, rgb_b, depth_a, depth_b, gt_a, gt_b=None, epoch: int = 0) -> torch.Tensor:
        self.optimizer.zero_grad(set_to_none=True)
        adjust_learning_rate(self.optimizer, self.base_lr, self.iteration, epoch, self.max_iter, self.power)
        pred1, pred2, _ = self.model(rgb_a, rgb_b, depth_a, depth_b)
        loss = segmentation_loss(pred1, gt_a, pred2 if gt_b is not None else None, gt_b)
        loss.backward()
        allreduce_gradients(p for g in self.optimizer.param_groups for p in g["params"])
        self.optimizer.step()
        self.iteration += 1
        return loss.detach()
