"""Training step around the drop-in model: the loop body of the reference's train.py:554-602, re-stated.

  loss       weighted BCE + 0.8 * L1 on both predictions               (train.py:176-216, :595-597)
  optimiser  SGD, two parameter groups (encoder | attention + depth + decoder)  (train.py:538-540, :220-346)
  schedule   poly LR; group 0 runs at 0.01 * lr, group 1 at 10 * lr     (train.py:161-174, :348-355)
  parallel   one process per GPU, frame pairs sharded across ranks; gradients live in flat buckets (decoder + co-attention
             first, then the encoders in ~25 MB buckets, i.e. the order the backward produces them) and each bucket's
             NCCL all-reduce is launched from a post-accumulate hook the moment its last gradient is written, so the
             collectives overlap the rest of the backward (replaces nn.DataParallel's per-step broadcast of all
             parameters + reduce-to-GPU0, train.py:493)

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


class GradientBuckets:
    """Bucketed gradient all-reduce issued DURING the backward.

    Parameters are bucketed in reverse registration order (decoder -> co-attention -> depth encoder -> RGB encoder:
    approximately the order in which autograd finishes them), a small first bucket to get the first collective going early,
    then `bucket_bytes` each.  Gradients stay what autograd makes them (with zero_grad(set_to_none=True) it hands over its
    own tensors, no accumulation kernels); a post-accumulate-grad hook counts a bucket's parameters down, and when the last
    one is written the bucket is packed with ONE multi-tensor copy and `all_reduce(async_op=True)` is launched: NCCL orders
    the collective after the work already queued on the current stream and runs it on its own stream, concurrently with the
    remaining backward kernels.  `finish()` packs and launches whatever is left (parameters that got no gradient this step
    contribute zeros and keep grad = None), waits, and unpacks the averages into the gradient tensors, again one
    multi-tensor copy per bucket.  (Making `.grad` a view of the bucket instead -- the first version -- costs an in-place
    add kernel per parameter and step, 1059 of them for this model: +4.5 ms on a 68 ms step.)"""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 25 << 20, first_bucket_bytes: int = 1 << 20):
        import torch.distributed as dist
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.world = self.dist.get_world_size() if self.dist else 1
        plist = [p for p in params if p.requires_grad]
        self.buckets = []          # dicts: flat, views, params, pending, work
        self._bucket_of = {}
        cur, cur_bytes = [], 0
        limit = first_bucket_bytes
        for p in reversed(plist):
            if cur and (cur_bytes + p.numel() * p.element_size() > limit or p.dtype != cur[0].dtype or p.device != cur[0].device):
                self._close(cur)
                cur, cur_bytes, limit = [], 0, bucket_bytes
            cur.append(p)
            cur_bytes += p.numel() * p.element_size()
        if cur:
            self._close(cur)
        # one closure per parameter holding ITS bucket: the hook body is a decrement and a compare (it runs ~1000 times per
        # step from autograd's thread, so every dictionary lookup in it shows in the step time)
        self._hooks = [p.register_post_accumulate_grad_hook(self._make_hook(self.buckets[self._bucket_of[id(p)]])) for p in plist]
        self.launched_in_backward = 0
        # NCCL averages inside the collective; gloo (CPU tests) only sums
        self._avg = None
        if self.dist is not None and self.world > 1:
            try:
                if self.dist.get_backend() == "nccl":
                    self._avg = self.dist.ReduceOp.AVG
            except Exception:
                self._avg = None

    def _close(self, plist):
        flat = torch.zeros(sum(p.numel() for p in plist), dtype=plist[0].dtype, device=plist[0].device)
        views, off = [], 0
        for p in plist:
            views.append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
            self._bucket_of[id(p)] = len(self.buckets)
        self.buckets.append({"flat": flat, "views": views, "params": plist, "pending": len(plist), "work": None})

    def zero(self):
        """Start of a step: drop the gradients (autograd then hands over fresh tensors) and re-arm the buckets."""
        for b in self.buckets:
            b["pending"] = len(b["params"])
            b["work"] = None
            for p in b["params"]:
                p.grad = None
        self.launched_in_backward = 0

    def _pack_and_launch(self, b):
        if b["work"] is not None:
            return
        have = [(v, p.grad) for v, p in zip(b["views"], b["params"]) if p.grad is not None]
        if len(have) != len(b["params"]):
            b["flat"].zero_()
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        if self.dist is not None and self.world > 1:
            if self._avg is not None:
                b["work"] = self.dist.all_reduce(b["flat"], op=self._avg, async_op=True)
            else:
                b["work"] = self.dist.all_reduce(b["flat"], async_op=True)
        else:
            b["work"] = False

    def _make_hook(self, b):
        def hook(_p):
            b["pending"] -= 1
            if b["pending"] == 0:
                self._pack_and_launch(b)
                self.launched_in_backward += 1
        return hook

    def finish(self):
        """After backward: launch the buckets that never completed, wait for all collectives, write the averages back."""
        if self.dist is None or self.world == 1:
            return
        for b in self.buckets:
            self._pack_and_launch(b)
        for b in self.buckets:
            b["work"].wait()
            if self._avg is None:
                b["flat"].div_(self.world)
            have = [(p.grad, v) for v, p in zip(b["views"], b["params"]) if p.grad is not None]
            if have:
                torch._foreach_copy_([g for g, _ in have], [v for _, v in have])

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


class TrainStep:
    """One optimiser step on a rank-local shard of frame pairs (train.py:582-602)."""

    def __init__(self, model, learning_rate: float = 2.5e-4, momentum: float = 0.9, weight_decay: float = 0.0005,
                 power: float = 0.9, max_iter: int = 30000):
        self.model = model
        self.base_lr, self.power, self.max_iter = learning_rate, power, max_iter
        self.optimizer = make_optimizer(model, learning_rate, momentum, weight_decay)
        self.iteration = 0
        # bucketed, overlapped gradient all-reduce when a process group exists (one process per GPU); otherwise plain
        # single-process gradients
        import torch.distributed as dist
        self.buckets = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self.buckets = GradientBuckets(p for g in self.optimizer.param_groups for p in g["params"])

    def __call__(self, rgb_a, rgb_b, depth_a, depth_b, gt_a, gt_b=None, epoch: int = 0) -> torch.Tensor:
        if self.buckets is not None:
            self.buckets.zero()
        else:
            self.optimizer.zero_grad(set_to_none=True)
        adjust_learning_rate(self.optimizer, self.base_lr, self.iteration, epoch, self.max_iter, self.power)
        pred1, pred2, _ = self.model(rgb_a, rgb_b, depth_a, depth_b)
        loss = segmentation_loss(pred1, gt_a, pred2 if gt_b is not None else None, gt_b)
        loss.backward()
        if self.buckets is not None:
            self.buckets.finish()
        self.optimizer.step()
        self.iteration += 1
        return loss.detach()
