#!/usr/bin/env python
"""Diagnostic (GPU): the train-step fixture with the CUDA co-attention vs an eager fp32 torch co-attention, per parameter."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from cosnet_b200.backbone import Bottleneck
from cosnet_b200.rgbd_segmentation_raa import RGBDSegmentation_RAA
from cosnet_b200.train_step import TrainStep
from oracle.make_golden import HOT_PARAMS, train_step_inputs
from oracle.ref_harness import seeded_state
from tests.helpers import load_golden, rel_l2


def eager(v_a, v_b, weight, gate_weight, gate_bias, bf16_operands=False, gated_only=False):
    n, c, h, w = v_a.shape
    A, B = v_a.view(n, c, -1), v_b.view(n, c, -1)
    q = torch.nn.functional.linear(A.transpose(1, 2), weight)
    s = torch.bmm(q, B)
    z_b = torch.bmm(A, torch.softmax(s, 1)).view(n, c, h, w)
    z_a = torch.bmm(B, torch.softmax(s.transpose(1, 2), 1)).view(n, c, h, w)
    m_a = torch.sigmoid(torch.nn.functional.conv2d(z_a, gate_weight.view(1, c, 1, 1), gate_bias))
    with torch.no_grad():
        m_b = torch.sigmoid(torch.nn.functional.conv2d(z_b, gate_weight.view(1, c, 1, 1), gate_bias))
    return torch.cat([z_a * m_a, v_a], 1), torch.cat([z_b * m_b, v_b], 1)


def run(impl, tf32):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    fx = load_golden("train_step_n2_97x97")
    dev = torch.device("cuda:0")
    model = RGBDSegmentation_RAA(Bottleneck, [1, 1, 1, 1], [1, 1, 1, 1], num_classes=1).train()
    seeded_state(model, int(fx["seed"]))
    model = model.to(dev)
    if impl is not None:
        model.coattention_impl = impl
    before = {k: v.detach().clone() for k, v in model.named_parameters() if k in HOT_PARAMS}
    rgb, dep, gt = (torch.from_numpy(x).to(dev) for x in train_step_inputs(int(fx["seed"]) + 1, int(fx["n"]), int(fx["hw"])))
    step = TrainStep(model, learning_rate=float(fx["lr"]), max_iter=int(fx["max_iter"]))
    loss = float(step(rgb[0], rgb[1], dep[0], dep[1], gt[0], gt[1]))
    after = dict(model.named_parameters())
    out = {"loss_rel": abs(loss / float(fx["loss"]) - 1)}
    for k in HOT_PARAMS:
        out[k] = rel_l2((after[k].detach() - before[k]).cpu().numpy(), fx["delta__" + k])
    return out


for name, impl, tf32 in (("cuda op, tf32 convs", None, True), ("cuda op, fp32 convs", None, False),
                         ("eager fp32 op, tf32 convs", eager, True), ("eager fp32 op, fp32 convs", eager, False)):
    print(name, {k: f"{v:.2e}" for k, v in run(impl, tf32).items()}, flush=True)
