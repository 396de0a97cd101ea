"""Drive the UNMODIFIED reference model's co-attention with synthetic features -- TEST INFRASTRUCTURE.

Imports `RGBDSegmentation_RAA` from /root/reference (read-only; nothing is copied), replaces its two
encoders by stubs that return queued tensors, and captures with forward-pre-hooks exactly what the
hot path (rgbd_segmentation_RAA.py:150-187 / :204-238) hands to its consumers:

  gate / depth_gate                  -> raw Z_a then Z_b          (:177,179 / :228,230)
  reduce_channels_A / _B             -> cat_a / cat_b (RGB)       (:188-189)
  depth_reduce_channels (two calls)  -> cat_a then cat_b (depth)  (:239-241)

/root/reference only exists in the build container: this module is used by `oracle/make_golden.py`
(fixtures travel, the reference does not) and by the optional `cpu_baseline` timing when present.
"""
from __future__ import annotations

import os
import sys
import warnings

REFERENCE_ROOT = os.environ.get("COSNET_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "rgbd_segmentation_RAA.py"))


def import_reference():
    """Returns (RGBDSegmentation_RAA, Bottleneck) classes of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # make sure we do not pick up this repo's own drop-in modules of the same name
    for name in ("rgbd_segmentation_RAA", "deeplab", "deeplab.residual_net", "deeplab.deeplabv3_encoder",
                 "deeplab.config"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(REFERENCE_ROOT):
            del sys.modules[name]
    from deeplab.residual_net import Bottleneck  # type: ignore
    from rgbd_segmentation_RAA import RGBDSegmentation_RAA  # type: ignore
    assert RGBDSegmentation_RAA.__module__ and sys.modules["rgbd_segmentation_RAA"].__file__.startswith(REFERENCE_ROOT)
    return RGBDSegmentation_RAA, Bottleneck


def build_stubbed_reference(no_grad_for_counterpart: bool = True, small_backbone: bool = True):
    """Reference model whose encoders are stubs fed from `.queue`.

    small_backbone=True builds the throw-away ResNets with one block per layer (the encoders are
    replaced anyway); the co-attention / fusion / decoder layers are the reference's own.
    """
    import torch
    import torch.nn as nn

    RAA, Bottleneck = import_reference()
    blocks = [1, 1, 1, 1] if small_backbone else [3, 4, 23, 3]
    blocks_d = [1, 1, 1, 1] if small_backbone else [3, 4, 6, 3]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = RAA(Bottleneck, blocks, blocks_d, num_classes=1, no_grad_for_counterpart=no_grad_for_counterpart)

    class RgbStub(nn.Module):
        def __init__(self):
            super().__init__()
            self.queue = []

        def forward(self, x):
            f = self.queue.pop(0)
            if not torch.is_grad_enabled():
                f = f.detach()
            return f, x.new_zeros(1)

    class DepthStub(nn.Module):
        def __init__(self):
            super().__init__()
            self.queue = []

        def forward(self, x):
            f = self.queue.pop(0)
            if not torch.is_grad_enabled():
                f = f.detach()
            return f

    model.encoder = RgbStub()
    model.depth_encoder = DepthStub()
    return model


def run_reference(model, v_a, v_b, d_a, d_b, image_hw=None):
    """Runs the reference forward on synthetic features; returns captured tensors + outputs.

    v_a, v_b, d_a, d_b: torch fp32 [N, 256, H', W'] (may require grad).
    """
    import torch

    cap = {"gate": [], "depth_gate": [], "reduce_A": [], "reduce_B": [], "depth_reduce": []}
    hooks = []

    def grab(key):
        def hook(mod, inputs):
            cap[key].append(inputs[0])
            return None
        return hook

    hooks.append(model.gate.register_forward_pre_hook(grab("gate")))
    hooks.append(model.depth_gate.register_forward_pre_hook(grab("depth_gate")))
    hooks.append(model.reduce_channels_A.register_forward_pre_hook(grab("reduce_A")))
    hooks.append(model.reduce_channels_B.register_forward_pre_hook(grab("reduce_B")))
    hooks.append(model.depth_reduce_channels.register_forward_pre_hook(grab("depth_reduce")))
    n, _, h, w = v_a.shape
    if image_hw is None:
        image_hw = (h * 8, w * 8)
    img = torch.zeros(n, 3, *image_hw)
    dimg = torch.zeros(n, 1, *image_hw)
    model.encoder.queue = [v_a, v_b]
    model.depth_encoder.queue = [d_a, d_b]
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            x1, x2, _ = model(img, img, dimg, dimg)
    finally:
        for hk in hooks:
            hk.remove()
    return {
        "rgb_z_a": cap["gate"][0], "rgb_z_b": cap["gate"][1],
        "rgb_cat_a": cap["reduce_A"][0], "rgb_cat_b": cap["reduce_B"][0],
        "depth_z_a": cap["depth_gate"][0], "depth_z_b": cap["depth_gate"][1],
        "depth_cat_a": cap["depth_reduce"][0], "depth_cat_b": cap["depth_reduce"][1],
        "x1": x1, "x2": x2,
    }


def seeded_state(model, seed: int):
    """Fill EVERY tensor of model.state_dict() from a numpy stream (in key order) so that the reference and the drop-in
    model -- identical keys and shapes, tests/test_module_dropin.py -- can be given bit-identical weights on any machine
    without shipping a checkpoint.  Scales keep activations O(1): He-style fan-in scaling for conv / linear weights,
    BN weight 1 +- 0.1, running_var in [0.5, 1.5], everything else small."""
    import numpy as np
    import torch
    rng = np.random.default_rng(seed)
    with torch.no_grad():
        for key, t in model.state_dict().items():
            if not torch.is_floating_point(t):
                t.zero_()
                continue
            x = rng.standard_normal(tuple(t.shape)).astype(np.float32) if t.dim() > 0 else np.float32(rng.standard_normal())
            if key.endswith("running_var"):
                x = 1.0 + 0.5 * np.tanh(x)
            elif key.endswith("running_mean") or key.endswith(".bias"):
                x = 0.05 * x
            elif t.dim() == 1:                       # BN / PReLU weights
                x = (0.25 if t.numel() == 1 else 1.0) + 0.1 * x
            else:
                fan_in = int(np.prod(t.shape[1:]))
                x = x * np.float32(np.sqrt(2.0 / fan_in))
                if "similarity_weights" in key:      # keep the logits at the scale of a trained model (S std ~ 3, not
                    x = x * np.float32(0.25)         # near-one-hot softmaxes whose gradients are ill-conditioned)
            t.copy_(torch.from_numpy(np.asarray(x, dtype=np.float32)))
    return model
