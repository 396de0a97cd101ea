"""Drop-in boundary of the nn.Module: ctor, state_dict keys, get_params, load_state and the
forward wiring, checked on CPU against the unmodified reference with the oracle injected as the co-attention
operator (test-only injection; the product default refuses CPU tensors)."""
import os
import subprocess
import sys
import warnings

import numpy as np
import pytest
import torch

from cosnet_b200.backbone import Bottleneck
from cosnet_b200.rgbd_segmentation_raa import RGBDSegmentation_RAA
from oracle import coattn_oracle as orc
from oracle import ref_harness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not ref_harness.reference_available(), reason="reference tree not present")


def oracle_impl(v_a, v_b, weight, gate_weight, gate_bias, gated_only=False):
    out = orc.coattention(v_a.detach().numpy(), v_b.detach().numpy(), weight.detach().numpy(),
                          gate_weight.detach().numpy(), None if gate_bias is None else gate_bias.detach().numpy(),
                          dtype=np.float32)
    cat_a, cat_b = torch.from_numpy(out["cat_a"]), torch.from_numpy(out["cat_b"])
    if gated_only:
        c = v_a.shape[1]
        return cat_a[:, :c].contiguous(), cat_b[:, :c].contiguous()
    return cat_a, cat_b


def small_model(**kw):
    return RGBDSegmentation_RAA(Bottleneck, [1, 1, 1, 1], [1, 1, 1, 1], num_classes=1, **kw)


@needs_ref
def test_state_dict_keys_match_reference():
    RefRAA, RefBottleneck = ref_harness.import_reference()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = RefRAA(RefBottleneck, [1, 2, 1, 1], [1, 1, 2, 1], num_classes=1)
    mine = RGBDSegmentation_RAA(Bottleneck, [1, 2, 1, 1], [1, 1, 2, 1], num_classes=1)
    rs, ms = ref.state_dict(), mine.state_dict()
    assert list(rs.keys()) == list(ms.keys())
    assert all(rs[k].shape == ms[k].shape for k in rs)
    # same set of trainable parameters (the projection-shortcut BN affines are frozen, residual_net.py:132-133)
    assert [n for n, p in ref.named_parameters() if p.requires_grad] == [n for n, p in mine.named_parameters() if p.requires_grad]
    for subset in ("none", "all", "encoder", "rgb_attention", "rgb", "depth", "decoder"):
        r = [type(m).__name__ for m in ref.get_params(subset)]
        m = [type(x).__name__ for x in mine.get_params(subset)]
        assert r == m, subset
        assert sum(p.numel() for mod in ref.get_params(subset) for p in mod.parameters()) == \
            sum(p.numel() for mod in mine.get_params(subset) for p in mod.parameters())


def test_full_size_constructor_signature_and_param_count():
    m = RGBDSegmentation_RAA(Bottleneck, [3, 4, 23, 3], [3, 4, 6, 3], num_classes=1)   # train.py:379
    assert sum(p.numel() for p in m.parameters()) == 142371334                       # SURVEY.md 2.1 [measured]
    assert len(m.state_dict()) == 1059
    assert m.rgb_similarity_weights.weight.shape == (256, 256) and m.gate.weight.shape == (1, 256, 1, 1)
    assert m.depth_gate.bias.shape == (1,) and m.gate.bias is None


@needs_ref
@pytest.mark.parametrize("frozen", [True, False])
def test_forward_matches_reference_with_oracle_operator(frozen):
    RefRAA, RefBottleneck = ref_harness.import_reference()
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = RefRAA(RefBottleneck, [1, 1, 1, 1], [1, 1, 1, 1], num_classes=1, no_grad_for_counterpart=frozen).eval()
    mine = small_model(no_grad_for_counterpart=frozen).eval()
    mine.load_state_dict(ref.state_dict(), strict=True)
    mine.coattention_impl = oracle_impl
    g = torch.Generator().manual_seed(1)
    ra, rb = torch.randn(1, 3, 65, 57, generator=g), torch.randn(1, 3, 65, 57, generator=g)
    da, db = torch.randn(1, 1, 65, 57, generator=g), torch.randn(1, 1, 65, 57, generator=g)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = ref(ra, rb, da, db)
        got = mine(ra, rb, da, db)
    for w, g_ in zip(want, got):
        assert w.shape == g_.shape
        assert (w - g_).abs().max() < 1e-5
    # the third output is frame B's auxiliary map (rgbd_segmentation_RAA.py:143-148, :268)
    assert torch.allclose(got[2], mine.encoder(rb)[1])


def test_split_reduce_conv_equals_concat_path():
    """conv(cat([Zg, V]), W) == conv(Zg, W[:, :C]) + conv(V, W[:, C:]) -- the concat is never built."""
    torch.manual_seed(5)
    m = small_model().eval()
    m.coattention_impl = oracle_impl
    g = torch.Generator().manual_seed(6)
    ra, rb = torch.randn(1, 3, 49, 57, generator=g), torch.randn(1, 3, 49, 57, generator=g)
    da, db = torch.randn(1, 1, 49, 57, generator=g), torch.randn(1, 1, 49, 57, generator=g)
    with torch.no_grad():
        want = m(ra, rb, da, db)
        m.split_reduce_conv = True
        got = m(ra, rb, da, db)
    for w_, g_ in zip(want, got):
        assert (w_ - g_).abs().max() < 1e-5
    assert list(m.state_dict().keys()) == list(small_model().state_dict().keys())


def test_load_state_renames_legacy_keys():
    m = small_model()
    sd = m.state_dict()
    legacy = {}
    for k, v in sd.items():
        new = k
        if k.startswith("encoder.aspp."):
            new = k.replace("encoder.aspp.", "encoder.layer5.")
        elif k.startswith("encoder.backbone."):
            new = k.replace("encoder.backbone.", "encoder.")
        elif k.startswith("rgb_similarity_weights."):
            new = k.replace("rgb_similarity_weights.", "linear_e.")
        elif k.startswith("reduce_channels_A."):
            new = k.replace("reduce_channels_A.", "conv1.")
        elif k.startswith("reduce_channels_B."):
            new = k.replace("reduce_channels_B.", "conv2.")
        elif k.startswith("bn_A."):
            new = k.replace("bn_A.", "bn1.")
        elif k.startswith("bn_B."):
            new = k.replace("bn_B.", "bn2.")
        elif k.startswith("segmentation_classifier_A."):
            new = k.replace("segmentation_classifier_A.", "main_classifier1.")
        elif k.startswith("segmentation_classifier_B."):
            new = k.replace("segmentation_classifier_B.", "main_classifier2.")
        legacy["module." + new] = torch.full_like(v, 0.5) if v.is_floating_point() else v
    fresh = small_model()
    fresh.load_state(legacy)
    for k, v in fresh.state_dict().items():
        if v.is_floating_point():
            assert torch.all(v == 0.5), k


def test_default_operator_has_no_cpu_fallback():
    from cosnet_b200._lib import CoattnError
    m = small_model().eval()
    x = torch.zeros(1, 3, 33, 33)
    d = torch.zeros(1, 1, 33, 33)
    with pytest.raises(CoattnError), torch.no_grad():
        m(x, x, d, d)


def test_dropin_import_paths():
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "dropin"), ROOT]))
    code = ("from rgbd_segmentation_RAA import RGBDSegmentation_RAA; from deeplab.residual_net import Bottleneck; "
            "m = RGBDSegmentation_RAA(Bottleneck, [1,1,1,1], [1,1,1,1], num_classes=1); print(type(m).__module__)")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "cosnet_b200.rgbd_segmentation_raa" in out.stdout


def torch_impl(v_a, v_b, weight, gate_weight, gate_bias, gated_only=False):
    """Differentiable restatement of :154-187 (test-only), B-side mask under no_grad like :178-182."""
    import torch.nn.functional as F
    n, c, h, w = v_a.shape
    a, b = v_a.reshape(n, c, h * w), v_b.reshape(n, c, h * w)
    s = torch.bmm(F.linear(a.transpose(1, 2), weight), b)
    z_b = torch.bmm(a, F.softmax(s, dim=1)).view(n, c, h, w)
    z_a = torch.bmm(b, F.softmax(s, dim=2).transpose(1, 2)).view(n, c, h, w)
    m_a = torch.sigmoid(F.conv2d(z_a, gate_weight, gate_bias))
    with torch.no_grad():
        m_b = torch.sigmoid(F.conv2d(z_b, gate_weight, gate_bias))
    if gated_only:
        return z_a * m_a, z_b * m_b
    return torch.cat([z_a * m_a, v_a], 1), torch.cat([z_b * m_b, v_b], 1)


@needs_ref
@pytest.mark.parametrize("frozen", [True, False])
def test_train_mode_forward_backward_matches_reference(frozen):
    """Train-mode wiring around the operator: BatchNorm statistics (incl. the second, no-grad depth_bn update of
    :240-247), which parameters receive gradient (counterpart frozen or not, B-side depth branch dead) and the
    gradient values, against the unmodified reference."""
    RefRAA, RefBottleneck = ref_harness.import_reference()
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = RefRAA(RefBottleneck, [1, 1, 1, 1], [1, 1, 1, 1], num_classes=1, no_grad_for_counterpart=frozen).train()
    mine = small_model(no_grad_for_counterpart=frozen).train()
    mine.load_state_dict(ref.state_dict(), strict=True)
    mine.coattention_impl = torch_impl
    g = torch.Generator().manual_seed(2)
    ra, rb = torch.randn(2, 3, 41, 49, generator=g), torch.randn(2, 3, 41, 49, generator=g)
    da, db = torch.randn(2, 1, 41, 49, generator=g), torch.randn(2, 1, 41, 49, generator=g)
    r = [torch.randn(2, 1, 41, 49, generator=g) for _ in range(3)]
    outs = []
    for model in (ref, mine):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = model(ra, rb, da, db)
        sum((o * w).sum() for o, w in zip(out, r)).backward()
        outs.append(out)
    for w_, g_ in zip(*outs):
        assert (w_ - g_).abs().max() < 1e-5
    ref_buf, my_buf = dict(ref.named_buffers()), dict(mine.named_buffers())
    assert ref_buf.keys() == my_buf.keys()
    for k in ref_buf:
        assert torch.allclose(ref_buf[k].float(), my_buf[k].float(), rtol=1e-4, atol=1e-6), k
    ref_par, my_par = dict(ref.named_parameters()), dict(mine.named_parameters())
    # biases in front of a BatchNorm have a mathematically zero gradient (pure rounding noise): compare every
    # gradient against its own norm plus a floor tied to the largest gradient in the model
    top = max(float(p_.grad.norm()) for p_ in ref_par.values() if p_.grad is not None)
    for k, pr in ref_par.items():
        pm = my_par[k]
        assert (pr.grad is None) == (pm.grad is None), k
        if pr.grad is not None:
            assert float((pr.grad - pm.grad).norm()) <= 2e-3 * float(pr.grad.norm()) + 1e-5 * top, k


def test_bn_and_pointwise_fold_equal_the_unfused_operators():
    """SURVEY.md 8f row N3 (second half): conv -> BN_eval (-> 1x1 conv) evaluated as ONE split convolution with folded
    weights equals the reference's operator sequence (rgbd_segmentation_RAA.py:188-191, :239-247)."""
    torch.manual_seed(0)
    m = small_model().eval()
    for bn in (m.bn_A, m.bn_B, m.depth_bn):
        bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 1.5); bn.weight.data.normal_(1, 0.1); bn.bias.data.normal_()
    g, v = torch.randn(2, 256, 9, 7), torch.randn(2, 256, 9, 7)
    with torch.no_grad():
        want = m.bn_A(m.reduce_channels_A(torch.cat([g, v], 1)))
        got = m._split_conv_folded(*m._folded_reduce("A", m.reduce_channels_A, m.bn_A), m.reduce_channels_A, g, v)
        assert (want - got).abs().max() < 1e-5
        want = m.depth_weights(m.depth_bn(m.depth_reduce_channels(torch.cat([g, v], 1))))
        got = m._split_conv_folded(*m._folded_reduce("D", m.depth_reduce_channels, m.depth_bn, m.depth_weights),
                                   m.depth_reduce_channels, g, v)
        assert (want - got).abs().max() < 1e-5
        # the cache follows parameter updates
        w0, _ = m._folded_reduce("A", m.reduce_channels_A, m.bn_A)
        m.bn_A.weight.mul_(2.0)
        w1, _ = m._folded_reduce("A", m.reduce_channels_A, m.bn_A)
        assert torch.allclose(w1, 2.0 * w0)


def test_aspp_pre_tail_composes_to_forward():
    from cosnet_b200.backbone import ASPP
    torch.manual_seed(1)
    aspp = ASPP(64, 256, 32, [2, 3, 7], [2, 3, 7]).eval()
    x = torch.randn(1, 64, 9, 11)
    with torch.no_grad():
        assert torch.equal(aspp(x), aspp.prelu(aspp.bn(aspp.pre_tail(x))))
