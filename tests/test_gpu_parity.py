"""Parity of the CUDA co-attention (through the C ABI) against golden vectors of the reference and the
CPU oracle.  Needs a B200: run with `pytest -m gpu`.

Tolerances (floating point path, fp32 accumulation; rel-L2 on the module output, the [N,2C,H,W] concat):
  TOL    = 1e-3  default operand format (fp16): the bound BASELINE.json states, at every feature-scale tier;
  TOL_BF = 5e-3  COATTN_FLAG_BF16 operands: bf16 quantisation alone costs ~1e-3 at tiny L and ~2e-3 at
                 train-like logit scales (sigma = 1.0, S std ~ 5) -- reported, not hidden.
"""
import os

import numpy as np
import pytest
import torch

from oracle import coattn_oracle as orc
from tests.helpers import golden_inputs, load_golden, rel_l2, subsample

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

TOL = 1e-3
TOL_BF = 5e-3
C = 256


@pytest.fixture(scope="module")
def op():
    import __graft_entry__ as ge
    ge.build()
    from cosnet_b200 import coattention_forward_raw
    assert torch.cuda.is_available()
    return coattention_forward_raw


def run(op, v_a, v_b, w, g, b, bf16=False):
    dev = torch.device("cuda:0")
    t = lambda x: None if x is None else torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    cat_a, cat_b, z, lse = op(t(v_a), t(v_b), t(w), t(g), t(b), bf16)
    torch.cuda.synchronize()
    return cat_a.cpu().numpy(), cat_b.cpu().numpy(), z.cpu().numpy(), lse.cpu().numpy()


@pytest.mark.parametrize("bf16,tol", [(False, TOL), (True, TOL_BF)])
@pytest.mark.parametrize("name", ["fwd_n2_4x5_s066", "fwd_n1_3x43_s100"])
def test_golden_vectors_of_the_reference(op, name, bf16, tol):
    fx = load_golden(name)
    inp = golden_inputs(fx)
    for mod, (va, vb, w, g, b) in {
        "rgb": (inp["v_a"], inp["v_b"], inp["w_rgb"], inp["g_rgb"], None),
        "depth": (inp["d_a"], inp["d_b"], inp["w_dep"], inp["g_dep"], inp["b_dep"]),
    }.items():
        cat_a, cat_b, z, _ = run(op, va, vb, w, g, b, bf16)
        ref_a = np.concatenate([fx[f"{mod}_gated_a"], va], axis=1)
        ref_b = np.concatenate([fx[f"{mod}_gated_b"], vb], axis=1)
        assert rel_l2(cat_a, ref_a) < tol, (mod, rel_l2(cat_a, ref_a))
        assert rel_l2(cat_b, ref_b) < tol, (mod, rel_l2(cat_b, ref_b))
        # the passthrough half is a bit-exact copy (:186-187)
        assert np.array_equal(cat_a[:, C:], va) and np.array_equal(cat_b[:, C:], vb)
        if mod == "rgb":
            assert rel_l2(z[0].reshape(fx["rgb_z_a"].shape), fx["rgb_z_a"]) < 5 * tol
            assert rel_l2(z[1].reshape(fx["rgb_z_b"].shape), fx["rgb_z_b"]) < 5 * tol


LARGE = ["fwd_n1_60x60_s066", "fwd_n1_60x60_s100", "fwd_n1_61x81_s066", "fwd_n1_61x107_s066"]


@pytest.mark.parametrize("name", LARGE)
def test_reference_fixtures_at_baseline_sizes(op, name):
    """Outputs of the UNMODIFIED reference at the BASELINE.json sizes (L = 3600, 4941, 6527; oracle/make_golden.py keeps a
    strided subsample of the gated halves + the norm of the full tensors).  Bar: rel-L2 <= 1e-3 on the module output, i.e.
    the [N, 2C, H, W] concat, evaluated on the subsample (gated half against the fixture, passthrough half bit exact)."""
    fx = load_golden(name)
    inp = golden_inputs(fx)
    for mod, (va, vb, w, g, b) in {
        "rgb": (inp["v_a"], inp["v_b"], inp["w_rgb"], inp["g_rgb"], None),
        "depth": (inp["d_a"], inp["d_b"], inp["w_dep"], inp["g_dep"], inp["b_dep"]),
    }.items():
        cat_a, cat_b, z, _ = run(op, va, vb, w, g, b)
        for side, cat, v in (("a", cat_a, va), ("b", cat_b, vb)):
            assert np.array_equal(cat[:, C:], v)
            got, ref = subsample(cat[:, :C], fx), fx[f"{mod}_gated_{side}_sub"]
            err = np.linalg.norm(got.astype(np.float64) - ref)
            concat_norm = np.sqrt(np.linalg.norm(ref.astype(np.float64)) ** 2 + np.linalg.norm(subsample(v, fx).astype(np.float64)) ** 2)
            assert err / concat_norm < TOL, (mod, side, err / concat_norm)
            # the gated half on its own (stricter than the bar; fp16 operands measure 1e-4 ... 6e-4 here)
            assert rel_l2(got, ref) < 2e-3, (mod, side, rel_l2(got, ref))
            assert abs(np.linalg.norm(cat[:, :C].astype(np.float64)) / float(fx[f"{mod}_gated_{side}_norm"]) - 1) < 1e-3
        if mod == "rgb":
            assert rel_l2(subsample(z[0], fx), fx["rgb_z_a_sub"]) < 2e-3
            assert rel_l2(subsample(z[1], fx), fx["rgb_z_b_sub"]) < 2e-3


SHAPES = [
    # n, h, w           L      what it exercises
    (1, 1, 1),        # 1      single position: both softmaxes are 1
    (2, 8, 8),        # 64     exactly one key tile
    (1, 8, 16),       # 128    exactly one query tile
    (3, 12, 11),      # 132    ragged tail, odd batch
    (1, 31, 41),      # 1271   240x320 input (config.yaml:81), L odd -> scalar gate path
    (2, 60, 60),      # 3600   473x473 input (headline shape)
    (1, 61, 81),      # 4941   480x640 input
]


@pytest.mark.parametrize("n,h,w", SHAPES)
@pytest.mark.parametrize("bias", [False, True])
def test_against_oracle(op, n, h, w, bias):
    v_a, v_b = orc.synthetic_features(1000 + h * w, n, h, w, 0.66)
    W, g, b = orc.synthetic_weights(2000 + h * w, bias=bias)
    ref = orc.coattention(v_a, v_b, W, g, b)
    cat_a, cat_b, z, lse = run(op, v_a, v_b, W, g, b)
    assert rel_l2(cat_a, ref["cat_a"]) < TOL
    assert rel_l2(cat_b, ref["cat_b"]) < TOL
    L = h * w
    assert np.abs(lse[0] - ref["lse_a"]).max() < 5e-2
    assert np.abs(lse[1] - ref["lse_b"]).max() < 5e-2
    assert np.array_equal(cat_a[:, C:], v_a) and np.array_equal(cat_b[:, C:], v_b)
    assert np.isfinite(cat_a).all() and np.isfinite(cat_b).all() and np.isfinite(z).all()


def test_480x854_shape(op):
    # 480x854 input -> 61x107 features (L = 6527, what the reference really produces)
    n, h, w = 1, 61, 107
    v_a, v_b = orc.synthetic_features(77, n, h, w, 0.66)
    W, g, b = orc.synthetic_weights(78, bias=True)
    ref = orc.coattention(v_a, v_b, W, g, b, dtype=np.float32)
    cat_a, cat_b, _, _ = run(op, v_a, v_b, W, g, b)
    assert rel_l2(cat_a, ref["cat_a"]) < TOL
    assert rel_l2(cat_b, ref["cat_b"]) < TOL


@pytest.mark.parametrize("sigma,bf16,tol", [
    (0.01, False, TOL), (0.25, False, TOL), (0.66, False, TOL), (1.0, False, TOL), (1.5, False, TOL),
    (0.01, True, TOL_BF), (0.66, True, TOL_BF), (1.0, True, TOL_BF), (1.5, True, 2e-2)])
def test_feature_scale_tiers(op, sigma, bf16, tol):
    # T0 (degenerate uniform softmax) .. T2+ (peaky softmax, exercises the lazy O rescale), SURVEY.md 7.3-2
    n, h, w = 1, 24, 24
    v_a, v_b = orc.synthetic_features(31, n, h, w, sigma)
    W, g, b = orc.synthetic_weights(32, bias=True)
    ref = orc.coattention(v_a, v_b, W, g, b)
    cat_a, cat_b, _, _ = run(op, v_a, v_b, W, g, b, bf16)
    assert rel_l2(cat_a, ref["cat_a"]) < tol, rel_l2(cat_a, ref["cat_a"])
    assert rel_l2(cat_b, ref["cat_b"]) < tol, rel_l2(cat_b, ref["cat_b"])


def test_large_logits_stay_finite(op):
    # |S| in the hundreds: exp() of raw logits would overflow without the running-max bookkeeping
    n, h, w = 1, 16, 16
    v_a, v_b = orc.synthetic_features(41, n, h, w, 4.0)
    W, g, b = orc.synthetic_weights(42, bias=False)
    ref = orc.coattention(v_a, v_b, W, g, b)
    cat_a, cat_b, z, lse = run(op, v_a, v_b, W, g, b)
    assert np.isfinite(cat_a).all() and np.isfinite(cat_b).all() and np.isfinite(lse).all()
    # near one-hot softmax: attended features must stay inside the convex hull of the values
    zb = z[1].reshape(n, C, h * w)
    A = v_a.reshape(n, C, -1)
    assert (zb <= A.max(axis=2, keepdims=True) + 1e-2).all() and (zb >= A.min(axis=2, keepdims=True) - 1e-2).all()
    assert np.abs(lse[0] - ref["lse_a"]).max() < 0.5


def test_bf16_headline_shape(op):
    n, h, w = 1, 60, 60
    v_a, v_b = orc.synthetic_features(61, n, h, w, 0.66)
    W, g, b = orc.synthetic_weights(62, bias=False)
    ref = orc.coattention(v_a, v_b, W, g, b, dtype=np.float32)
    cat_a, cat_b, _, _ = run(op, v_a, v_b, W, g, b, bf16=True)
    assert rel_l2(cat_a, ref["cat_a"]) < TOL      # bf16 does meet 1e-3 at the headline shape (L = 3600, S std ~ 2)
    assert rel_l2(cat_b, ref["cat_b"]) < TOL


def test_fused_and_unfused_gate_agree(op):
    n, h, w = 2, 12, 11
    v_a, v_b = orc.synthetic_features(71, n, h, w, 0.66)
    W, g, b = orc.synthetic_weights(72, bias=True)
    dev = torch.device("cuda:0")
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    fused = op(t(v_a), t(v_b), t(W), t(g), t(b), False, False, True)          # + mask
    unfused = op(t(v_a), t(v_b), t(W), t(g), t(b), False, True)
    torch.cuda.synchronize()
    assert (fused[0] - unfused[0]).abs().max() < 2e-6 and (fused[1] - unfused[1]).abs().max() < 2e-6
    assert torch.equal(fused[2], unfused[2]) and torch.equal(fused[3], unfused[3])      # raw z and lse are identical
    ref = orc.coattention(v_a, v_b, W, g, b)
    mask = fused[4].cpu().numpy()
    assert np.abs(mask[0] - ref["mask_a"].reshape(n, -1)).max() < 1e-4
    assert np.abs(mask[1] - ref["mask_b"].reshape(n, -1)).max() < 1e-4
    # without the optional z output the concat is unchanged
    no_z = op(t(v_a), t(v_b), t(W), t(g), t(b), False, False, False, False)
    torch.cuda.synchronize()
    assert no_z[2] is None and torch.equal(no_z[0], fused[0]) and torch.equal(no_z[1], fused[1])


def test_deterministic_and_batch_invariant(op):
    n, h, w = 3, 20, 20
    v_a, v_b = orc.synthetic_features(51, n, h, w, 0.66)
    W, g, b = orc.synthetic_weights(52, bias=True)
    first = run(op, v_a, v_b, W, g, b)
    again = run(op, v_a, v_b, W, g, b)
    for x, y in zip(first, again):
        assert np.array_equal(x, y)
    # a sample's result does not depend on what else is in the batch
    solo = run(op, v_a[1:2], v_b[1:2], W, g, b)
    assert np.array_equal(solo[0][0], first[0][1]) and np.array_equal(solo[1][0], first[1][1])


def test_full_size_properties_batch32(op):
    """BASELINE cfg 2 size (batch 32, 60x60): size-independent properties instead of an O(L^2) oracle."""
    dev = torch.device("cuda:0")
    n, h, w = 32, 60, 60
    L = h * w
    gen = torch.Generator(device=dev); gen.manual_seed(1234)
    x = torch.randn((2, n, C, h, w), generator=gen, device=dev)
    feats = torch.where(x >= 0, x, 0.25 * x) * 0.66
    v_a, v_b = feats[0].contiguous(), feats[1].contiguous()
    v_b[:, 7] = 1.0            # a constant value channel must be reproduced exactly by any softmax average
    v_a[:, 9] = -2.0
    W, g, b = (torch.from_numpy(t).to(dev) for t in orc.synthetic_weights(5, bias=True))
    cat_a, cat_b, z, lse = op(v_a, v_b, W, g, b)
    torch.cuda.synchronize()
    assert torch.isfinite(cat_a).all() and torch.isfinite(cat_b).all()
    assert (z[0][:, 7] - 1.0).abs().max() < 2e-3      # sum_j P_a[i, j] == 1
    assert (z[1][:, 9] + 2.0).abs().max() < 4e-3      # sum_i P_b[i, j] == 1
    # convex-hull property on every channel
    zmax, zmin = z[0].amax(dim=2), z[0].amin(dim=2)
    vb = v_b.view(n, C, L)
    assert (zmax <= vb.amax(dim=2) + 1e-2).all() and (zmin >= vb.amin(dim=2) - 1e-2).all()
    # gate: cat[:, :C] = z * sigmoid(g.z + b), checked with torch on the kernel's own z
    za = z[0].view(n, C, h, w)
    mask = torch.sigmoid((za * g.view(1, C, 1, 1)).sum(1, keepdim=True) + b)
    assert (cat_a[:, :C] - za * mask).abs().max() < 1e-5
    assert torch.equal(cat_a[:, C:], v_a) and torch.equal(cat_b[:, C:], v_b)
    # batch invariance at full size: sample 17 alone reproduces its slice bit for bit
    solo = op(v_a[17:18].contiguous(), v_b[17:18].contiguous(), W, g, b)
    torch.cuda.synchronize()
    assert torch.equal(solo[0][0], cat_a[17]) and torch.equal(solo[1][0], cat_b[17])
    # spot-check 3 samples against the oracle restricted to a handful of query rows (O(L) per row)
    for s in (0, 13, 31):
        A = v_a[s].view(C, L).double().cpu().numpy(); B = v_b[s].view(C, L).double().cpu().numpy()
        Wd = W.double().cpu().numpy()
        rows = np.array([0, 1, 127, 128, 2047, 3599])
        q = (Wd @ A[:, rows]).T                       # [rows, C]
        srow = q @ B                                   # [rows, L]
        p = orc.softmax(srow, axis=1)
        ref = (B @ p.T)                                # [C, rows]
        got = z[0][s][:, rows].double().cpu().numpy()
        assert rel_l2(got, ref) < 1e-3, rel_l2(got, ref)


def test_cfg3_batch16_61x107(op):
    """BASELINE cfg 3 as worded: 60x107-class features (480x854 input -> 61x107, L = 6527), batch 16.  Size-independent
    properties on the whole batch + two samples against the fp32 oracle + bit-exact batch invariance."""
    dev = torch.device("cuda:0")
    n, h, w = 16, 61, 107
    L = h * w
    v_a, v_b = (torch.from_numpy(x).to(dev) for x in orc.synthetic_features(303, n, h, w, 0.66))
    W, g, b = (torch.from_numpy(t).to(dev) for t in orc.synthetic_weights(304, bias=True))
    v_b[:, 5] = 1.0
    cat_a, cat_b, z, lse = op(v_a, v_b, W, g, b)
    torch.cuda.synchronize()
    assert torch.isfinite(cat_a).all() and torch.isfinite(cat_b).all()
    assert (z[0][:, 5] - 1.0).abs().max() < 2e-3
    assert torch.equal(cat_a[:, C:], v_a) and torch.equal(cat_b[:, C:], v_b)
    for s in (0, 15):
        ref = orc.coattention(v_a[s:s + 1].cpu().numpy(), v_b[s:s + 1].cpu().numpy(), W.cpu().numpy(), g.cpu().numpy(),
                              b.cpu().numpy(), dtype=np.float32)
        assert rel_l2(cat_a[s:s + 1].cpu().numpy(), ref["cat_a"]) < TOL
        assert rel_l2(cat_b[s:s + 1].cpu().numpy(), ref["cat_b"]) < TOL
        solo = op(v_a[s:s + 1].contiguous(), v_b[s:s + 1].contiguous(), W, g, b)
        torch.cuda.synchronize()
        assert torch.equal(solo[0][0], cat_a[s]) and torch.equal(solo[1][0], cat_b[s])


def test_fp16_range_guard(op):
    """fp16 operands clamp at +-65504.  A feature or a projected value beyond that is REPORTED (status block of the
    workspace -> CoattnError from `check_overflow` / the next call), never silently clipped; bf16 operands (fp32 exponent
    range) take the same inputs without a flag and match the oracle."""
    import cosnet_b200
    from cosnet_b200 import _lib
    dev = torch.device("cuda:0")
    n, h, w = 1, 12, 11
    v_a, v_b = (torch.from_numpy(x).to(dev) for x in orc.synthetic_features(401, n, h, w, 0.66))
    W, g, b = (torch.from_numpy(t).to(dev) for t in orc.synthetic_weights(402, bias=True))
    cosnet_b200.check_overflow()                       # clean slate
    op(v_a, v_b, W, g, b)
    assert cosnet_b200.check_overflow()["absmax"] > 1.0   # in range: no error, the cast kernels report max |v|
    # (1) a feature beyond the fp16 range
    big = v_b.clone(); big[0, 3, 2, 2] = 7.0e4
    op(v_a, big, W, g, b)
    with pytest.raises(_lib.CoattnError, match="V_b"):
        cosnet_b200.check_overflow()
    cosnet_b200.check_overflow()                       # the flag was cleared by the report
    # (2) features in range, Q = W V_a beyond it
    op(v_a * 1.0e3, v_b, W * 400.0, g, b)
    with pytest.raises(_lib.CoattnError, match="Q = W V_a"):
        cosnet_b200.check_overflow()
    # (3) lazy mode: the NEXT call on the same stream raises
    nan = v_a.clone(); nan[0, 0, 0, 0] = float("nan")
    op(nan, v_b, W, g, b)
    torch.cuda.synchronize()
    with pytest.raises(_lib.CoattnError, match="V_a"):
        op(v_a, v_b, W, g, b)
    # (4) bf16 operands: same large input, no flag, finite results close to the oracle
    scaled_a, scaled_b = v_a * 300.0, v_b * 300.0       # |S| in the tens of thousands: near one-hot softmaxes
    scaled_b[0, 3, 2, 2] = 7.0e4
    out = op(scaled_a, scaled_b, W, g, b, True)
    cosnet_b200.check_overflow()
    assert torch.isfinite(out[0]).all() and torch.isfinite(out[1]).all()


@pytest.mark.parametrize("scale", [2.0 ** -10, 2.0 ** -16])
def test_tiny_features(op, scale):
    """Features far below 1 (an eval-mode random-init encoder gives std 0.0066; 2^-16 is the fp16 subnormal
    range): fp16 operands keep 1e-3 on the module output down to the subnormal range, where bf16 is the remedy."""
    n, h, w = 1, 16, 16
    v_a, v_b = orc.synthetic_features(411, n, h, w, 0.66)
    v_a, v_b = (v_a * np.float32(scale)), (v_b * np.float32(scale))
    W, g, b = orc.synthetic_weights(412, bias=True)
    ref = orc.coattention(v_a, v_b, W, g, b)
    subnormal = scale < 2.0 ** -12
    cat_a, cat_b, _, _ = run(op, v_a, v_b, W, g, b, bf16=subnormal)
    tol = TOL_BF if subnormal else TOL
    assert rel_l2(cat_a, ref["cat_a"]) < tol, rel_l2(cat_a, ref["cat_a"])
    assert rel_l2(cat_b, ref["cat_b"]) < tol, rel_l2(cat_b, ref["cat_b"])


def test_argument_checks(op):
    dev = torch.device("cuda:0")
    x = torch.zeros(1, 256, 4, 4, device=dev)
    with pytest.raises(ValueError):
        op(x, torch.zeros(1, 256, 4, 5, device=dev), torch.zeros(256, 256, device=dev), torch.zeros(256, device=dev))
    with pytest.raises(TypeError):
        op(x.half(), x.half(), torch.zeros(256, 256, device=dev), torch.zeros(256, device=dev))
    with pytest.raises(ValueError):
        y = torch.zeros(1, 128, 4, 4, device=dev)
        op(y, y, torch.zeros(256, 256, device=dev), torch.zeros(128, device=dev))


class _Stub(torch.nn.Module):
    """Encoder stand-in that replays queued synthetic features (same trick as oracle/ref_harness.py)."""

    def __init__(self, as_tuple):
        super().__init__()
        self.queue, self.as_tuple = [], as_tuple

    def forward(self, x):
        f = self.queue.pop(0)
        return (f, x.new_zeros(1)) if self.as_tuple else f


@pytest.mark.parametrize("sigma,hw", [(0.66, (60, 60)), (1.0, (30, 30))])
def test_end_to_end_masks_agree_with_oracle_operator(op, sigma, hw):
    """BASELINE.json: binarised masks must agree on >= 99.9 % of pixels.  The whole drop-in module (reduce convs,
    BN, depth fusion, classifiers, x8 upsampling) runs twice on the same weights: once with the CUDA co-attention,
    once with the CPU oracle injected as the operator."""
    from cosnet_b200.backbone import Bottleneck
    from cosnet_b200.rgbd_segmentation_raa import RGBDSegmentation_RAA
    dev = torch.device("cuda:0")
    torch.manual_seed(1234)
    model = RGBDSegmentation_RAA(Bottleneck, [1, 1, 1, 1], [1, 1, 1, 1], num_classes=1).eval()
    model.encoder, model.depth_encoder = _Stub(True), _Stub(False)
    model = model.to(dev)
    n, (h, w) = 2, hw
    feats = [torch.from_numpy(f).to(dev) for f in orc.synthetic_features(91, n, h, w, sigma, count=4)]
    img = torch.zeros(n, 3, h * 8, w * 8, device=dev)
    dimg = torch.zeros(n, 1, h * 8, w * 8, device=dev)

    def run_model():
        model.encoder.queue = [feats[0], feats[1]]
        model.depth_encoder.queue = [feats[2], feats[3]]
        with torch.no_grad():
            return model(img, img, dimg, dimg)

    x1, x2, _ = run_model()

    def oracle_impl(v_a, v_b, weight, gate_weight, gate_bias):
        out = orc.coattention(v_a.cpu().numpy(), v_b.cpu().numpy(), weight.detach().cpu().numpy(),
                              gate_weight.detach().cpu().numpy(),
                              None if gate_bias is None else gate_bias.detach().cpu().numpy(), dtype=np.float32)
        return torch.from_numpy(out["cat_a"]).to(dev), torch.from_numpy(out["cat_b"]).to(dev)

    model.coattention_impl = oracle_impl
    r1, r2, _ = run_model()
    for got, ref in ((x1, r1), (x2, r2)):
        assert got.shape == (n, 1, h * 8, w * 8)
        # the test is only meaningful if the maps straddle the threshold
        assert float(ref.min()) < 0.5 < float(ref.max())
        agree = ((got > 0.5) == (ref > 0.5)).float().mean().item()
        assert agree >= 0.999, agree
        assert (got - ref).abs().max().item() < 2e-3


def test_fused_and_unfused_prep_agree(op):
    for (n, h, w) in ((2, 12, 11), (1, 31, 41), (1, 60, 60)):
        v_a, v_b = orc.synthetic_features(85, n, h, w, 0.66)
        W, g, b = orc.synthetic_weights(86, bias=True)
        dev = torch.device("cuda:0")
        t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
        mn = op(t(v_a), t(v_b), t(W), t(g), t(b), unfolded=True)        # channel-major (MN-major) operands, stand-alone projection
        kmajor = op(t(v_a), t(v_b), t(W), t(g), t(b), kmajor=True)      # transposing prep + fused convert/projection
        unfused = op(t(v_a), t(v_b), t(W), t(g), t(b), unfused_prep=True)
        folded = op(t(v_a), t(v_b), t(W), t(g), t(b))                   # default: projection inside the attend kernel
        torch.cuda.synchronize()
        for x, y in zip(kmajor, unfused):
            assert torch.equal(x, y)      # same conversions, same MMAs: bit-identical
        for x, y in zip(mn, unfused):     # same operand values, different operand layout
            assert (x - y).abs().max() < 1e-5
        # the in-kernel projection: frame-A side from the same Q = W V_a (same MMAs, fp32 accumulation order may differ in the
        # last bit); frame-B side from W^T V_b rounded to 16 bits instead of W V_a -- equal to that rounding
        assert (folded[0] - mn[0]).abs().max() < 1e-5 * max(1.0, float(mn[0].abs().max()))
        assert rel_l2(folded[1][:, :C].cpu().numpy(), mn[1][:, :C].cpu().numpy()) < 5e-4
        assert torch.equal(folded[1][:, C:], mn[1][:, C:])


@pytest.mark.parametrize("n,h,w", [(2, 12, 11), (1, 40, 40)])
def test_single_cta_cross_check_kernel(op, n, h, w):
    """COATTN_FLAG_SINGLE_CTA: the earlier single-CTA attend kernel + stand-alone passthrough kernel (different tiling,
    same math) must agree with the oracle and with the CTA-pair kernel."""
    v_a, v_b = orc.synthetic_features(87, n, h, w, 0.66)
    W, g, b = orc.synthetic_weights(88, bias=True)
    dev = torch.device("cuda:0")
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    pair = op(t(v_a), t(v_b), t(W), t(g), t(b))
    single = op(t(v_a), t(v_b), t(W), t(g), t(b), single_cta=True)
    torch.cuda.synchronize()
    ref = orc.coattention(v_a, v_b, W, g, b)
    assert rel_l2(single[0].cpu().numpy(), ref["cat_a"]) < TOL and rel_l2(single[1].cpu().numpy(), ref["cat_b"]) < TOL
    assert (single[0] - pair[0]).abs().max() < 5e-3 and (single[1] - pair[1]).abs().max() < 5e-3
    assert torch.equal(single[0][:, C:], t(v_a)) and torch.equal(single[1][:, C:], t(v_b))


def test_gated_only_output(op):
    n, h, w = 2, 12, 11
    v_a, v_b = orc.synthetic_features(91, n, h, w, 0.66)
    W, g, b = orc.synthetic_weights(92, bias=True)
    dev = torch.device("cuda:0")
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    full = op(t(v_a), t(v_b), t(W), t(g), t(b))
    gated = op(t(v_a), t(v_b), t(W), t(g), t(b), gated_only=True)
    torch.cuda.synchronize()
    assert gated[0].shape == (n, C, h, w) and gated[1].shape == (n, C, h, w)
    assert torch.equal(gated[0], full[0][:, :C]) and torch.equal(gated[1], full[1][:, :C])


def test_frame_a_only_matches_full(op):
    n, h, w = 3, 12, 11
    v_a, v_b = orc.synthetic_features(81, n, h, w, 0.66)
    W, g, b = orc.synthetic_weights(82, bias=True)
    dev = torch.device("cuda:0")
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    full = op(t(v_a), t(v_b), t(W), t(g), t(b))
    a_only = op(t(v_a), t(v_b), t(W), t(g), t(b), False, False, False, True, False, True)
    torch.cuda.synchronize()
    assert a_only[1] is None
    assert torch.equal(a_only[0], full[0])
    assert torch.equal(a_only[2][0], full[2][0]) and torch.equal(a_only[3][0], full[3][0])


def test_multi_reference_inference_matches_pairwise_loop():
    """test.py:278-305: mean over references of the frame-A mask; the batched, query-hoisted path must equal running the
    drop-in module once per (query, reference) pair."""
    from cosnet_b200.backbone import Bottleneck
    from cosnet_b200.inference import segment_with_references
    from cosnet_b200.rgbd_segmentation_raa import RGBDSegmentation_RAA
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    model = RGBDSegmentation_RAA(Bottleneck, [1, 1, 1, 1], [1, 1, 1, 1], num_classes=1).to(dev).eval()
    q, r, hw = 2, 3, 97
    tgt, tgt_d = torch.randn(q, 3, hw, hw, device=dev), torch.randn(q, 1, hw, hw, device=dev)
    refs, refs_d = torch.randn(q, r, 3, hw, hw, device=dev), torch.randn(q, r, 1, hw, hw, device=dev)
    got = segment_with_references(model, tgt, tgt_d, refs, refs_d)
    want = torch.zeros_like(got)
    with torch.no_grad():
        for i in range(r):                                       # the reference's loop (test.py:287-301)
            want += model(tgt, refs[:, i], tgt_d, refs_d[:, i])[0]
    want /= r
    assert got.shape == (q, 1, hw, hw)
    assert (got - want).abs().max().item() < 1e-5


@pytest.mark.parametrize("host_passthrough", [False, True])
def test_host_pipeline_matches_resident_path(op, host_passthrough):
    """The host-buffer entry point (pinned host in/out, chunks rotating over three streams) reproduces the resident
    path bit for bit, with the device or the host writing the passthrough half."""
    from cosnet_b200.coattention import HostPipeline
    dev = torch.device("cuda:0")
    n, c, h, w = 7, 256, 12, 11          # 7 pairs in chunks of 3: a ragged last chunk
    v_a, v_b = (torch.from_numpy(x) for x in orc.synthetic_features(95, n, h, w, 0.66))
    W, g, b = (torch.from_numpy(x).to(dev) for x in orc.synthetic_weights(96, bias=True))
    want = op(v_a.to(dev), v_b.to(dev), W, g, b)
    pipe = HostPipeline(n, c, h, w, chunk=3, slots=3, device=dev, host_passthrough=host_passthrough)
    out_a = torch.empty(n, 2 * c, h, w).pin_memory()
    out_b = torch.empty(n, 2 * c, h, w).pin_memory()
    for _ in range(2):                   # second call reuses the slots
        out_a.zero_(); out_b.zero_()
        pipe(v_a.pin_memory(), v_b.pin_memory(), W, g, b, out_a, out_b)
        torch.cuda.synchronize()
        pipe.wait_host()
        assert torch.equal(out_a, want[0].cpu()) and torch.equal(out_b, want[1].cpu())
    # streamed: consecutive calls without the join in between (the bench's e2e leg), one join at the end
    if not host_passthrough:
        out_a.zero_(); out_b.zero_()
        out_a2 = torch.empty_like(out_a).pin_memory(); out_b2 = torch.empty_like(out_b).pin_memory()
        pipe(v_a.pin_memory(), v_b.pin_memory(), W, g, b, out_a, out_b, join=False)
        pipe(v_b.pin_memory(), v_a.pin_memory(), W, g, b, out_a2, out_b2, join=False)
        pipe.join()
        torch.cuda.synchronize()
        swapped = op(v_b.to(dev), v_a.to(dev), W, g, b)
        assert torch.equal(out_a, want[0].cpu()) and torch.equal(out_b, want[1].cpu())
        assert torch.equal(out_a2, swapped[0].cpu()) and torch.equal(out_b2, swapped[1].cpu())


def test_forward_is_cuda_graph_capturable(op):
    """The whole modality call (4 launches, no host synchronisation, no allocation inside the library) can be captured
    into a CUDA graph and replayed on new inputs -- the way to run launch-bound small shapes (test.py-style inference
    on a few pairs)."""
    from cosnet_b200 import _lib
    from cosnet_b200.coattention import workspace_bytes
    dev = torch.device("cuda:0")
    lib = _lib.load()
    n, c, h, w = 2, 256, 12, 11
    v_a = torch.empty(n, c, h, w, device=dev)
    v_b = torch.empty_like(v_a)
    W, g, b = (torch.from_numpy(x).to(dev) for x in orc.synthetic_weights(71, bias=True))
    cat_a = torch.empty(n, 2 * c, h, w, device=dev)
    cat_b = torch.empty_like(cat_a)
    nbytes = workspace_bytes(n, c, h, w)
    ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
    wsp = (ws.data_ptr() + 1023) // 1024 * 1024

    def call(stream):
        _lib.check(lib.coattn_forward(v_a.data_ptr(), v_b.data_ptr(), W.data_ptr(), g.data_ptr(), b.data_ptr(),
                                      cat_a.data_ptr(), cat_b.data_ptr(), None, None, None, wsp, nbytes, n, c, h, w, 0,
                                      stream.cuda_stream), "coattn_forward")

    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        call(side)                                  # warm-up outside the capture (function attributes, descriptors)
    torch.cuda.current_stream(dev).wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        call(torch.cuda.current_stream(dev))
    for seed in (72, 73):
        fa, fb = (torch.from_numpy(x).to(dev) for x in orc.synthetic_features(seed, n, h, w, 0.66))
        v_a.copy_(fa); v_b.copy_(fb)
        graph.replay()
        torch.cuda.synchronize()
        want = op(fa, fb, W, g, b)
        assert torch.equal(cat_a, want[0]) and torch.equal(cat_b, want[1])


@pytest.mark.parametrize("n,h,w", [(1, 1, 1), (2, 12, 11), (1, 31, 41), (1, 60, 60), (2, 61, 81)])
def test_softmax16_cross_check_kernel(op, n, h, w):
    """COATTN_FLAG_SOFTMAX16: the attend kernel with 16 softmax warps (four column groups, 32 key columns per thread)
    against the default 8-warp layout (two groups): same MMAs, same exponentials, only the order of a few fp32
    additions differs; both against the oracle.  Shapes cover single-tile, ragged and multi-item cases."""
    v_a, v_b = orc.synthetic_features(89, n, h, w, 0.66)
    W, g, b = orc.synthetic_weights(90, bias=True)
    dev = torch.device("cuda:0")
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    narrow = op(t(v_a), t(v_b), t(W), t(g), t(b), unfolded=True)     # the 16-warp variant keeps the stand-alone projection
    wide = op(t(v_a), t(v_b), t(W), t(g), t(b), softmax16=True)
    torch.cuda.synchronize()
    ref = orc.coattention(v_a, v_b, W, g, b)
    for out in (wide, narrow):
        assert rel_l2(out[0].cpu().numpy(), ref["cat_a"]) < TOL and rel_l2(out[1].cpu().numpy(), ref["cat_b"]) < TOL
    for x, y in zip(wide, narrow):    # cat_a, cat_b, z, lse
        assert (x - y).abs().max() <= 1e-5 * max(1.0, float(y.abs().max()))


def test_concurrent_host_threads():
    """SURVEY.md 8b threading contract (nn.DataParallel's parallel_apply, train.py:493): concurrent calls from several
    host threads -- one per visible device, or two streams of one device -- give the single-threaded results bit for bit."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "gpu_threads.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr


def test_unaligned_tensors_take_the_scalar_paths(op):
    """Feature tensors that are only 4-byte aligned (views into a larger buffer at an odd element offset): the cast,
    passthrough and gate stages fall back to scalar accesses; the result is the aligned one bit for bit."""
    n, c, h, w = 2, 256, 12, 12
    dev = torch.device("cuda:0")
    v_a, v_b = (torch.from_numpy(x).to(dev) for x in orc.synthetic_features(97, n, h, w, 0.66))
    W, g, b = (torch.from_numpy(x).to(dev) for x in orc.synthetic_weights(98, bias=True))

    def shifted(x):
        buf = torch.empty(x.numel() + 1, device=dev)
        view = buf[1:].view_as(x)
        view.copy_(x)
        assert view.data_ptr() % 16 == 4 and view.is_contiguous()
        return view
    want = op(v_a, v_b, W, g, b)
    got = op(shifted(v_a), shifted(v_b), W, g, b)
    torch.cuda.synchronize()
    for x, y in zip(got, want):
        assert torch.equal(x, y)
    got = op(shifted(v_a), shifted(v_b), W, g, b, unfused_gate=True)
    ref = op(v_a, v_b, W, g, b, unfused_gate=True)
    torch.cuda.synchronize()
    for x, y in zip(got, ref):
        assert torch.equal(x, y)


@pytest.mark.parametrize("q,refs,h,w,gated", [(3, 5, 12, 11, False), (2, 2, 31, 41, False), (1, 1, 8, 8, False), (2, 3, 20, 20, True)])
def test_grouped_queries_equal_repeated_queries(op, q, refs, h, w, gated):
    """coattn_forward_queries (query side prepared once per query frame, test.py:287-305) == the frame-A forward on the
    query features repeated `refs` times, bit for bit."""
    from cosnet_b200.coattention import coattention_queries_raw
    dev = torch.device("cuda:0")
    v_a = torch.from_numpy(orc.synthetic_features(91, q, h, w, 0.66)[0]).to(dev)
    v_b = torch.from_numpy(orc.synthetic_features(92, q * refs, h, w, 0.66)[1]).to(dev)
    W, g, b = (torch.from_numpy(x).to(dev) for x in orc.synthetic_weights(93, bias=True))
    want = op(v_a.repeat_interleave(refs, 0), v_b, W, g, b, want_z=False, a_only=True, gated_only=gated)[0]
    got = coattention_queries_raw(v_a, v_b, W, g, b, refs=refs, gated_only=gated)
    torch.cuda.synchronize()
    assert got.shape == want.shape and torch.equal(got, want)


@pytest.mark.parametrize("n,h,w", [(2, 12, 11), (1, 60, 60), (1, 31, 41)])
def test_fused_encoder_tail_and_planes_ready(op, n, h, w):
    """SURVEY.md 8f row N4 (producer side): coattn_stage_tail = PReLU(BN_eval(x)) + the 16-bit operand cast in one kernel.
    Its fp32 features equal torch's to rounding, and the co-attention started from its planes (COATTN_FLAG_PLANES_READY)
    equals the ordinary call on those features BIT FOR BIT (the plane is the cast of the same fp32 values)."""
    from cosnet_b200.coattention import bn_eval_affine, coattention_planes_ready, encoder_tail
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    bn = torch.nn.BatchNorm2d(C).to(dev).eval()
    bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 1.5); bn.weight.data.normal_(1, 0.2); bn.bias.data.normal_(0, 0.3)
    prelu = torch.nn.PReLU().to(dev)
    xa, xb = torch.randn(n, C, h, w, device=dev), torch.randn(n, C, h, w, device=dev)
    W, g, b = (torch.from_numpy(t).to(dev) for t in orc.synthetic_weights(21, bias=True))
    scale, shift = bn_eval_affine(bn)
    with torch.no_grad():
        v_a = encoder_tail(xa, scale, shift, prelu.weight, 0, "t")
        v_b = encoder_tail(xb, scale, shift, prelu.weight, 1, "t")
        assert (v_a - prelu(bn(xa))).abs().max() < 2e-6 and (v_b - prelu(bn(xb))).abs().max() < 2e-6
        for gated in (False, True):
            got = coattention_planes_ready(v_a, v_b, W, g, b, "t", gated_only=gated)
            want = op(v_a, v_b, W, g, b, want_z=False, gated_only=gated)
            torch.cuda.synchronize()
            assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
        # features are not needed by a gated-only consumer that keeps them elsewhere: planes only
        assert encoder_tail(xa, scale, shift, prelu.weight, 0, "t2", want_features=False) is None


def test_fused_eval_path_matches_plain_eval_path(op):
    """The drop-in model's eval forward with the fused tail + planes-ready operators + folded split reduce convs (rows N3 /
    N4) against the same model with fuse_eval_path = False: same parameters, outputs equal to fp32 rounding."""
    from cosnet_b200.backbone import Bottleneck
    from cosnet_b200.rgbd_segmentation_raa import RGBDSegmentation_RAA
    dev = torch.device("cuda:0")
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(11)
        model = RGBDSegmentation_RAA(Bottleneck, [1, 1, 1, 1], [1, 1, 1, 1], num_classes=1).to(dev).eval()
        for m in model.modules():           # non-trivial running statistics everywhere
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.05); m.running_var.uniform_(0.8, 1.2)
        x = torch.randn(2, 3, 97, 129, device=dev); d = torch.randn(2, 1, 97, 129, device=dev)
        with torch.no_grad():
            model.fuse_eval_path = True
            got = model(x, x.flip(0), d, d.flip(0))
            model.fuse_eval_path = False
            want = model(x, x.flip(0), d, d.flip(0))
        for a, b_ in zip(got, want):
            assert a.shape == b_.shape
            assert (a - b_).abs().max().item() < 2e-5, (a - b_).abs().max().item()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
