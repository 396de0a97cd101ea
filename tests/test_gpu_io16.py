"""16-bit feature interface (`coattn_forward16`, SURVEY.md 8(b) "fp32 (or bf16)" / 8(f) N4) against the fp32 interface
and the CPU oracle.  Needs a B200: run with `pytest -m gpu`.

The 16-bit path performs the arithmetic of `coattn_forward` on the same 16-bit values (the fp32 entry point casts its
features to exactly these values), so its outputs must be the fp32 entry point's outputs rounded once to 16 bits --
BIT FOR BIT -- whether TMA reads the features in place (L % 8 == 0) or from padded copies.
"""
import numpy as np
import pytest
import torch

from oracle import coattn_oracle as orc
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu

C = 256
TOL = 1e-3      # fp16 features and outputs, module output vs the fp64 oracle on the ORIGINAL fp32 features
TOL_BF = 1e-2   # bf16 features and outputs: 2^-9 rounding of inputs, operands and outputs


@pytest.fixture(scope="module")
def ops():
    import __graft_entry__ as ge
    ge.build()
    from cosnet_b200.coattention import coattention_forward16_raw, coattention_forward_raw
    assert torch.cuda.is_available()
    return coattention_forward_raw, coattention_forward16_raw


def _inputs(seed, n, h, w, dtype, bias=True, sigma=0.66):
    dev = torch.device("cuda:0")
    v_a, v_b = orc.synthetic_features(seed, n, h, w, sigma)
    W, g, b = orc.synthetic_weights(seed + 1, bias=bias)
    t = lambda x: None if x is None else torch.from_numpy(x).to(dev)
    return (v_a, v_b, W, g, b), (t(v_a).to(dtype), t(v_b).to(dtype), t(W), t(g), t(b))


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("n,h,w", [
    (2, 12, 12),    # L = 144: read in place, ragged key and query tiles
    (1, 16, 16),    # L = 256: read in place, exactly one query tile pair
    (3, 20, 20),    # L = 400: read in place, odd batch
    (2, 12, 11),    # L = 132: L % 8 != 0 -> padded copies
    (1, 7, 9),      # L = 63 : odd L
    (1, 40, 47),    # L = 1880: in place, several query tiles
])
def test_io16_is_the_fp32_interface_rounded_once(ops, dtype, n, h, w):
    fwd32, fwd16 = ops
    _, (a16, b16, W, g, b) = _inputs(31, n, h, w, dtype)
    bf = dtype == torch.bfloat16
    want_a, want_b, _, lse32, mask32 = fwd32(a16.float(), b16.float(), W, g, b, bf16_operands=bf, want_mask=True)
    got_a, got_b, lse, mask = fwd16(a16, b16, W, g, b, want_lse=True)
    torch.cuda.synchronize()
    assert got_a.dtype == dtype and got_a.shape == want_a.shape
    assert torch.equal(got_a, want_a.to(dtype)) and torch.equal(got_b, want_b.to(dtype))
    assert torch.equal(got_a[:, C:], a16) and torch.equal(got_b[:, C:], b16)       # passthrough half: bit copy
    assert torch.equal(lse, lse32) and torch.equal(mask, mask32)


@pytest.mark.parametrize("dtype,tol", [(torch.float16, TOL), (torch.bfloat16, TOL_BF)])
@pytest.mark.parametrize("n,h,w,bias", [(2, 12, 12, False), (1, 31, 41, True), (1, 60, 60, True)])
def test_io16_against_oracle(ops, dtype, tol, n, h, w, bias):
    _, fwd16 = ops
    (v_a, v_b, W, g, b), (a16, b16, tW, tg, tb) = _inputs(47, n, h, w, dtype, bias=bias)
    ref = orc.coattention(v_a, v_b, W, g, b)
    got_a, got_b = fwd16(a16, b16, tW, tg, tb)
    torch.cuda.synchronize()
    ea, eb = rel_l2(got_a.float().cpu().numpy(), ref["cat_a"]), rel_l2(got_b.float().cpu().numpy(), ref["cat_b"])
    assert ea < tol and eb < tol, (ea, eb)


@pytest.mark.parametrize("q,refs,h,w,gated", [(3, 5, 12, 12, False), (2, 2, 31, 41, False), (2, 3, 20, 20, True)])
def test_io16_grouped_queries_and_gated_only(ops, q, refs, h, w, gated):
    """refs > 1 (test.py:287-305): pair p = (query p // refs, reference p), frame-A outputs only."""
    _, fwd16 = ops
    dev = torch.device("cuda:0")
    v_a = torch.from_numpy(orc.synthetic_features(91, q, h, w, 0.66)[0]).to(dev).half()
    v_b = torch.from_numpy(orc.synthetic_features(92, q * refs, h, w, 0.66)[1]).to(dev).half()
    W, g, b = (torch.from_numpy(x).to(dev) for x in orc.synthetic_weights(93, bias=True))
    want, none_b = fwd16(v_a.repeat_interleave(refs, 0), v_b, W, g, b, a_only=True, gated_only=gated)
    got, got_b = fwd16(v_a, v_b, W, g, b, refs=refs, gated_only=gated)
    torch.cuda.synchronize()
    assert none_b is None and got_b is None
    assert got.shape == (q * refs, C if gated else 2 * C, h, w) and torch.equal(got, want)
    full_a, _ = fwd16(v_a.repeat_interleave(refs, 0), v_b, W, g, b)
    assert torch.equal(got[:, :C], full_a[:, :C])


def test_io16_unaligned_views_take_the_copy_path(ops):
    """A feature tensor whose base is not 16-byte aligned cannot be a TMA source: same bits through the padded copies."""
    _, fwd16 = ops
    _, (a16, b16, W, g, b) = _inputs(5, 2, 16, 16, torch.float16)
    want_a, want_b = fwd16(a16, b16, W, g, b)

    def shifted(x):      # same values, base pointer moved by 2 bytes
        buf = torch.empty(x.numel() + 1, dtype=x.dtype, device=x.device)
        v = buf[1:].view(x.shape)
        v.copy_(x)
        assert v.data_ptr() % 16 != 0 and v.is_contiguous()
        return v
    got_a, got_b = fwd16(shifted(a16), shifted(b16), W, g, b)
    torch.cuda.synchronize()
    assert torch.equal(got_a, want_a) and torch.equal(got_b, want_b)


def test_io16_argument_checks(ops):
    from cosnet_b200 import _lib
    _, fwd16 = ops
    _, (a16, b16, W, g, b) = _inputs(5, 1, 8, 8, torch.float16)
    with pytest.raises(TypeError):
        fwd16(a16.float(), b16.float(), W, g, b)
    with pytest.raises(TypeError):
        fwd16(a16, b16.bfloat16(), W, g, b)
    with pytest.raises(ValueError):
        fwd16(a16, torch.cat([b16, b16, b16]), W, g, b, refs=2)
    with pytest.raises(_lib.CoattnError):
        fwd16(a16.cpu(), b16.cpu(), W, g, b)
    lib = _lib.load()
    ws = torch.empty(lib.coattn_workspace_bytes(1, C, 8, 8) + 1024, dtype=torch.uint8, device="cuda:0")
    wp = (ws.data_ptr() + 1023) // 1024 * 1024
    out = torch.empty((1, 2 * C, 8, 8), dtype=torch.float16, device="cuda:0")
    args = lambda flags, cat_b=out.data_ptr(): (a16.data_ptr(), b16.data_ptr(), W.data_ptr(), g.data_ptr(), None,
                                               out.data_ptr(), cat_b, None, None, wp, ws.numel() - 1024, 1, 1, C, 8, 8,
                                               flags, None)
    assert lib.coattn_forward16(*args(_lib.FLAG_UNFUSED_GATE)) == -7
    assert lib.coattn_forward16(*args(_lib.FLAG_SINGLE_CTA)) == -7
    assert lib.coattn_forward16(*args(0, None)) == -1          # cat_b required unless A_ONLY
    assert lib.coattn_forward16(*args(_lib.FLAG_A_ONLY, None)) == 0
    torch.cuda.synchronize()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("gated", [False, True])
def test_io16_host_pipeline_matches_resident_path(ops, dtype, gated):
    """HostPipeline(feature_dtype=16-bit): pinned 16-bit host buffers in and out, chunks over three streams."""
    from cosnet_b200.coattention import HostPipeline
    _, fwd16 = ops
    n, h, w = 7, 16, 12      # 7 pairs in chunks of 2: a ragged last chunk
    _, (a16, b16, W, g, b) = _inputs(77, n, h, w, dtype)
    want_a, want_b = fwd16(a16, b16, W, g, b, gated_only=gated)
    pipe = HostPipeline(n, C, h, w, chunk=2, slots=3, device="cuda:0", feature_dtype=dtype, gated_only=gated)
    oc = C if gated else 2 * C
    out_a = torch.empty((n, oc, h, w), dtype=dtype).pin_memory()
    out_b = torch.empty((n, oc, h, w), dtype=dtype).pin_memory()
    pipe(a16.cpu().pin_memory(), b16.cpu().pin_memory(), W, g, b, out_a, out_b)
    torch.cuda.synchronize()
    assert pipe.h2d_bytes == 2 * n * C * h * w * 2 and pipe.d2h_bytes == 2 * n * oc * h * w * 2
    assert torch.equal(out_a, want_a.cpu()) and torch.equal(out_b, want_b.cpu())


class _Stub(torch.nn.Module):
    """Encoder stand-in that replays queued synthetic features (as in tests/test_gpu_parity.py)."""

    def __init__(self, as_tuple):
        super().__init__()
        self.queue, self.as_tuple = [], as_tuple

    def forward(self, x):
        f = self.queue.pop(0)
        return (f, x.new_zeros(1)) if self.as_tuple else f


@pytest.mark.parametrize("split", [False, True])
def test_module_with_16bit_operator_and_half_precision_module(ops, split):
    """(1) The fp32 drop-in module with the co-attention block swapped for the 16-bit interface (features rounded to
    fp16 on the way in, fp16 concat widened on the way out): binarised masks agree with the fp32 operator on >= 99.9 %
    of pixels (BASELINE.json's bar).  (2) `model.half()` -- what an fp16 deployment of test.py runs: the encoders hand
    over fp16 features, the module routes them through coattn_forward16 and the fp16 concat feeds fp16 reduce convs; the
    remaining distance to the fp32 model is that of the half-precision convolutions around the block."""
    from cosnet_b200.backbone import Bottleneck
    from cosnet_b200.coattention import coattention
    from cosnet_b200.rgbd_segmentation_raa import RGBDSegmentation_RAA
    dev = torch.device("cuda:0")
    torch.manual_seed(1234)
    model = RGBDSegmentation_RAA(Bottleneck, [1, 1, 1, 1], [1, 1, 1, 1], num_classes=1).eval()
    model.encoder, model.depth_encoder = _Stub(True), _Stub(False)
    model = model.to(dev)
    model.split_reduce_conv = split
    n, h, w = 2, 40, 40
    feats = [torch.from_numpy(f).to(dev) for f in orc.synthetic_features(91, n, h, w, 0.66, count=4)]
    img = torch.zeros(n, 3, h * 8, w * 8, device=dev)
    dimg = torch.zeros(n, 1, h * 8, w * 8, device=dev)

    def run_model(m, dtype):
        m.encoder.queue = [feats[0].to(dtype), feats[1].to(dtype)]
        m.depth_encoder.queue = [feats[2].to(dtype), feats[3].to(dtype)]
        with torch.no_grad():
            return m(img.to(dtype), img.to(dtype), dimg.to(dtype), dimg.to(dtype))

    r1, r2, _ = run_model(model, torch.float32)

    def impl16(v_a, v_b, weight, gate_weight, gate_bias, gated_only=False):
        a, b = coattention(v_a.half(), v_b.half(), weight, gate_weight, gate_bias, gated_only=gated_only)
        assert a.dtype == torch.float16
        return a.float(), b.float()
    model.coattention_impl = impl16
    x1, x2, _ = run_model(model, torch.float32)
    for got, ref in ((x1, r1), (x2, r2)):
        assert float(ref.min()) < 0.5 < float(ref.max())      # only meaningful if the maps straddle the threshold
        agree = ((got > 0.5) == (ref > 0.5)).float().mean().item()
        assert agree >= 0.999, agree
        assert (got - ref).abs().max().item() < 2e-3

    model.coattention_impl = coattention
    h1, h2, _ = run_model(model.half(), torch.float16)
    for got, ref in ((h1, r1), (h2, r2)):
        assert got.dtype == torch.float16 and got.shape == ref.shape and bool(torch.isfinite(got).all())
        assert (got.float() - ref).abs().max().item() < 1e-2


def test_16bit_features_with_gradients(ops):
    """Mixed-precision training: fp16 features that require grad go through the fp32 interface on the widened values --
    same forward bits as coattn_forward16, gradients those of the fp32 path rounded to fp16."""
    from cosnet_b200.coattention import coattention
    _, fwd16 = ops
    _, (a16, b16, W, g, b) = _inputs(63, 2, 12, 12, torch.float16)
    want_a, want_b = fwd16(a16, b16, W, g, b)
    W1, g1, b1 = (t.clone().requires_grad_(True) for t in (W, g, b))
    a1 = a16.clone().requires_grad_(True)
    cat_a, cat_b = coattention(a1, b16, W1, g1, b1)
    assert cat_a.dtype == torch.float16 and torch.equal(cat_a, want_a) and torch.equal(cat_b, want_b)
    ra, rb = torch.randn_like(cat_a), torch.randn_like(cat_b)
    ((cat_a * ra).sum() + (cat_b * rb).sum()).backward()
    W2, g2, b2 = (t.clone().requires_grad_(True) for t in (W, g, b))
    a2 = a16.float().requires_grad_(True)
    ca2, cb2 = coattention(a2, b16.float(), W2, g2, b2)
    ((ca2 * ra.float()).sum() + (cb2 * rb.float()).sum()).backward()
    torch.cuda.synchronize()
    # (the backward accumulates d_w / d_gate_w with floating-point atomics: equal to rounding, not bit for bit)
    rel = lambda x, y: float((x.float() - y.float()).norm() / y.float().norm())
    assert a1.grad.dtype == torch.float16 and rel(a1.grad, a2.grad) < 1e-3
    for x, y in ((W1, W2), (g1, g2), (b1, b2)):
        assert rel(x.grad, y.grad) < 1e-3


def test_half_precision_multi_reference_inference():
    """test.py:278-305 with `model.half()`: the query-hoisted batched path (coattn_forward16 with refs > 1) equals running
    the half-precision drop-in module once per (query, reference) pair."""
    from cosnet_b200.backbone import Bottleneck
    from cosnet_b200.inference import segment_with_references
    from cosnet_b200.rgbd_segmentation_raa import RGBDSegmentation_RAA
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    model = RGBDSegmentation_RAA(Bottleneck, [1, 1, 1, 1], [1, 1, 1, 1], num_classes=1).to(dev).eval().half()
    q, r, hw = 2, 3, 97
    tgt, tgt_d = torch.randn(q, 3, hw, hw, device=dev).half(), torch.randn(q, 1, hw, hw, device=dev).half()
    refs, refs_d = torch.randn(q, r, 3, hw, hw, device=dev).half(), torch.randn(q, r, 1, hw, hw, device=dev).half()
    got = segment_with_references(model, tgt, tgt_d, refs, refs_d)
    want = torch.zeros_like(got, dtype=torch.float32)
    with torch.no_grad():
        for i in range(r):                                       # the reference's loop (test.py:287-301)
            want += model(tgt, refs[:, i], tgt_d, refs_d[:, i])[0].float()
    want /= r
    assert got.dtype == torch.float16 and got.shape == (q, 1, hw, hw)
    # the encoders see other batch compositions (cuDNN picks other fp16 algorithms): equal to half-precision rounding
    assert (got.float() - want).abs().max().item() < 5e-3


def test_forward16_is_cuda_graph_capturable(ops):
    """3 launches (cast_w, project_mn, attend2), no host synchronisation, no allocation: capturable, and the tensor maps
    over the caller's feature tensors stay valid for replays on new contents of the same buffers."""
    from cosnet_b200 import _lib
    from cosnet_b200.coattention import workspace_bytes
    _, fwd16 = ops
    dev = torch.device("cuda:0")
    lib = _lib.load()
    n, c, h, w = 2, 256, 12, 12
    v_a = torch.empty(n, c, h, w, device=dev, dtype=torch.float16)
    v_b = torch.empty_like(v_a)
    W, g, b = (torch.from_numpy(x).to(dev) for x in orc.synthetic_weights(71, bias=True))
    cat_a = torch.empty(n, 2 * c, h, w, device=dev, dtype=torch.float16)
    cat_b = torch.empty_like(cat_a)
    nbytes = workspace_bytes(n, c, h, w)
    ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
    wsp = (ws.data_ptr() + 1023) // 1024 * 1024

    def call(stream):
        _lib.check(lib.coattn_forward16(v_a.data_ptr(), v_b.data_ptr(), W.data_ptr(), g.data_ptr(), b.data_ptr(),
                                        cat_a.data_ptr(), cat_b.data_ptr(), None, None, wsp, nbytes, n, 1, c, h, w, 0,
                                        stream.cuda_stream), "coattn_forward16")

    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        call(side)                                  # warm-up outside the capture (function attributes)
    torch.cuda.current_stream(dev).wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        call(torch.cuda.current_stream(dev))
    for seed in (72, 73):
        fa, fb = (torch.from_numpy(x).to(dev).half() for x in orc.synthetic_features(seed, n, h, w, 0.66))
        v_a.copy_(fa); v_b.copy_(fb)
        graph.replay()
        torch.cuda.synchronize()
        want = fwd16(fa, fb, W, g, b)
        assert torch.equal(cat_a, want[0]) and torch.equal(cat_b, want[1])
