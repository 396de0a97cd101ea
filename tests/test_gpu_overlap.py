"""RGB and depth co-attention calls on two CUDA streams (`run_modalities`): the drop-in module's eval forward and the
test.py-style inference helper must give bit-identical results with and without the overlap.  `pytest -m gpu`."""
import pytest
import torch

from oracle import coattn_oracle as orc

pytestmark = pytest.mark.gpu


class _Stub(torch.nn.Module):
    """Encoder stand-in that replays queued synthetic features."""

    def __init__(self, as_tuple):
        super().__init__()
        self.queue, self.as_tuple = [], as_tuple

    def forward(self, x):
        f = self.queue.pop(0)
        return (f, x.new_zeros(1)) if self.as_tuple else f


def test_overlap_rule():
    """Overlap exactly when the combined item count needs fewer waves of 74 CTA pairs than the two calls separately."""
    import __graft_entry__ as ge
    ge.build()
    from cosnet_b200.coattention import modality_overlap_pays
    assert modality_overlap_pays(1, 60, 60)            # 30 + 30 items: one wave instead of two (121 -> 65 us measured)
    assert not modality_overlap_pays(4, 60, 60)        # 120 + 120: four waves either way
    assert not modality_overlap_pays(32, 60, 60)       # 960 + 960: 26 waves either way (12.97 waves per call)
    assert modality_overlap_pays(2, 61, 107)           # 104 + 104: three waves instead of four (+27 % measured)
    assert modality_overlap_pays(16, 61, 107)          # 832 + 832: 23 instead of 24
    assert modality_overlap_pays(5, 61, 81, passes=1)  # test.py: one query x 5 references, frame-A items only


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("n,h,w", [(1, 60, 60), (2, 31, 41)])
def test_module_eval_forward_is_identical_with_and_without_overlap(split, n, h, w):
    import __graft_entry__ as ge
    ge.build()
    from cosnet_b200.backbone import Bottleneck
    from cosnet_b200.rgbd_segmentation_raa import RGBDSegmentation_RAA
    dev = torch.device("cuda:0")
    torch.manual_seed(1234)
    model = RGBDSegmentation_RAA(Bottleneck, [1, 1, 1, 1], [1, 1, 1, 1], num_classes=1).eval()
    model.encoder, model.depth_encoder = _Stub(True), _Stub(False)
    model = model.to(dev)
    model.split_reduce_conv = split
    feats = [torch.from_numpy(f).to(dev) for f in orc.synthetic_features(91, n, h, w, 0.66, count=4)]
    img = torch.zeros(n, 3, h * 8, w * 8, device=dev)
    dimg = torch.zeros(n, 1, h * 8, w * 8, device=dev)

    def run(overlap):
        model.overlap_modalities = overlap
        model.encoder.queue = [feats[0], feats[1]]
        model.depth_encoder.queue = [feats[2], feats[3]]
        with torch.no_grad():
            out = model(img, img, dimg, dimg)
        torch.cuda.synchronize()
        return out

    ref = run(False)          # the reference's operator order, one stream
    for overlap in (True, None):
        for _ in range(3):    # repeated: a missing stream dependency would show up as a flaky mismatch
            got = run(overlap)
            assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1])


def test_inference_helper_is_identical_with_and_without_overlap(monkeypatch):
    import __graft_entry__ as ge
    ge.build()
    from cosnet_b200 import inference
    from cosnet_b200.backbone import Bottleneck
    from cosnet_b200.rgbd_segmentation_raa import RGBDSegmentation_RAA
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    model = RGBDSegmentation_RAA(Bottleneck, [1, 1, 1, 1], [1, 1, 1, 1], num_classes=1).to(dev).eval()
    q, r, hw = 1, 5, 97
    tgt, tgt_d = torch.randn(q, 3, hw, hw, device=dev), torch.randn(q, 1, hw, hw, device=dev)
    refs, refs_d = torch.randn(q, r, 3, hw, hw, device=dev), torch.randn(q, r, 1, hw, hw, device=dev)
    monkeypatch.setattr(inference, "modality_overlap_pays", lambda *a, **k: False)
    want = inference.segment_with_references(model, tgt, tgt_d, refs, refs_d)
    monkeypatch.setattr(inference, "modality_overlap_pays", lambda *a, **k: True)
    for _ in range(3):
        got = inference.segment_with_references(model, tgt, tgt_d, refs, refs_d)
        torch.cuda.synchronize()
        assert torch.equal(got, want)


@pytest.mark.parametrize("n,h,w,gated", [(1, 60, 60, False), (2, 12, 11, False), (3, 20, 17, True)])
def test_coattention_pair_equals_two_calls(n, h, w, gated):
    """`coattention_pair` (RGB call on the current stream, depth call on a side stream by handle) returns the bits of two
    plain `coattention()` calls -- also when it is called repeatedly (workspace / event reuse) and on a non-default stream."""
    from cosnet_b200 import coattention, coattention_pair
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev); gen.manual_seed(77 + n)
    feats = [torch.randn(n, 256, h, w, generator=gen, device=dev) * 0.66 for _ in range(4)]
    W = [torch.randn(256, 256, generator=gen, device=dev) / 16 for _ in range(2)]
    G = [torch.randn(256, generator=gen, device=dev) * 0.05 for _ in range(2)]
    b = torch.randn(1, generator=gen, device=dev) * 0.1
    with torch.no_grad():
        ref_rgb = coattention(feats[0], feats[1], W[0], G[0], None, gated_only=gated)
        ref_dep = coattention(feats[2], feats[3], W[1], G[1], b, gated_only=gated)
        for stream in (None, torch.cuda.Stream(dev)):
            ctx = torch.cuda.stream(stream) if stream is not None else torch.no_grad()
            if stream is not None:
                stream.wait_stream(torch.cuda.current_stream(dev))
            with ctx:
                for _ in range(3):
                    rgb, dep = coattention_pair((feats[0], feats[1], W[0], G[0], None), (feats[2], feats[3], W[1], G[1], b),
                                                gated_only=gated)
                    # consumer kernels on the calling stream right behind the call: both results must be complete there
                    s_rgb = [t.clone() for t in rgb]
                    s_dep = [t.clone() for t in dep]
            torch.cuda.synchronize()
            for got, ref in zip(s_rgb + s_dep, list(ref_rgb) + list(ref_dep)):
                assert torch.equal(got, ref)
