"""GraphedCoAttention (CUDA-graph replay of the RGB + depth modality calls, two streams when that pays) against the eager
operator: bit-identical, replayable on new inputs, weights read at replay time.  `pytest -m gpu`."""
import pytest
import torch

from oracle import coattn_oracle as orc

pytestmark = pytest.mark.gpu


def _weights(dev):
    Wr, gr, _ = orc.synthetic_weights(6, bias=False)
    Wd, gd, bd = orc.synthetic_weights(7, bias=True)
    t = lambda x: torch.from_numpy(x).to(dev)
    return (t(Wr), t(gr), None), (t(Wd), t(gd), t(bd))


@pytest.mark.parametrize("n,h,w,overlap", [(1, 60, 60, None), (2, 31, 41, True), (3, 12, 12, False)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_graph_replay_equals_eager(n, h, w, overlap, dtype):
    import __graft_entry__ as ge
    ge.build()
    from cosnet_b200.coattention import coattention_forward16_raw, coattention_forward_raw
    from cosnet_b200.graphed import GraphedCoAttention
    dev = torch.device("cuda:0")
    rgb, depth = _weights(dev)
    g = GraphedCoAttention(n, h, w, rgb, depth, dtype=dtype, device=dev, overlap=overlap)
    if overlap is None:
        assert g.overlap            # one 60x60 pair: 30 + 30 work items share the 74 CTA pairs

    def eager(v_a, v_b, p):
        if dtype == torch.float32:
            return coattention_forward_raw(v_a, v_b, *p, want_z=False)[:2]
        return coattention_forward16_raw(v_a, v_b, *p)

    for seed in (11, 12, 13):
        f = [torch.from_numpy(x).to(dev).to(dtype) for x in orc.synthetic_features(seed, n, h, w, 0.66, count=4)]
        cat_a, cat_b, dcat_a, dcat_b = g(*f)
        torch.cuda.synchronize()
        wa, wb = eager(f[0], f[1], rgb)
        da, db = eager(f[2], f[3], depth)
        assert torch.equal(cat_a, wa) and torch.equal(cat_b, wb) and torch.equal(dcat_a, da) and torch.equal(dcat_b, db)
    # the graph reads the weights from the parameters' storage: an in-place update is seen by the next replay
    rgb[0].mul_(0.5)
    cat_a, cat_b, _, _ = g.replay()
    torch.cuda.synchronize()
    wa, wb = eager(f[0], f[1], rgb)
    assert torch.equal(cat_a, wa) and torch.equal(cat_b, wb)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_graph_replay_grouped_queries(dtype):
    """test.py's shape: one query frame against 5 reference frames, frame-A outputs only."""
    import __graft_entry__ as ge
    ge.build()
    from cosnet_b200.coattention import coattention_forward16_raw, coattention_queries_raw
    from cosnet_b200.graphed import GraphedCoAttention
    dev = torch.device("cuda:0")
    rgb, depth = _weights(dev)
    q, r, h, w = 1, 5, 24, 24
    g = GraphedCoAttention(q * r, h, w, rgb, depth, refs=r, dtype=dtype, device=dev)
    assert g.cat_b is None
    va, da = (torch.from_numpy(x).to(dev).to(dtype) for x in orc.synthetic_features(3, q, h, w, 0.66))
    vb, db = (torch.from_numpy(x).to(dev).to(dtype) for x in orc.synthetic_features(4, q * r, h, w, 0.66))
    cat_a, _, dcat_a, _ = g(va, vb, da, db)
    torch.cuda.synchronize()
    if dtype == torch.float32:
        wa = coattention_queries_raw(va, vb, *rgb, refs=r)
        wd = coattention_queries_raw(da, db, *depth, refs=r)
    else:
        wa = coattention_forward16_raw(va, vb, *rgb, refs=r)[0]
        wd = coattention_forward16_raw(da, db, *depth, refs=r)[0]
    assert torch.equal(cat_a, wa) and torch.equal(dcat_a, wd)


def test_whole_model_eval_graph_matches_eager():
    """GraphedEvalModel: the drop-in model's eval forward (fused tails, planes-ready operators, folded reduce convs) captured
    once and replayed on new inputs equals the eager forward."""
    from cosnet_b200.backbone import Bottleneck
    from cosnet_b200.graphed import GraphedEvalModel
    from cosnet_b200.rgbd_segmentation_raa import RGBDSegmentation_RAA
    dev = torch.device("cuda:0")
    torch.manual_seed(21)
    model = RGBDSegmentation_RAA(Bottleneck, [1, 1, 1, 1], [1, 1, 1, 1], num_classes=1).to(dev).eval()
    mk = lambda c: torch.randn(1, c, 97, 97, device=dev)
    gm = GraphedEvalModel(model, mk(3), mk(3), mk(1), mk(1))
    for _ in range(2):
        x = (mk(3), mk(3), mk(1), mk(1))
        got = [t.clone() for t in gm(*x)]
        with torch.no_grad():
            want = model(*x)
        torch.cuda.synchronize()
        for a, b in zip(got, want):
            assert torch.isfinite(a).all() and (a - b).abs().max().item() < 1e-5
