"""The numpy oracle against golden vectors produced by the unmodified reference (CPU only)."""
import numpy as np
import pytest

from oracle import coattn_oracle as orc
from tests.helpers import golden_inputs, load_golden, rel_l2, subsample

FWD = ["fwd_n2_4x5_s066", "fwd_n1_3x43_s100"]
BWD = ["bwd_n1_4x5_s066_frozen", "bwd_n1_4x5_s066_both"]

# fp32 reference (MKL sgemm + ATen softmax) vs fp64 numpy restatement
FWD_TOL = 2e-6


@pytest.mark.parametrize("name", FWD)
def test_forward_matches_reference(name):
    fx = load_golden(name)
    inp = golden_inputs(fx)
    rgb = orc.coattention(inp["v_a"], inp["v_b"], inp["w_rgb"], inp["g_rgb"], None)
    dep = orc.coattention(inp["d_a"], inp["d_b"], inp["w_dep"], inp["g_dep"], inp["b_dep"])
    c = 256
    assert rel_l2(rgb["cat_a"][:, :c], fx["rgb_gated_a"]) < FWD_TOL
    assert rel_l2(rgb["cat_b"][:, :c], fx["rgb_gated_b"]) < FWD_TOL
    assert rel_l2(dep["cat_a"][:, :c], fx["depth_gated_a"]) < FWD_TOL
    assert rel_l2(dep["cat_b"][:, :c], fx["depth_gated_b"]) < FWD_TOL
    assert rel_l2(rgb["z_a"], fx["rgb_z_a"]) < FWD_TOL
    assert rel_l2(rgb["z_b"], fx["rgb_z_b"]) < FWD_TOL
    # passthrough half is a bit-exact copy (rgbd_segmentation_RAA.py:186-187)
    assert np.array_equal(rgb["cat_a"][:, c:].astype(np.float32), inp["v_a"])
    assert np.array_equal(dep["cat_b"][:, c:].astype(np.float32), inp["d_b"])


# reference-generated fixtures at the BASELINE.json sizes (strided subsample + full-tensor norms, oracle/make_golden.py)
LARGE = ["fwd_n1_60x60_s066", "fwd_n1_60x60_s100", "fwd_n1_61x81_s066", "fwd_n1_61x107_s066"]


@pytest.mark.parametrize("name", LARGE)
def test_forward_matches_reference_at_baseline_sizes(name):
    fx = load_golden(name)
    inp = golden_inputs(fx)
    c = 256
    for mod, args in {"rgb": (inp["v_a"], inp["v_b"], inp["w_rgb"], inp["g_rgb"], None),
                      "depth": (inp["d_a"], inp["d_b"], inp["w_dep"], inp["g_dep"], inp["b_dep"])}.items():
        out = orc.coattention(*args)
        for side in ("a", "b"):
            gated = out[f"cat_{side}"][:, :c]
            # fp32 reference (MKL sgemm over K = 3600..6527 + ATen softmax) against the fp64 restatement
            assert rel_l2(subsample(gated, fx), fx[f"{mod}_gated_{side}_sub"]) < 2e-5, (mod, side)
            assert abs(np.linalg.norm(gated) / float(fx[f"{mod}_gated_{side}_norm"]) - 1) < 1e-5
        if mod == "rgb":
            assert rel_l2(subsample(out["z_a"], fx), fx["rgb_z_a_sub"]) < 2e-5
            assert rel_l2(subsample(out["z_b"], fx), fx["rgb_z_b_sub"]) < 2e-5


def test_backward_matches_reference_autograd_at_headline_size():
    fx = load_golden("bwd_n1_60x60_s066_frozen")
    inp = golden_inputs(fx)
    n, h, w, seed = int(fx["n"]), int(fx["h"]), int(fx["w"]), int(fx["seed"])
    rng = np.random.default_rng(seed + 7)
    r_a = rng.standard_normal((n, 512, h, w), dtype=np.float32)
    r_b = rng.standard_normal((n, 512, h, w), dtype=np.float32)
    g = orc.coattention_grads(inp["v_a"], inp["v_b"], inp["w_rgb"], inp["g_rgb"], None, r_a, r_b)
    assert rel_l2(subsample(g["d_v_a"], fx), fx["d_v_a_sub"]) < 2e-5
    assert abs(np.linalg.norm(g["d_v_a"]) / float(fx["d_v_a_norm"]) - 1) < 1e-5
    assert rel_l2(g["d_w"], fx["d_w"]) < 2e-5
    assert rel_l2(g["d_gate_w"], fx["d_gate_w"]) < 2e-5


@pytest.mark.parametrize("name", BWD)
def test_backward_matches_reference_autograd(name):
    fx = load_golden(name)
    inp = golden_inputs(fx)
    n, h, w, seed = int(fx["n"]), int(fx["h"]), int(fx["w"]), int(fx["seed"])
    rng = np.random.default_rng(seed + 7)
    r_a = rng.standard_normal((n, 512, h, w), dtype=np.float32)
    r_b = rng.standard_normal((n, 512, h, w), dtype=np.float32)
    frozen = bool(fx["frozen"])
    g = orc.coattention_grads(inp["v_a"], inp["v_b"], inp["w_rgb"], inp["g_rgb"], None, r_a, r_b,
                              counterpart_grad=not frozen)
    assert rel_l2(g["d_v_a"], fx["d_v_a"]) < 1e-5
    assert rel_l2(g["d_w"], fx["d_w"]) < 1e-5
    assert rel_l2(g["d_gate_w"], fx["d_gate_w"]) < 1e-5
    if frozen:
        assert fx["d_v_b"].size == 0 and g["d_v_b"] is None
    else:
        assert rel_l2(g["d_v_b"], fx["d_v_b"]) < 1e-5


def test_lse_consistency():
    v_a, v_b = orc.synthetic_features(5, 1, 3, 4, 1.0)
    w, g, b = orc.synthetic_weights(6, bias=True)
    out = orc.coattention(v_a, v_b, w, g, b)
    a = v_a.reshape(1, 256, 12).astype(np.float64)
    bb = v_b.reshape(1, 256, 12).astype(np.float64)
    s = np.matmul(np.matmul(a.transpose(0, 2, 1), w.T.astype(np.float64)), bb)
    assert np.allclose(np.exp(s - out["lse_a"][:, :, None]).sum(axis=2), 1.0)
    assert np.allclose(np.exp(s - out["lse_b"][:, None, :]).sum(axis=1), 1.0)


def test_degenerate_single_position():
    # L = 1: both softmaxes are identically 1 -> Z_a = V_b, Z_b = V_a
    v_a, v_b = orc.synthetic_features(9, 2, 1, 1, 1.0)
    w, g, _ = orc.synthetic_weights(10)
    out = orc.coattention(v_a, v_b, w, g, None)
    assert np.allclose(out["z_a"], v_b)
    assert np.allclose(out["z_b"], v_a)
