"""CPU oracle for the co-attention hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain numpy restatement of the inline co-attention block of the reference model
(/root/reference/rgbd_segmentation_RAA.py:150-187 for RGB, :204-238 for depth).  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import
this module; the product package `cosnet_b200` never does.

Parity pinning: the reference ships no tests, golden vectors or checkpoints, so
the oracle is pinned against the reference code itself: `oracle/make_golden.py` imports the
unmodified `RGBDSegmentation_RAA` from /root/reference, drives it with stub encoders on seeded
synthetic features and stores its outputs under `tests/golden/`; `tests/test_oracle.py` checks this
restatement against those fixtures (max-abs error ~1e-6 in fp32, see the test for the bound).
"""
from __future__ import annotations

import numpy as np

__all__ = ["softmax", "coattention", "coattention_grads", "synthetic_features", "synthetic_weights"]


def softmax(x: np.ndarray, axis: int) -> np.ndarray:
    """F.softmax semantics (max-subtracted), rgbd_segmentation_RAA.py:164-165."""
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=axis, keepdims=True)


def coattention(v_a, v_b, w, gate_w, gate_b=None, dtype=np.float64):
    """One modality of the hot path.

    v_a, v_b : [N, C, H, W] encoder features of frame A / B        (:143-148 / :198-203)
    w        : [C, C]  `*_similarity_weights.weight` (out, in)      (:27 / :38)
    gate_w   : [C] (or [1, C, 1, 1])  `gate.weight`                 (:28 / :39)
    gate_b   : scalar or None         `depth_gate.bias`             (:39)

    Returns a dict with
      cat_a, cat_b : [N, 2C, H, W]  inputs of reduce_channels_A/B   (:186-187 / :237-238)
      z_a, z_b     : [N, C, H, W]   raw attended features           (:169-170 / :220-221)
      mask_a/b     : [N, 1, H, W]   sigmoid gates                   (:177-182 / :228-233)
      lse_a        : [N, L]  log sum_j exp S[i, j]   (normaliser of S_column, :165)
      lse_b        : [N, L]  log sum_i exp S[i, j]   (normaliser of S_row,    :164)
    """
    v_a = np.asarray(v_a, dtype=dtype)
    v_b = np.asarray(v_b, dtype=dtype)
    w = np.asarray(w, dtype=dtype)
    g = np.asarray(gate_w, dtype=dtype).reshape(-1)
    b = dtype(0) if gate_b is None else dtype(np.asarray(gate_b).reshape(-1)[0])
    n, c, h, wd = v_a.shape
    l = h * wd
    a_flat = v_a.reshape(n, c, l)                      # :154
    b_flat = v_b.reshape(n, c, l)                      # :155
    q = np.matmul(a_flat.transpose(0, 2, 1), w.T)      # :158-159  [N, L, C]  (Linear: x W^T)
    s = np.matmul(q, b_flat)                           # :160      [N, L, L]
    s_row = softmax(s, axis=1)                         # :164  normalise over i (A positions)
    s_col = softmax(s.transpose(0, 2, 1), axis=1)      # :165  [N, Lb, La], normalise over j
    z_b = np.matmul(a_flat, s_row)                     # :169  [N, C, L]
    z_a = np.matmul(b_flat, s_col)                     # :170  [N, C, L]
    smax_i = s.max(axis=2)
    lse_a = smax_i + np.log(np.exp(s - smax_i[:, :, None]).sum(axis=2))
    smax_j = s.max(axis=1)
    lse_b = smax_j + np.log(np.exp(s - smax_j[:, None, :]).sum(axis=1))

    def gate(z):                                       # :177-184
        t = np.einsum("c,ncl->nl", g, z) + b
        m = 1.0 / (1.0 + np.exp(-t))
        return z * m[:, None, :], m

    za_g, mask_a = gate(z_a)
    zb_g, mask_b = gate(z_b)
    cat_a = np.concatenate([za_g, a_flat], axis=1).reshape(n, 2 * c, h, wd)   # :186
    cat_b = np.concatenate([zb_g, b_flat], axis=1).reshape(n, 2 * c, h, wd)   # :187
    return {
        "cat_a": cat_a, "cat_b": cat_b,
        "z_a": z_a.reshape(n, c, h, wd), "z_b": z_b.reshape(n, c, h, wd),
        "mask_a": mask_a.reshape(n, 1, h, wd), "mask_b": mask_b.reshape(n, 1, h, wd),
        "lse_a": lse_a, "lse_b": lse_b,
    }


def coattention_grads(v_a, v_b, w, gate_w, gate_b, d_cat_a, d_cat_b, counterpart_grad=False,
                      dtype=np.float64):
    """Analytic backward of `coattention` with the reference's autograd semantics:

    * mask_b is computed under torch.no_grad() (:178-182) -> it is a constant multiplier: the gate
      parameters receive gradient from the A side only;
    * with no_grad_for_counterpart (:144-148) v_b is a constant: d_v_b is None.

    Returns dict(d_v_a, d_v_b, d_w, d_gate_w, d_gate_b).
    """
    v_a = np.asarray(v_a, dtype=dtype); v_b = np.asarray(v_b, dtype=dtype)
    w = np.asarray(w, dtype=dtype)
    g = np.asarray(gate_w, dtype=dtype).reshape(-1)
    b = dtype(0) if gate_b is None else dtype(np.asarray(gate_b).reshape(-1)[0])
    n, c, h, wd = v_a.shape
    l = h * wd
    A = v_a.reshape(n, c, l); B = v_b.reshape(n, c, l)
    dca = np.asarray(d_cat_a, dtype=dtype).reshape(n, 2 * c, l)
    dcb = np.asarray(d_cat_b, dtype=dtype).reshape(n, 2 * c, l)
    q = np.matmul(A.transpose(0, 2, 1), w.T)           # [N, L, C]
    s = np.matmul(q, B)                                # [N, La, Lb]
    p_b = softmax(s, axis=1)                           # S_row   (normalised over i)
    p_a = softmax(s, axis=2)                           # S_column^T (normalised over j)
    z_b = np.matmul(A, p_b)                            # [N, C, Lb]
    z_a = np.matmul(B, p_a.transpose(0, 2, 1))         # [N, C, La]
    t_a = np.einsum("c,ncl->nl", g, z_a) + b
    m_a = 1.0 / (1.0 + np.exp(-t_a))
    t_b = np.einsum("c,ncl->nl", g, z_b) + b
    m_b = 1.0 / (1.0 + np.exp(-t_b))
    # cat_a = [z_a * m_a, A]; cat_b = [z_b * m_b (const mask), B]
    d_zag = dca[:, :c]                                  # grad wrt gated z_a
    d_zbg = dcb[:, :c]
    d_ta = (d_zag * z_a).sum(axis=1) * m_a * (1 - m_a)  # through the A-side mask only
    d_za = d_zag * m_a[:, None, :] + g[None, :, None] * d_ta[:, None, :]
    d_zb = d_zbg * m_b[:, None, :]
    d_gate_w = np.einsum("nl,ncl->c", d_ta, z_a)
    d_gate_b = d_ta.sum()
    # z_a[:, i] = sum_j B[:, j] p_a[i, j]      z_b[:, j] = sum_i A[:, i] p_b[i, j]
    d_pa = np.matmul(d_za.transpose(0, 2, 1), B)        # [N, La, Lb]
    d_pb = np.matmul(A.transpose(0, 2, 1), d_zb)        # [N, La, Lb]
    d_s = p_a * (d_pa - (d_pa * p_a).sum(axis=2, keepdims=True)) \
        + p_b * (d_pb - (d_pb * p_b).sum(axis=1, keepdims=True))
    d_q = np.matmul(d_s, B.transpose(0, 2, 1))          # [N, La, C]
    d_w = np.einsum("nld,nlc->dc", d_q, A.transpose(0, 2, 1))
    d_v_a = np.matmul(d_q, w).transpose(0, 2, 1) + np.matmul(d_zb, p_b.transpose(0, 2, 1)) + dca[:, c:]
    d_v_b = None
    if counterpart_grad:
        d_v_b = np.matmul(q.transpose(0, 2, 1), d_s) + np.matmul(d_za, p_a) + dcb[:, c:]
        d_v_b = d_v_b.reshape(n, c, h, wd)
    return {"d_v_a": d_v_a.reshape(n, c, h, wd), "d_v_b": d_v_b, "d_w": d_w,
            "d_gate_w": d_gate_w, "d_gate_b": d_gate_b}


def synthetic_features(seed: int, n: int, h: int, w: int, sigma: float, c: int = 256, count: int = 2):
    """Synthetic post-PReLU encoder features: V = prelu_0.25(N(0,1)) * sigma, fp32."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(count):
        x = rng.standard_normal((n, c, h, w), dtype=np.float32)
        out.append((np.where(x >= 0, x, np.float32(0.25) * x) * np.float32(sigma)).astype(np.float32))
    return out


def synthetic_weights(seed: int, c: int = 256, bias: bool = False):
    """Reference ctor defaults: Linear U(+-1/sqrt(C)) (:27, not touched by the N(0,0.01) loop :53-62),
    gate conv N(0, 0.01) (:56), depth_gate bias U(+-1/sqrt(C)) (Conv2d default, :39)."""
    rng = np.random.default_rng(seed)
    k = 1.0 / np.sqrt(c)
    w = rng.uniform(-k, k, size=(c, c)).astype(np.float32)
    g = (rng.standard_normal(c) * 0.01).astype(np.float32)
    b = rng.uniform(-k, k, size=(1,)).astype(np.float32) if bias else None
    return w, g, b
