"""Out-of-bounds canaries around every output buffer of the C ABI (compute-sanitizer is not available on the pool):
each output lives inside a larger allocation whose guard bands are filled with a sentinel; after forward and backward
the guard bands must be untouched.  Ragged / odd shapes on purpose."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import coattn_oracle as orc

pytestmark = pytest.mark.gpu
GUARD = 4096          # floats on each side
SENT = 12345.678


class Guarded:
    def __init__(self, numel, dev):
        self.buf = torch.full((numel + 2 * GUARD,), SENT, dtype=torch.float32, device=dev)
        self.numel = numel

    @property
    def ptr(self):
        return self.buf.data_ptr() + GUARD * 4

    def view(self, *shape):
        return self.buf[GUARD:GUARD + self.numel].view(*shape)

    def intact(self):
        return bool((self.buf[:GUARD] == SENT).all() and (self.buf[GUARD + self.numel:] == SENT).all())


@pytest.mark.parametrize("n,h,w,flags", [(1, 1, 1, 0), (2, 7, 9, 0), (1, 31, 41, 0), (1, 13, 20, 1), (2, 12, 11, 4), (1, 16, 16, 16),
                                         (1, 12, 11, 2), (1, 31, 41, 256), (1, 60, 60, 256)])      # 256: split key ranges
def test_forward_and_backward_stay_inside_their_buffers(n, h, w, flags):
    import __graft_entry__ as ge
    ge.build()
    from cosnet_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    c, L = 256, h * w
    v_a, v_b = (torch.from_numpy(x).to(dev) for x in orc.synthetic_features(900 + L, n, h, w, 0.66))
    W, g, b = (torch.from_numpy(x).to(dev) for x in orc.synthetic_weights(901, bias=True))
    cat_a, cat_b = Guarded(n * 2 * c * L, dev), Guarded(n * 2 * c * L, dev)
    z, lse, mask = Guarded(2 * n * c * L, dev), Guarded(2 * n * L, dev), Guarded(2 * n * L, dev)
    nbytes = lib.coattn_workspace_bytes(n, c, h, w)
    ws = Guarded(nbytes // 4 + 512, dev)
    wsp = (ws.ptr + 1023) // 1024 * 1024
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.coattn_forward(v_a.data_ptr(), v_b.data_ptr(), W.data_ptr(), g.data_ptr(), b.data_ptr(), cat_a.ptr, cat_b.ptr,
                            z.ptr, lse.ptr, mask.ptr, wsp, nbytes, n, c, h, w, flags, st)
    assert rc == 0, lib.coattn_b200_strerror(rc)
    torch.cuda.synchronize()
    for name, gbuf in (("cat_a", cat_a), ("cat_b", cat_b), ("z", z), ("lse", lse), ("mask", mask), ("workspace", ws)):
        assert gbuf.intact(), f"forward wrote outside {name}"
    ref = orc.coattention(v_a.cpu().numpy(), v_b.cpu().numpy(), W.cpu().numpy(), g.cpu().numpy(), b.cpu().numpy())
    err = np.linalg.norm(cat_a.view(n, 2 * c, h, w).cpu().numpy() - ref["cat_a"]) / np.linalg.norm(ref["cat_a"])
    assert err < (5e-3 if flags & 1 else 1e-3)
    if flags & 2:      # the unfused gate path does not produce `mask`; the backward needs it
        return
    # backward (with counterpart gradients) through the same guarded buffers
    d_cat_a = torch.randn(n, 2 * c, h, w, device=dev)
    d_cat_b = torch.randn(n, 2 * c, h, w, device=dev)
    d_va, d_vb = Guarded(n * c * L, dev), Guarded(n * c * L, dev)
    d_w, d_gw, d_gb = Guarded(c * c, dev), Guarded(c, dev), Guarded(1, dev)
    bbytes = lib.coattn_backward_workspace_bytes(n, c, h, w, 1)
    bws = Guarded(bbytes // 4 + 512, dev)
    bwsp = (bws.ptr + 1023) // 1024 * 1024
    rc = lib.coattn_backward(v_a.data_ptr(), v_b.data_ptr(), W.data_ptr(), g.data_ptr(), z.ptr, lse.ptr, mask.ptr,
                             d_cat_a.data_ptr(), d_cat_b.data_ptr(), d_va.ptr, d_vb.ptr, d_w.ptr, d_gw.ptr, d_gb.ptr,
                             bwsp, bbytes, n, c, h, w, flags & 1, st)
    assert rc == 0, lib.coattn_b200_strerror(rc)
    torch.cuda.synchronize()
    for name, gbuf in (("d_v_a", d_va), ("d_v_b", d_vb), ("d_w", d_w), ("d_gate_w", d_gw), ("d_gate_b", d_gb), ("workspace", bws),
                       ("z", z), ("lse", lse), ("mask", mask)):
        assert gbuf.intact(), f"backward wrote outside {name}"
    gr = orc.coattention_grads(v_a.cpu().numpy(), v_b.cpu().numpy(), W.cpu().numpy(), g.cpu().numpy(), b.cpu().numpy(),
                               d_cat_a.cpu().numpy(), d_cat_b.cpu().numpy(), counterpart_grad=True)
    rel = lambda x, r: float(np.linalg.norm(x - r) / max(np.linalg.norm(r), 1e-30))
    assert np.isfinite(d_w.view(c, c).cpu().numpy()).all()
    if L < 16:
        # degenerate sizes (L = 1: both softmaxes are identically 1 and the exact dS is 0): dP - delta cancels only up
        # to the bf16 rounding of the GEMM operands, so a relative comparison against ~0 is meaningless here
        return
    assert rel(d_va.view(n, c, h, w).cpu().numpy(), gr["d_v_a"]) < 1e-2
    assert rel(d_vb.view(n, c, h, w).cpu().numpy(), gr["d_v_b"]) < 1e-2
    assert rel(d_w.view(c, c).cpu().numpy(), gr["d_w"]) < 1.5e-2


class Guarded16:
    """The same for 16-bit buffers (sentinel bit pattern 0x7BFF: the largest finite half)."""
    SENT16 = 0x7BFF

    def __init__(self, numel, dev):
        self.buf = torch.full((numel + 2 * GUARD,), self.SENT16, dtype=torch.int16, device=dev)
        self.numel = numel

    @property
    def ptr(self):
        return self.buf.data_ptr() + GUARD * 2

    def view(self, dtype, *shape):
        return self.buf[GUARD:GUARD + self.numel].view(dtype).view(*shape)

    def intact(self):
        return bool((self.buf[:GUARD] == self.SENT16).all() and (self.buf[GUARD + self.numel:] == self.SENT16).all())


@pytest.mark.parametrize("nq,refs,h,w,flags", [
    (2, 1, 12, 12, 0),     # features read in place by TMA (L % 8 == 0): boxes reach past L and past the last row block
    (1, 1, 16, 16, 1),     # bf16
    (2, 1, 7, 9, 0),       # odd L: padded copies
    (1, 1, 31, 41, 32),    # gated half only: 256-channel outputs
    (2, 3, 20, 20, 8),     # grouped queries, frame-A outputs only
])
def test_forward16_stays_inside_its_buffers(nq, refs, h, w, flags):
    import __graft_entry__ as ge
    ge.build()
    from cosnet_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    c, L, n = 256, h * w, nq * refs
    dt = torch.bfloat16 if flags & 1 else torch.float16
    fa = torch.from_numpy(orc.synthetic_features(700 + L, nq, h, w, 0.66)[0]).to(dev).to(dt)
    fb = torch.from_numpy(orc.synthetic_features(701 + L, n, h, w, 0.66)[1]).to(dev).to(dt)
    # the inputs sit in guarded allocations too: a TMA box that reaches past the tensor must read zeros, not neighbours
    ga, gb = Guarded16(fa.numel(), dev), Guarded16(fb.numel(), dev)
    ga.view(dt, *fa.shape).copy_(fa); gb.view(dt, *fb.shape).copy_(fb)
    W, g, b = (torch.from_numpy(x).to(dev) for x in orc.synthetic_weights(901, bias=True))
    oc = c if flags & 32 else 2 * c
    a_only = bool(flags & 8) or refs > 1
    cat_a, cat_b = Guarded16(n * oc * L, dev), Guarded16(n * oc * L, dev)
    passes = 1 if a_only else 2
    lse, mask = Guarded(passes * n * L, dev), Guarded(passes * n * L, dev)
    nbytes = lib.coattn_workspace_bytes(n, c, h, w)
    ws = Guarded(nbytes // 4 + 512, dev)
    wsp = (ws.ptr + 1023) // 1024 * 1024
    rc = lib.coattn_forward16(ga.ptr, gb.ptr, W.data_ptr(), g.data_ptr(), b.data_ptr(), cat_a.ptr,
                              None if a_only else cat_b.ptr, lse.ptr, mask.ptr, wsp, nbytes, nq, refs, c, h, w, flags,
                              torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.coattn_b200_strerror(rc)
    torch.cuda.synchronize()
    for name, gbuf in (("v_a", ga), ("v_b", gb), ("cat_a", cat_a), ("cat_b", cat_b), ("lse", lse), ("mask", mask), ("workspace", ws)):
        assert gbuf.intact(), f"coattn_forward16 wrote outside {name}"
    ref = orc.coattention(fa.float().repeat_interleave(refs, 0).cpu().numpy(), fb.float().cpu().numpy(), W.cpu().numpy(),
                          g.cpu().numpy(), b.cpu().numpy())
    got = cat_a.view(dt, n, oc, h, w).float().cpu().numpy()
    want = ref["cat_a"][:, :oc]
    err = np.linalg.norm(got - want) / np.linalg.norm(want)
    assert err < (1e-2 if flags & 1 else 1e-3), err
