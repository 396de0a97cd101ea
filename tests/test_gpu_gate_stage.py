"""The stand-alone gate / sigmoid / scale / concat epilogue (`coattn_stage_gate`, rgbd_segmentation_RAA.py:175-187 and :226-238)
against a plain PyTorch fp32 restatement of the same lines: the passthrough half bit for bit, the gated half to fp32 rounding
of the 256-term gate logit.  Shapes cover the vectorised kernel (L even: full 64-position tiles, ragged tail tiles) and
the scalar fallback (odd L, views that are only 4-byte aligned).  (Written for the TMA-pipeline variant of the kernel, which is kept as
profiles/r2_gate_tma_pipeline.patch: correct, but 0.78 of the copy bandwidth against the register-tile kernel's 0.92.)"""
import pytest
import torch

from cosnet_b200 import _lib

pytestmark = pytest.mark.gpu
C = 256


def _reference(z, v_a, v_b, g, b):
    outs = []
    for side, v in enumerate((v_a, v_b)):
        zz = z[side]
        logit = torch.einsum("c,ncl->nl", g.double(), zz.double()) + (0.0 if b is None else b.double())
        mask = torch.sigmoid(logit).float().unsqueeze(1)
        outs.append(torch.cat([zz * mask, v.flatten(2)], 1))
    return outs


def _run(n, h, w, bias, offset=0):
    lib = _lib.load()
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev); gen.manual_seed(1000 * n + 10 * h + w)
    L = h * w
    def buf(*shape):      # offset > 0: a 4-byte-aligned view that is not 16-byte aligned
        t = torch.randn(int(torch.tensor(shape).prod()) + offset, generator=gen, device=dev)
        return t[offset:].view(*shape)
    z, v_a, v_b = buf(2, n, C, L), buf(n, C, h, w), buf(n, C, h, w)
    g = torch.randn(C, generator=gen, device=dev) * 0.1
    b = torch.randn(1, generator=gen, device=dev) if bias else None
    cat_a = torch.full((n, 2 * C, h, w), float("nan"), device=dev)
    cat_b = torch.full((n, 2 * C, h, w), float("nan"), device=dev)
    code = lib.coattn_stage_gate(z.data_ptr(), v_a.data_ptr(), v_b.data_ptr(), g.data_ptr(), None if b is None else b.data_ptr(),
                                 cat_a.data_ptr(), cat_b.data_ptr(), n, C, h, w, torch.cuda.current_stream().cuda_stream)
    _lib.check(code, "coattn_stage_gate")
    torch.cuda.synchronize()
    ref_a, ref_b = _reference(z, v_a, v_b, g, b)
    for got, ref, v in ((cat_a, ref_a, v_a), (cat_b, ref_b, v_b)):
        got = got.flatten(2)
        assert torch.equal(got[:, C:], v.flatten(2))                       # passthrough half: a copy
        assert torch.isfinite(got).all()
        err = (got[:, :C] - ref[:, :C]).abs().max().item()
        assert err <= 2e-6 * max(1.0, ref[:, :C].abs().max().item()), err


@pytest.mark.parametrize("n,h,w,bias", [
    (1, 2, 2, False),        # one ragged tile per side
    (1, 4, 8, True),
    (2, 12, 11, True),       # L = 132: two full tiles + a tail of 4 positions
    (2, 7, 6, False),        # L = 42: even, not a multiple of 4
    (3, 20, 20, False),      # L = 400
    (2, 60, 60, True),       # headline L = 3600: 56 full tiles + a tail of 16
    (5, 61, 108, False),     # L = 6588 (multiple of 4)
])
def test_gate_stage_vectorised(n, h, w, bias):
    _run(n, h, w, bias)


@pytest.mark.parametrize("n,h,w,bias,offset", [
    (2, 61, 81, True, 0),    # L = 4941, odd: scalar kernel
    (1, 61, 107, False, 0),  # L = 6527
    (2, 12, 12, True, 1),    # L even but the views are only 4-byte aligned
])
def test_gate_stage_scalar_fallback(n, h, w, bias, offset):
    _run(n, h, w, bias, offset)


def test_gate_stage_batch32():
    _run(32, 60, 60, True)   # the benchmark's shape
