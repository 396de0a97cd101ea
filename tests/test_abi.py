"""The C-ABI library loads and exports every symbol include/coattn_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "coattn_b200.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from cosnet_b200 import _lib
    return _lib.load()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(coattn_\w+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = declared_functions()
    assert "coattn_forward" in names and "coattn_stage_attend" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_binding_covers_header():
    from cosnet_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_functions()


def test_version_and_errors(lib):
    from cosnet_b200 import _lib
    assert lib.coattn_b200_abi_version() == _lib.ABI_VERSION
    assert lib.coattn_b200_strerror(0) == b"ok"
    assert b"256" in lib.coattn_b200_strerror(-2)
    assert b"fallback" in lib.coattn_b200_strerror(-4)


def test_workspace_size_and_shape_errors(lib):
    # 60x60 -> Lp = 3840: six 16-bit planes of N*Lp*C (Bt, Qt, At, B16, A16, Q16) plus W16, z, lse; the z / lse segments
    # hold one part per key-range split COATTN_FLAG_SPLIT_KEYS may use at this size (2 pairs: 30 frame-A items -> 2 parts;
    # 8 pairs fill the 74 CTA pairs -> 1)
    c, h, w = 256, 60, 60
    lp = 3840
    for n, parts in ((2, 2), (8, 1), (32, 1)):
        plane = n * lp * c * 2
        expect_min = 6 * plane + c * c * 2 + parts * (2 * n * c * h * w * 4 + 2 * n * h * w * 4)
        got = lib.coattn_workspace_bytes(n, c, h, w)
        assert expect_min <= got <= expect_min + 8 * 1024, (n, parts)
    n = 2
    plane = n * lp * c * 2
    assert lib.coattn_workspace_bytes(n, 128, h, w) == -2      # C must be 256
    assert lib.coattn_workspace_bytes(0, c, h, w) == -2
    off, nb = ctypes.c_int64(), ctypes.c_int64()
    assert lib.coattn_workspace_segment(b"qt", n, c, h, w, ctypes.byref(off), ctypes.byref(nb)) == 0
    assert off.value == 1024 + plane and nb.value == plane      # the status block takes the first 1024 bytes
    assert lib.coattn_workspace_segment(b"status", n, c, h, w, ctypes.byref(off), ctypes.byref(nb)) == 0
    assert off.value == 0 and nb.value == 4 * 8                  # COATTN_STATUS_WORDS
    assert lib.coattn_status_clear(None, None) == -1 and lib.coattn_status_read(None, None, None) == -1
    assert lib.coattn_workspace_segment(b"nope", n, c, h, w, ctypes.byref(off), ctypes.byref(nb)) == -1


def test_argument_errors_before_any_cuda_work(lib):
    # NULL pointers and bad shapes are rejected on the host, before touching a device
    assert lib.coattn_forward(None, None, None, None, None, None, None, None, None, None, None, 0, 1, 256, 4, 4, 0, None) == -1
    assert lib.coattn_stage_gate(None, None, None, None, None, None, None, 1, 256, 4, 4, None) == -1
    assert lib.coattn_stage_project(None, 0, 1, 64, 4, 4, 0, None) == -2
    P = 0x10000      # fake, never dereferenced pointer (1024-byte aligned): every check below fails before any device work
    # the grouped-query and 16-bit entry points: NULL, shape, flag and alignment checks come first as well
    assert lib.coattn_forward_queries(None, None, None, None, None, None, None, 0, 1, 1, 256, 4, 4, 0, None) == -1
    assert lib.coattn_forward_queries(P, P, P, P, None, P, P, 1 << 40, 0, 1, 256, 4, 4, 0, None) == -2      # nq < 1
    assert lib.coattn_forward_queries(P, P, P, P, None, P, P, 1 << 40, 1, 1, 256, 4, 4, 2, None) == -7      # UNFUSED_GATE
    f16 = lambda *a: lib.coattn_forward16(*a)
    assert f16(None, None, None, None, None, None, None, None, None, None, 0, 1, 1, 256, 4, 4, 0, None) == -1
    assert f16(P, P, P, P, None, P, P, None, None, P, 1 << 40, 1, 0, 256, 4, 4, 0, None) == -2              # refs < 1
    assert f16(P, P, P, P, None, P, P, None, None, P, 1 << 40, 1, 1, 256, 4, 4, 64, None) == -7             # KMAJOR
    assert f16(P, P, P, P, None, P, P, None, None, P, 1 << 40, 1, 1, 256, 4, 4, 256, None) == -7            # SPLIT_KEYS
    assert f16(P, P, P, P, None, P, None, None, None, P, 1 << 40, 1, 1, 256, 4, 4, 0, None) == -1           # cat_b needed
    assert f16(P, P, P, P, None, P, P, None, None, P, 1 << 40, 1, 1, 128, 4, 4, 0, None) == -2              # C != 256
    assert f16(P, P, P, P, None, P, P, None, None, P + 8, 1 << 40, 1, 1, 256, 4, 4, 0, None) == -3          # workspace alignment
    # coattn_backward: the weight gradient is reduced with 16-byte vector operations -> a misaligned d_w is refused
    # (fake, never dereferenced pointers: the check comes before any device work)
    args = [P] * 8 + [None, P, None, P + 4, P, None, P, 1 << 40, 1, 256, 4, 4, 0, None]
    assert lib.coattn_backward(*args) == -6
    assert b"aligned" in lib.coattn_b200_strerror(-6)


def test_host_operator_refuses_cpu_tensors():
    import torch
    from cosnet_b200 import coattention
    from cosnet_b200._lib import CoattnError
    x = torch.zeros(1, 256, 2, 2)
    with pytest.raises(CoattnError):
        coattention(x, x, torch.zeros(256, 256), torch.zeros(256))
