"""Shared helpers of the test-suite (tests only)."""
import hashlib
import os

import numpy as np

from oracle import coattn_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_inputs(fx):
    """Regenerate the inputs of a golden fixture from its seed recipe and check their digest."""
    seed, n, h, w, sigma = int(fx["seed"]), int(fx["n"]), int(fx["h"]), int(fx["w"]), float(fx["sigma"])
    v_a, v_b, d_a, d_b = orc.synthetic_features(seed, n, h, w, sigma, count=4)
    w_rgb, g_rgb, _ = orc.synthetic_weights(seed + 1, bias=False)
    w_dep, g_dep, b_dep = orc.synthetic_weights(seed + 2, bias=True)
    inp = dict(v_a=v_a, v_b=v_b, d_a=d_a, d_b=d_b, w_rgb=w_rgb, g_rgb=g_rgb, w_dep=w_dep, g_dep=g_dep, b_dep=b_dep)
    hsh = hashlib.sha256()
    for k in sorted(inp):
        hsh.update(np.ascontiguousarray(inp[k]).tobytes())
    assert hsh.hexdigest() == str(fx["sha256"]), "numpy RNG stream changed: regenerate tests/golden with oracle/make_golden.py"
    return inp


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel_l2(x, ref):
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.linalg.norm(x - ref) / max(np.linalg.norm(ref), 1e-30))


def subsample(x, fx):
    """The strided subsample the full-size fixtures keep: x [N, C, ...] -> [N, C / sub_c, L / sub_p]."""
    x = np.asarray(x)
    x = x.reshape(x.shape[0], x.shape[1], -1)
    return x[:, ::int(fx["sub_c"]), ::int(fx["sub_p"])]
