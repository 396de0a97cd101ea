"""CPU timing port of the reference co-attention -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

Restates the op sequence of /root/reference/rgbd_segmentation_RAA.py:154-187 with the same ATen
operators (fp32, same materialisations: the transposed-contiguous copy :158, the S clone :164, the
softmax over a non-innermost dim :164 and over a transposed view :165), so that timing it on the GPU
box's host cores measures what the reference's own CPU path would cost there.  `bench.py` uses it for
`cpu_baseline` / `--impl reference` (kind "port") when /root/reference is not present; when it is, the
unmodified reference is timed instead (kind "reference", see oracle/ref_harness.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


@torch.no_grad()
def coattention_cpu(v_a, v_b, weight, gate_weight, gate_bias=None):
    n, c, h, w = v_a.shape
    l = h * w
    a_flat = v_a.view(n, c, l)                                             # :154
    b_flat = v_b.view(n, c, l)                                             # :155
    q = F.linear(a_flat.transpose(1, 2).contiguous(), weight)              # :158-159
    s = torch.bmm(q, b_flat)                                               # :160
    s_row = F.softmax(s.clone(), dim=1)                                    # :164
    s_col = F.softmax(s.transpose(1, 2), dim=1)                            # :165
    z_b = torch.bmm(a_flat, s_row).contiguous().view(n, c, h, w)           # :169, :176
    z_a = torch.bmm(b_flat, s_col).contiguous().view(n, c, h, w)           # :170, :175
    gw = gate_weight.view(1, c, 1, 1)
    m_a = torch.sigmoid(F.conv2d(z_a, gw, gate_bias))                      # :177, :180
    m_b = torch.sigmoid(F.conv2d(z_b, gw, gate_bias))                      # :179, :182
    cat_a = torch.cat([z_a * m_a, v_a], 1)                                 # :183, :186
    cat_b = torch.cat([z_b * m_b, v_b], 1)                                 # :184, :187
    return cat_a, cat_b
