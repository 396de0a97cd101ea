"""Host-side operator for the co-attention block (torch is only used for memory and streams).

`coattention(v_a, v_b, weight, gate_weight, gate_bias)` replaces the inline block of the reference
forward (rgbd_segmentation_RAA.py:150-187 for RGB, :204-238 for depth) and returns the two concat
tensors that feed `reduce_channels_A/B` (:188-189) / `depth_reduce_channels` (:239-241).
"""
from __future__ import annotations

import os
import threading

import torch

from . import _lib

_ws_lock = threading.Lock()
_ws_cache = {}  # (device index, stream handle, host thread) -> _Workspace, grown on demand

# fp16 operand range guard (include/coattn_b200.h, "Status block").  "lazy" (default): after every forward call the status
# words are copied to pinned host memory on the call's stream, and the NEXT call on the same workspace (or
# `check_overflow()`) raises if the previous one clipped -- one call late, never silent, no host synchronisation.
# "sync": every call waits for its own status (debugging).  "off": nothing is copied; `check_overflow()` still works.
OVERFLOW_CHECK = os.environ.get("COSNET_OVERFLOW_CHECK", "lazy")


class _Workspace:
    """One scratch buffer + its status mirror.  Keyed per (device, stream, host thread): two host threads never share
    one (ctypes releases the GIL, so calls of different threads interleave their launches)."""

    def __init__(self, device, nbytes):
        self.buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
        self.ptr = (self.buf.data_ptr() + 1023) // 1024 * 1024
        off = self.ptr - self.buf.data_ptr()
        self.status = self.buf[off:off + 4 * _lib.STATUS_WORDS].view(torch.int32)
        self.status.zero_()
        # no pinned host mirror for a buffer created inside a CUDA-graph capture (host allocations are not capturable; the
        # status of graphed calls is read with check_overflow-style synchronous reads by whoever owns the graph)
        self.host = None if torch.cuda.is_current_stream_capturing() else torch.zeros(_lib.STATUS_WORDS, dtype=torch.int32).pin_memory()
        self.event = None
        self._event = None
        self.stream_obj = None      # the Stream this entry is keyed on (eager cache only), for post_call without a lookup

    def numel(self):
        return self.buf.numel()

    def post_call(self, stream=None, capturing=None):
        """Queue the status read-back of the call just issued on `stream` (lazy / sync modes).  stream=None: the stream this
        entry is keyed on, which is the current one; capturing: the caller's answer to is_current_stream_capturing()."""
        if OVERFLOW_CHECK == "off" or self.host is None:
            return
        if torch.cuda.is_current_stream_capturing() if capturing is None else capturing:
            return
        if stream is None:
            stream = self.stream_obj
            if stream is None:
                stream = self.stream_obj = torch.cuda.current_stream(self.buf.device)
        self.host.copy_(self.status, non_blocking=True)
        if self._event is None:
            self._event = torch.cuda.Event()
        self.event = self._event
        self.event.record(stream)
        if OVERFLOW_CHECK == "sync":
            self.raise_if_clipped(wait=True)

    def raise_if_clipped(self, wait=False):
        if self.event is None:
            return
        if wait:
            self.event.synchronize()
        elif not self.event.query():
            return
        self.event = None
        flags = int(self.host[0])
        if flags & 7:
            self.status.zero_()
            self.host.zero_()
            what = [n for b, n in ((_lib.STATUS_OVERFLOW_B, "V_b"), (_lib.STATUS_OVERFLOW_A, "V_a"),
                                   (_lib.STATUS_OVERFLOW_Q, "Q = W V_a")) if flags & b]
            raise _lib.CoattnError(
                f"co-attention: {', '.join(what)} exceeded the fp16 operand range (+-65504) or was not finite in an earlier "
                "call on this stream; its result was clipped.  Use bf16_operands=True (fp32 exponent range) or rescale "
                "the features.")


def workspace_bytes(n: int, c: int, h: int, w: int) -> int:
    r = _lib.load().coattn_workspace_bytes(n, c, h, w)
    if r < 0:
        _lib.check(int(r), "coattn_workspace_bytes")
    return int(r)


_capture_local = threading.local()


class capture_scope:
    """Context for capturing calls of this module into a CUDA graph: workspaces that must be SHARED between calls (the
    tagged ones: `encoder_tail` writes the operand planes, `coattention_planes_ready` reads them) are kept in a cache that
    lives exactly as long as the scope's owner keeps it -- never in the eager cache, whose buffers must not come from a
    graph's private pool."""

    def __enter__(self):
        self.cache = {}
        _capture_local.cache = self.cache
        return self.cache

    def __exit__(self, *exc):
        _capture_local.cache = None


def _workspace_entry(device: torch.device, nbytes: int, tag=None, cur=None, raw=None) -> _Workspace:
    """raw: the current stream's handle when the caller has it already (saves building a Stream object per call)."""
    if torch.cuda.is_current_stream_capturing():
        # a buffer allocated during capture lives in the graph's private pool: never cache it for eager use
        scope = getattr(_capture_local, "cache", None)
        if scope is None or tag is None:
            return _Workspace(device, nbytes)
        ent = scope.get(tag)
        if ent is None or ent.numel() < nbytes + 1024:
            ent = scope[tag] = _Workspace(device, nbytes)
        return ent
    if raw is None:
        raw = (cur if cur is not None else torch.cuda.current_stream(device)).cuda_stream
    key = (device.index if device.index is not None else torch.cuda.current_device(), raw, threading.get_ident(), tag)
    with _ws_lock:
        ent = _ws_cache.get(key)
    if ent is not None:
        ent.raise_if_clipped()
    if ent is None or ent.numel() < nbytes + 1024:
        ent = _Workspace(device, nbytes)
        ent.stream_obj = cur
        with _ws_lock:
            _ws_cache[key] = ent
    return ent


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    return _workspace_entry(device, nbytes).buf


def check_overflow(device=None) -> dict:
    """Synchronous fp16-range check of every cached workspace of `device` (all devices if None): raises CoattnError if any
    call since the last check clipped an operand, otherwise returns the largest |feature| the cast kernels have seen."""
    with _ws_lock:
        ents = [(k, e) for k, e in _ws_cache.items() if device is None or k[0] == torch.device(device).index]
    amax = 0.0
    for _, e in ents:
        torch.cuda.synchronize(e.buf.device)
        words = e.status.cpu()
        e.host.copy_(words)
        e.event = torch.cuda.Event()
        e.event.record(torch.cuda.current_stream(e.buf.device))
        amax = max(amax, float(words[1:3].view(torch.float32).max()))
        e.raise_if_clipped(wait=True)
    return {"absmax": amax}


def release_workspaces(device=None):
    """Drop the cached scratch buffers (they live outside the caching allocator's reuse until released)."""
    with _ws_lock:
        for k in [k for k in _ws_cache if device is None or k[0] == torch.device(device).index]:
            del _ws_cache[k]


def _aligned_ptr(buf: torch.Tensor) -> int:
    p = buf.data_ptr()
    return (p + 1023) // 1024 * 1024


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def _on_device(dev: torch.device):
    """Device guard that costs nothing when `dev` is current already (the usual case; `torch.cuda.device` is ~4 us of
    host time per modality call, which matters at batch 1 where a call is ~55 us of GPU time)."""
    return _NO_GUARD if torch._C._cuda_getDevice() == dev.index else torch.cuda.device(dev)


def _f32(t, dev):
    """fp32 contiguous view of a parameter on `dev` -- the tensor itself when it already is one (the usual case)."""
    t = t.detach()
    if t.dtype is torch.float32 and t.device == dev and t.is_contiguous():
        return t
    return t.to(device=dev, dtype=torch.float32).contiguous()


_ws_bytes_cache = {}


def _workspace_bytes_cached(n, c, h, w):
    key = (n, c, h, w)
    r = _ws_bytes_cache.get(key)
    if r is None:
        r = _ws_bytes_cache[key] = workspace_bytes(n, c, h, w)
    return r


def _check_inputs(v_a, v_b, weight, gate_weight, gate_bias):
    if not (v_a.is_cuda and v_b.is_cuda):
        raise _lib.CoattnError("co-attention runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
    if v_a.shape != v_b.shape or v_a.dim() != 4:
        raise ValueError(f"expected two [N, C, H, W] feature maps of equal shape, got {tuple(v_a.shape)} / {tuple(v_b.shape)}")
    if v_a.dtype != torch.float32 or v_b.dtype != torch.float32:
        raise TypeError("features must be float32 (the reference encoders emit fp32, deeplabv3_encoder.py:80-82)")
    n, c, h, w = v_a.shape
    if weight.shape != (c, c):
        raise ValueError(f"weight must be [{c}, {c}], got {tuple(weight.shape)}")
    if gate_weight.numel() != c:
        raise ValueError(f"gate weight must have {c} elements, got {tuple(gate_weight.shape)}")
    if gate_bias is not None and gate_bias.numel() != 1:
        raise ValueError("gate bias must have one element")
    return n, c, h, w


def coattention_forward_raw(v_a, v_b, weight, gate_weight, gate_bias=None, bf16_operands=False,
                            unfused_gate=False, want_mask=False, want_z=True, single_cta=False, a_only=False, unfused_prep=False, gated_only=False, kmajor=False,
                            softmax16=False, split_keys=False, unfolded=False, want_lse=True):
    """Runs the CUDA stages.  Returns (cat_a, cat_b, z, lse) with z [2,N,C,L] and lse [2,N,L]
    (plus mask [2,N,L] when want_mask=True; fused path only); z / lse are None with want_z / want_lse False.

    bf16_operands=False (default): fp16 tensor-core operands with fp32 accumulation (COATTN_FLAG_BF16 unset);
    True: bf16 operands (see include/coattn_b200.h for the trade-off).
    unfolded=True: COATTN_FLAG_UNFOLDED, Q = W V_a by the stand-alone projection kernel instead of inside the attend kernel.
    split_keys=True: COATTN_FLAG_SPLIT_KEYS, the latency mode for one or two pairs (key range of every item swept in parts
    by different CTA pairs + a merge kernel; equal to the default path to fp32 rounding, not bit for bit).
    """
    n, c, h, w = _check_inputs(v_a, v_b, weight, gate_weight, gate_bias)
    lib = _lib.load()
    dev = v_a.device
    with _on_device(dev):
        v_a = v_a.contiguous()
        v_b = v_b.contiguous()
        wt = _f32(weight, dev)
        gw = _f32(gate_weight, dev).view(-1)
        gb = None if gate_bias is None else _f32(gate_bias, dev).view(-1)
        oc = c if gated_only else 2 * c
        cat_a = torch.empty((n, oc, h, w), dtype=torch.float32, device=dev)
        cat_b = None if a_only else torch.empty((n, oc, h, w), dtype=torch.float32, device=dev)
        z = torch.empty((2, n, c, h * w), dtype=torch.float32, device=dev) if (want_z or unfused_gate) else None
        # want_lse=False (inference): the library keeps the row statistics in its workspace segment instead
        lse = torch.empty((2, n, h * w), dtype=torch.float32, device=dev) if (want_lse or unfused_gate or split_keys) else None
        nbytes = _workspace_bytes_cached(n, c, h, w)
        stream = torch._C._cuda_getCurrentRawStream(dev.index)      # the handle only: no Stream object per call
        capturing = torch.cuda.is_current_stream_capturing()
        ws = _workspace_entry(dev, nbytes, raw=stream) if not capturing else _workspace_entry(dev, nbytes)
        mask = torch.empty((2, n, h * w), dtype=torch.float32, device=dev) if want_mask else None
        flags = 0
        if bf16_operands or unfused_gate or single_cta or a_only or unfused_prep or gated_only or kmajor or softmax16 or split_keys or unfolded:
            flags = ((_lib.FLAG_BF16 if bf16_operands else 0) | (_lib.FLAG_UNFUSED_GATE if unfused_gate else 0)
                     | (_lib.FLAG_SINGLE_CTA if single_cta else 0) | (_lib.FLAG_A_ONLY if a_only else 0)
                     | (_lib.FLAG_UNFUSED_PREP if unfused_prep else 0) | (_lib.FLAG_GATED_ONLY if gated_only else 0) | (_lib.FLAG_KMAJOR if kmajor else 0) | (_lib.FLAG_SOFTMAX16 if softmax16 else 0)
                     | (_lib.FLAG_SPLIT_KEYS if split_keys else 0) | (_lib.FLAG_UNFOLDED if unfolded else 0))
        code = lib.coattn_forward(v_a.data_ptr(), v_b.data_ptr(), wt.data_ptr(), gw.data_ptr(),
                                  None if gb is None else gb.data_ptr(), cat_a.data_ptr(),
                                  None if cat_b is None else cat_b.data_ptr(),
                                  None if z is None else z.data_ptr(), None if lse is None else lse.data_ptr(),
                                  None if mask is None else mask.data_ptr(),
                                  ws.ptr, nbytes, n, c, h, w, flags, stream)
        if code:
            _lib.check(code, "coattn_forward")
        if not bf16_operands:
            ws.post_call(capturing=capturing)
    if want_mask:
        return cat_a, cat_b, z, lse, mask
    return cat_a, cat_b, z, lse


def backward_workspace_bytes(n: int, c: int, h: int, w: int, counterpart: bool = False) -> int:
    r = _lib.load().coattn_backward_workspace_bytes(n, c, h, w, 1 if counterpart else 0)
    if r < 0:
        _lib.check(int(r), "coattn_backward_workspace_bytes")
    return int(r)


def coattention_queries_raw(v_a, v_b, weight, gate_weight, gate_bias=None, refs: int = 1, bf16_operands=False,
                            gated_only=False, split_keys=False):
    """test.py-style inference (test.py:287-305): v_a [Q, C, H, W] query features, v_b [Q * refs, C, H, W] reference
    features (pair p = query p // refs); returns cat_a [Q * refs, 2C (C with gated_only), H, W].  The query side (16-bit
    cast and Q = W V_a) is prepared once per query frame; the result equals the A_ONLY forward on repeated queries."""
    n, c, h, w = _check_inputs(v_a, v_a, weight, gate_weight, gate_bias)      # n = number of query frames
    if v_b.shape[0] != n * refs or tuple(v_b.shape[1:]) != (c, h, w) or not v_b.is_cuda or v_b.dtype != torch.float32:
        raise _lib.CoattnError(f"v_b must be a CUDA fp32 tensor [{n * refs}, {c}, {h}, {w}], got {tuple(v_b.shape)}")
    lib = _lib.load()
    dev = v_a.device
    with torch.cuda.device(dev):
        v_a = v_a.contiguous(); v_b = v_b.contiguous()
        weight = weight.contiguous(); gw = gate_weight.reshape(-1).contiguous()
        pairs = n * refs
        cat_a = torch.empty((pairs, (c if gated_only else 2 * c), h, w), dtype=torch.float32, device=dev)
        nbytes = workspace_bytes(pairs, c, h, w)
        ws = _workspace_entry(dev, nbytes)
        flags = ((_lib.FLAG_BF16 if bf16_operands else 0) | (_lib.FLAG_GATED_ONLY if gated_only else 0)
                 | (_lib.FLAG_SPLIT_KEYS if split_keys else 0))
        code = lib.coattn_forward_queries(v_a.data_ptr(), v_b.data_ptr(), weight.data_ptr(), gw.data_ptr(),
                                          None if gate_bias is None else gate_bias.data_ptr(), cat_a.data_ptr(),
                                          ws.ptr, nbytes, n, refs, c, h, w, flags,
                                          torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(code, "coattn_forward_queries")
        if not bf16_operands:
            ws.post_call(torch.cuda.current_stream(dev))
    return cat_a


def coattention_forward16_raw(v_a, v_b, weight, gate_weight, gate_bias=None, refs: int = 1, a_only=False,
                              gated_only=False, want_lse=False):
    """16-bit feature interface (`coattn_forward16`): v_a [Q, C, H, W] and v_b [Q * refs, C, H, W] are both float16
    (IEEE half operands) or both bfloat16 (bf16 operands); returns (cat_a, cat_b) in the same dtype (cat_b None with
    a_only / refs > 1), plus lse [passes, n, L] and mask [passes, n, L] in fp32 with want_lse.  With H*W % 8 == 0 the
    tensor cores read the features in place (no cast pass); the module's parameters stay fp32.  Forward only."""
    if not (v_a.is_cuda and v_b.is_cuda):
        raise _lib.CoattnError("co-attention runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
    if v_a.dtype not in (torch.float16, torch.bfloat16) or v_b.dtype != v_a.dtype:
        raise TypeError(f"features must both be float16 or both bfloat16, got {v_a.dtype} / {v_b.dtype}")
    if v_a.dim() != 4 or v_b.dim() != 4 or refs < 1 or v_b.shape[0] != v_a.shape[0] * refs or v_b.shape[1:] != v_a.shape[1:]:
        raise ValueError(f"expected v_a [Q, C, H, W] and v_b [Q * {refs}, C, H, W], got {tuple(v_a.shape)} / {tuple(v_b.shape)}")
    nq, c, h, w = v_a.shape
    if weight.shape != (c, c):
        raise ValueError(f"weight must be [{c}, {c}], got {tuple(weight.shape)}")
    if gate_weight.numel() != c:
        raise ValueError(f"gate weight must have {c} elements, got {tuple(gate_weight.shape)}")
    if gate_bias is not None and gate_bias.numel() != 1:
        raise ValueError("gate bias must have one element")
    a_only = a_only or refs > 1
    lib = _lib.load()
    dev = v_a.device
    with _on_device(dev):
        v_a = v_a.contiguous(); v_b = v_b.contiguous()
        wt = _f32(weight, dev)
        gw = _f32(gate_weight, dev).view(-1)
        gb = None if gate_bias is None else _f32(gate_bias, dev).view(-1)
        n = nq * refs
        oc = c if gated_only else 2 * c
        cat_a = torch.empty((n, oc, h, w), dtype=v_a.dtype, device=dev)
        cat_b = None if a_only else torch.empty((n, oc, h, w), dtype=v_a.dtype, device=dev)
        passes = 1 if a_only else 2
        lse = torch.empty((passes, n, h * w), dtype=torch.float32, device=dev) if want_lse else None
        mask = torch.empty((passes, n, h * w), dtype=torch.float32, device=dev) if want_lse else None
        nbytes = _workspace_bytes_cached(n, c, h, w)
        stream = torch._C._cuda_getCurrentRawStream(dev.index)
        capturing = torch.cuda.is_current_stream_capturing()
        ws = _workspace_entry(dev, nbytes, raw=stream) if not capturing else _workspace_entry(dev, nbytes)
        flags = ((_lib.FLAG_BF16 if v_a.dtype == torch.bfloat16 else 0) | (_lib.FLAG_A_ONLY if a_only else 0)
                 | (_lib.FLAG_GATED_ONLY if gated_only else 0))
        code = lib.coattn_forward16(v_a.data_ptr(), v_b.data_ptr(), wt.data_ptr(), gw.data_ptr(),
                                    None if gb is None else gb.data_ptr(), cat_a.data_ptr(),
                                    None if cat_b is None else cat_b.data_ptr(),
                                    None if lse is None else lse.data_ptr(), None if mask is None else mask.data_ptr(),
                                    ws.ptr, nbytes, nq, refs, c, h, w, flags, stream)
        if code:
            _lib.check(code, "coattn_forward16")
        if v_a.dtype == torch.float16:
            ws.post_call(capturing=capturing)
    if want_lse:
        return cat_a, cat_b, lse, mask
    return cat_a, cat_b


def bn_eval_affine(bn: torch.nn.BatchNorm2d):
    """Eval-mode BatchNorm as y = x * scale + shift (running statistics)."""
    scale = bn.weight.detach() * torch.rsqrt(bn.running_var + bn.eps)
    shift = bn.bias.detach() - bn.running_mean * scale
    return scale.float().contiguous(), shift.float().contiguous()


def encoder_tail(x, scale, shift, slope, frame: int, tag, bf16_operands=False, want_features=True):
    """Producer side of the hot path (deeplab/deeplabv3_encoder.py:80-82), inference only:
    features = PReLU(BN_eval(x)) for the bottleneck-conv output x [N, 256, H, W], fused with the co-attention's operand
    cast: ONE kernel writes the fp32 features (returned; None with want_features=False) and the 16-bit operand plane of
    frame A (frame=0) or B (frame=1) into the workspace tagged `tag`; `coattention_planes_ready(..., tag=tag)` then starts
    at the projection.  Both frames of a pair batch must go through this before that call, on the same stream."""
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 4:
        raise _lib.CoattnError("encoder_tail needs a CUDA fp32 [N, 256, H, W] tensor; there is no CPU fallback")
    n, c, h, w = x.shape
    lib = _lib.load()
    dev = x.device
    with torch.cuda.device(dev):
        x = x.contiguous()
        y = torch.empty_like(x) if want_features else None
        nbytes = workspace_bytes(n, c, h, w)
        ws = _workspace_entry(dev, nbytes, tag=tag)
        code = lib.coattn_stage_tail(x.data_ptr(), scale.data_ptr(), shift.data_ptr(), slope.detach().float().contiguous().data_ptr(),
                                     None if y is None else y.data_ptr(), ws.ptr, nbytes, frame, n, c, h, w,
                                     _lib.FLAG_BF16 if bf16_operands else 0, torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(code, "coattn_stage_tail")
    return y


def coattention_planes_ready(v_a, v_b, weight, gate_weight, gate_bias, tag, bf16_operands=False, gated_only=False):
    """`coattention` for features whose 16-bit operand planes were already written by `encoder_tail(..., tag=tag)`:
    no cast kernel.  v_a / v_b (fp32) are only read for the passthrough half of the concat (not at all with gated_only)."""
    n, c, h, w = _check_inputs(v_a, v_b, weight, gate_weight, gate_bias)
    lib = _lib.load()
    dev = v_a.device
    with torch.cuda.device(dev):
        wt = weight.detach().float().contiguous()
        gw = gate_weight.detach().float().contiguous().view(-1)
        gb = None if gate_bias is None else gate_bias.detach().float().contiguous().view(-1)
        oc = c if gated_only else 2 * c
        cat_a = torch.empty((n, oc, h, w), dtype=torch.float32, device=dev)
        cat_b = torch.empty((n, oc, h, w), dtype=torch.float32, device=dev)
        nbytes = workspace_bytes(n, c, h, w)
        ws = _workspace_entry(dev, nbytes, tag=tag)
        cur = torch.cuda.current_stream(dev)
        flags = _lib.FLAG_PLANES_READY | (_lib.FLAG_BF16 if bf16_operands else 0) | (_lib.FLAG_GATED_ONLY if gated_only else 0)
        code = lib.coattn_forward(v_a.data_ptr(), v_b.data_ptr(), wt.data_ptr(), gw.data_ptr(),
                                  None if gb is None else gb.data_ptr(), cat_a.data_ptr(), cat_b.data_ptr(), None, None, None,
                                  ws.ptr, nbytes, n, c, h, w, flags, cur.cuda_stream)
        _lib.check(code, "coattn_forward")
        if not bf16_operands:
            ws.post_call(cur)
    return cat_a, cat_b


class _CoAttentionFn(torch.autograd.Function):
    """Autograd bridge: forward = coattn_forward (keeps z, lse, mask), backward = coattn_backward.

    Gradient semantics of the reference: the B-side gate mask is a constant (rgbd_segmentation_RAA.py:178-182), so the
    gate parameters only receive gradient through cat_a; the gradient for the counterpart features v_b is only
    computed when v_b requires grad (no_grad_for_counterpart=False, :147-148).
    """

    @staticmethod
    def forward(ctx, v_a, v_b, weight, gate_weight, gate_bias, bf16_operands, gated_only=False):
        # the backward hands raw pointers of the saved features to the library: save the contiguous NCHW tensors the
        # forward kernels actually read (channels_last / strided encoder outputs are copied here, once)
        v_a = v_a.contiguous()
        v_b = v_b.contiguous()
        cat_a, cat_b, z, lse, mask = coattention_forward_raw(v_a, v_b, weight, gate_weight, gate_bias, bf16_operands,
                                                             want_mask=True, want_z=True, gated_only=gated_only)
        ctx.gated_only = gated_only
        ctx.save_for_backward(v_a, v_b, weight, gate_weight, z, lse, mask)
        ctx.has_bias = gate_bias is not None
        ctx.bf16 = bool(bf16_operands)
        ctx.v_b_needs_grad = v_b.requires_grad
        ctx.set_materialize_grads(False)
        return cat_a, cat_b

    @staticmethod
    def backward(ctx, d_cat_a, d_cat_b):
        v_a, v_b, weight, gate_weight, z, lse, mask = ctx.saved_tensors
        n, c, h, w = v_a.shape
        dev = v_a.device
        lib = _lib.load()
        with torch.cuda.device(dev):
            if d_cat_a is None:
                d_cat_a = torch.zeros((n, (c if ctx.gated_only else 2 * c), h, w), dtype=torch.float32, device=dev)
            d_cat_a = d_cat_a.contiguous().float()
            d_cat_b = None if d_cat_b is None else d_cat_b.contiguous().float()
            wt = weight.detach().float().contiguous()
            gw = gate_weight.detach().float().contiguous().view(-1)
            d_v_a = torch.empty(v_a.shape, dtype=torch.float32, device=dev)
            d_w = torch.empty((c, c), dtype=torch.float32, device=dev)
            d_gw = torch.empty((c,), dtype=torch.float32, device=dev)
            d_gb = torch.empty((1,), dtype=torch.float32, device=dev) if ctx.has_bias else None
            d_v_b = torch.empty(v_b.shape, dtype=torch.float32, device=dev) if ctx.v_b_needs_grad else None
            nbytes = backward_workspace_bytes(n, c, h, w, ctx.v_b_needs_grad)
            # per call, through the caching allocator (stream-ordered reuse is safe; nothing stays pinned between steps)
            ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            code = lib.coattn_backward(v_a.data_ptr(), v_b.data_ptr(), wt.data_ptr(), gw.data_ptr(), z.data_ptr(),
                                       lse.data_ptr(), mask.data_ptr(), d_cat_a.data_ptr(),
                                       None if d_cat_b is None else d_cat_b.data_ptr(), d_v_a.data_ptr(),
                                       None if d_v_b is None else d_v_b.data_ptr(), d_w.data_ptr(), d_gw.data_ptr(), None if d_gb is None else d_gb.data_ptr(),
                                       _aligned_ptr(ws), nbytes, n, c, h, w,
                                       (_lib.FLAG_BF16 if ctx.bf16 else 0) | (_lib.FLAG_GATED_ONLY if ctx.gated_only else 0), stream)
            _lib.check(code, "coattn_backward")
        return d_v_a, d_v_b, d_w.to(weight.dtype), d_gw.view_as(gate_weight).to(gate_weight.dtype), d_gb, None, None


def coattention(v_a, v_b, weight, gate_weight, gate_bias=None, bf16_operands=False, gated_only=False):
    """Drop-in for rgbd_segmentation_RAA.py:150-187: returns (cat_a, cat_b), each [N, 2C, H, W].  Differentiable
    w.r.t. v_a, v_b, weight, gate_weight and gate_bias (hand-written CUDA backward).  float16 / bfloat16 features give
    outputs of the same dtype: `coattention_forward16_raw` without gradients, the fp32 path on the widened features with.

    gated_only=True returns only the gated attended features Z * sigmoid(gate(Z)), [N, C, H, W] each (the first half of
    the concat), for consumers that apply the reduce conv in two halves and never need the concat itself."""
    needs_grad = torch.is_grad_enabled() and any(
        t is not None and t.requires_grad for t in (v_a, v_b, weight, gate_weight, gate_bias))
    if v_a.dtype in (torch.float16, torch.bfloat16):
        # 16-bit features (a half-precision or autocast encoder): outputs in the same dtype, the operand format follows
        # the feature dtype.  Without gradients: coattn_forward16 (features read in place).  With gradients (mixed-precision
        # training): the hand-written backward takes fp32 features, so they are widened first -- the same 16-bit values
        # reach the tensor cores, the concat is rounded once at the end and autograd narrows d_v_a / d_v_b.
        if needs_grad:
            cat_a, cat_b = _CoAttentionFn.apply(v_a.float(), v_b.float(), weight, gate_weight, gate_bias,
                                                v_a.dtype == torch.bfloat16, gated_only)
            return cat_a.to(v_a.dtype), cat_b.to(v_a.dtype)
        return coattention_forward16_raw(v_a, v_b, weight, gate_weight, gate_bias, gated_only=gated_only)
    if needs_grad:
        return _CoAttentionFn.apply(v_a, v_b, weight, gate_weight, gate_bias, bf16_operands, gated_only)
    cat_a, cat_b, _, _ = coattention_forward_raw(v_a, v_b, weight, gate_weight, gate_bias, bf16_operands, want_z=False,
                                                 gated_only=gated_only, want_lse=False)
    return cat_a, cat_b


_side_streams = {}


def _side_stream(device: torch.device) -> torch.cuda.Stream:
    key = device.index if device.index is not None else torch.cuda.current_device()
    with _ws_lock:
        st = _side_streams.get(key)
        if st is None:
            st = _side_streams[key] = torch.cuda.Stream(device)
    return st


def modality_overlap_pays(n: int, h: int, w: int, passes: int = 2, device=None, clusters: int = 0) -> bool:
    """True when running the RGB and the depth modality call concurrently needs fewer waves of the attend kernel than
    running them back to back.  A call is `passes * n * ceil(L / 256)` equal work items for SMs / 2 CTA pairs; the last
    wave of a call is partly empty, and a second call on another stream back-fills it.  Measured:
    one 60x60 pair 121 -> 65 us, two 61x107 pairs +27 %, 16 of them +3 %; no gain (or a few % loss) when the combined
    item count needs as many waves as the two calls separately, e.g. 32 pairs at 60x60."""
    if clusters <= 0:      # CTA pairs of the device (74 on a B200); pass `clusters` to evaluate the rule without a GPU
        clusters = torch.cuda.get_device_properties(device if device is not None else torch.cuda.current_device()).multi_processor_count // 2
    items = passes * n * ((h * w + 255) // 256)
    waves = lambda x: (x + clusters - 1) // clusters
    return 2 * waves(items) > waves(2 * items)


def run_modalities(rgb_call, depth_call, depth_inputs, overlap: bool):
    """(rgb_call(), depth_call()) -- with overlap=True the depth call is issued on a side stream, so the kernels of the two
    modality calls (independent: own weights, own workspace per stream) share the GPU.  No autograd: inference paths only.
    depth_inputs: the tensors depth_call reads (kept alive for the side stream)."""
    if not overlap:
        return rgb_call(), depth_call()
    dev = depth_inputs[0].device
    cur = torch.cuda.current_stream(dev)
    side = _side_stream(dev)
    side.wait_stream(cur)
    torch.cuda.set_stream(side)      # same device: cheaper than the `torch.cuda.stream` context (host time counts at batch 1)
    try:
        d_out = depth_call()
    finally:
        torch.cuda.set_stream(cur)
    r_out = rgb_call()
    cur.wait_stream(side)
    for t in depth_inputs:
        t.record_stream(side)
    for t in (d_out if isinstance(d_out, (tuple, list)) else (d_out,)):
        if isinstance(t, torch.Tensor):
            t.record_stream(cur)
    return r_out, d_out


_pair_events = {}


def coattention_pair(rgb, depth, gated_only=False, bf16_operands=False):
    """Both modality calls of a batch of frame pairs (rgbd_segmentation_RAA.py:150-187 and :204-238), inference only:
    `rgb` and `depth` are `(v_a, v_b, weight, gate_weight, gate_bias)`; returns `((cat_a, cat_b), (dcat_a, dcat_b))`.

    The RGB call goes to the current stream and the depth call to a side stream BY HANDLE -- torch's current stream is never
    switched, the two streams are forked and joined with two cached events -- so that the kernels of the two calls share the
    GPU (what `run_modalities(..., overlap=True)` does) at about half its host cost: at batch 1 an eager pair is bound by the
    host, not by the 62 us of kernels.  Results are bit-identical to two `coattention()` calls.  All
    tensors are allocated on the current stream and every side-stream use is joined before returning, so the caching
    allocator needs no `record_stream`."""
    n, c, h, w = _check_inputs(*rgb)
    if _check_inputs(*depth) != (n, c, h, w):
        raise ValueError("the RGB and the depth features must have the same shape")
    lib = _lib.load()
    dev = rgb[0].device
    if depth[0].device != dev:
        raise ValueError("the RGB and the depth features must live on the same device")
    if torch.cuda.is_current_stream_capturing():      # inside a capture the generic path does the right thing
        return (coattention(*rgb, bf16_operands=bf16_operands, gated_only=gated_only),
                coattention(*depth, bf16_operands=bf16_operands, gated_only=gated_only))
    with _on_device(dev), torch.no_grad():
        cur = torch.cuda.current_stream(dev)
        side = _side_stream(dev)
        ev = _pair_events.get((dev.index, threading.get_ident()))
        if ev is None:
            ev = _pair_events[(dev.index, threading.get_ident())] = (torch.cuda.Event(), torch.cuda.Event())
        oc = c if gated_only else 2 * c
        nbytes = _workspace_bytes_cached(n, c, h, w)
        flags = (_lib.FLAG_BF16 if bf16_operands else 0) | (_lib.FLAG_GATED_ONLY if gated_only else 0)
        # every copy / conversion of an input is queued on the current stream BEFORE the fork
        prepared = []
        for v_a, v_b, weight, gate_weight, gate_bias in (depth, rgb):
            prepared.append((v_a.contiguous(), v_b.contiguous(), _f32(weight, dev), _f32(gate_weight, dev).view(-1),
                             None if gate_bias is None else _f32(gate_bias, dev).view(-1)))
        raws = (side.cuda_stream, cur.cuda_stream)
        wss = [_workspace_entry(dev, nbytes, raw=r) for r in raws]      # (a new workspace zeroes its status block on `cur`)
        outs = []
        ev[0].record(cur)
        side.wait_event(ev[0])                           # the depth features were produced on the current stream
        for (v_a, v_b, wt, gw, gb), raw, ws in zip(prepared, raws, wss):
            cat_a = torch.empty((n, oc, h, w), dtype=torch.float32, device=dev)
            cat_b = torch.empty((n, oc, h, w), dtype=torch.float32, device=dev)
            code = lib.coattn_forward(v_a.data_ptr(), v_b.data_ptr(), wt.data_ptr(), gw.data_ptr(),
                                      None if gb is None else gb.data_ptr(), cat_a.data_ptr(), cat_b.data_ptr(), None, None,
                                      None, ws.ptr, nbytes, n, c, h, w, flags, raw)
            if code:
                _lib.check(code, "coattn_forward")
            outs.append((cat_a, cat_b))
        ev[1].record(side)
        cur.wait_event(ev[1])                            # join: everything below and after is ordered behind both calls
        if not bf16_operands:
            for ws in wss:
                ws.post_call(cur, capturing=False)      # status read-back of both workspaces on the current stream
    return outs[1], outs[0]


class HostPipeline:
    """Host-buffer entry point: features live in (pinned) host memory, results return to host memory.

    The batch is cut into chunks that rotate over `slots` CUDA streams; each stream runs
    H2D -> kernels -> D2H in order, so the copies of neighbouring chunks overlap the kernels.
    Device buffers and workspaces are allocated once and reused across calls.

    host_passthrough=False (default): the device produces the whole concat and all of it crosses PCIe back.
    True: the second half of each concat tensor (a bit-exact copy of the input features, :186-187, which the caller
    already holds in host memory) is copied host-to-host by worker threads and only the gated half is read back.
    Measured on the round-1 box this is NOT faster (1.7k vs 1.6k pairs/s): the host memcpy competes with the DMA
    traffic for host memory bandwidth, so it stays opt-in.
    """

    def __init__(self, n: int, c: int, h: int, w: int, chunk: int = 8, slots: int = 4, device="cuda:0",
                 bf16_operands: bool = False, host_passthrough: bool = False, gated_only: bool = False,
                 feature_dtype: torch.dtype = torch.float32):
        self.n, self.c, self.h, self.w = n, c, h, w
        # feature_dtype float16 / bfloat16: host and device buffers hold 16-bit features and results (coattn_forward16;
        # the operand format follows the dtype) -- half the bytes on the wire in both directions
        if feature_dtype not in (torch.float32, torch.float16, torch.bfloat16):
            raise TypeError(f"feature_dtype must be float32, float16 or bfloat16, got {feature_dtype}")
        self.dtype = feature_dtype
        self.io16 = feature_dtype != torch.float32
        if self.io16:
            bf16_operands = feature_dtype == torch.bfloat16
            host_passthrough = False
        esize = 2 if self.io16 else 4
        self.flags = _lib.FLAG_BF16 if bf16_operands else 0
        # gated_only: outputs are [n, C, h, w] (Z * sigmoid(gate) only) for a consumer that applies the reduce conv to the
        # two halves separately (split_reduce_conv) -- the passthrough half never exists, on the device or on the wire
        self.gated_only = gated_only
        if gated_only:
            self.flags |= _lib.FLAG_GATED_ONLY
            host_passthrough = False
        self.chunk = max(1, min(chunk, n))
        self.device = torch.device(device)
        self.host_passthrough = host_passthrough
        self.lib = _lib.load()
        self.slots = []
        with torch.cuda.device(self.device):
            nbytes = workspace_bytes(self.chunk, c, h, w)
            for _ in range(slots):
                s = {
                    "stream": torch.cuda.Stream(self.device),
                    "va": torch.empty((self.chunk, c, h, w), dtype=self.dtype, device=self.device),
                    "vb": torch.empty((self.chunk, c, h, w), dtype=self.dtype, device=self.device),
                    "ca": torch.empty((self.chunk, (c if gated_only else 2 * c), h, w), dtype=self.dtype, device=self.device),
                    "cb": torch.empty((self.chunk, (c if gated_only else 2 * c), h, w), dtype=self.dtype, device=self.device),
                    "ws": torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device),
                    "nbytes": nbytes,
                }
                self.slots.append(s)
        self.h2d_bytes = 2 * n * c * h * w * esize
        self.d2h_bytes = 2 * n * (c if (host_passthrough or gated_only) else 2 * c) * h * w * esize
        # fp32: cast, cast_w, attend2 (which projects its query tiles itself);  16-bit features read in place: cast_w, attend2
        in_place = self.io16 and (h * w) % 8 == 0
        self.launches_per_call = (2 if in_place else 3) * ((n + self.chunk - 1) // self.chunk)
        self._worker = None

    def _host_copy(self, v_a, v_b, out_a, out_b, lo, hi):
        c = self.c
        for i in range(lo, hi):     # contiguous 4*C*H*W-byte memcpys; torch releases the GIL inside copy_
            out_a[i, c:].copy_(v_a[i])
            out_b[i, c:].copy_(v_b[i])

    def join(self):
        """Make the current stream wait for everything the slot streams have been given (after calls with join=False)."""
        cur = torch.cuda.current_stream(self.device)
        for s in self.slots:
            cur.wait_stream(s["stream"])

    def __call__(self, v_a, v_b, weight, gate_weight, gate_bias, out_a, out_b, join: bool = True):
        """v_a, v_b, out_a, out_b: host tensors (pin them for asynchronous copies); weights on the device.

        join=False: do not make the current stream wait for the slot streams at the end.  Consecutive calls (the RGB and the
        depth modality of a step, the next step's batch) then stream through the slots back to back -- the first H2D of a
        call overlaps the last D2H of the one before instead of waiting for it (with 8 chunks per call the fill and drain of
        the pipeline are ~20 % of a call).  Call `join()` (or synchronise the device) before reading the outputs."""
        import threading
        n, c, h, w = self.n, self.c, self.h, self.w
        gw = gate_weight.view(-1)
        if self.host_passthrough:
            nthr = min(8, n)
            per = (n + nthr - 1) // nthr
            self._worker = [threading.Thread(target=self._host_copy, args=(v_a, v_b, out_a, out_b, t * per, min(n, (t + 1) * per)))
                            for t in range(nthr) if t * per < n]
            for th in self._worker:
                th.start()
        cur = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        for k, lo in enumerate(range(0, n, self.chunk)):
            hi = min(n, lo + self.chunk)
            m = hi - lo
            s = self.slots[k % len(self.slots)]
            st = s["stream"]
            st.wait_event(ready)
            with torch.cuda.stream(st):
                s["va"][:m].copy_(v_a[lo:hi], non_blocking=True)
                s["vb"][:m].copy_(v_b[lo:hi], non_blocking=True)
                no_pass = self.host_passthrough or self.gated_only
                pass_a = None if no_pass else s["va"].data_ptr()
                pass_b = None if no_pass else s["vb"].data_ptr()
                nb = s["nbytes"]
                wsp = _aligned_ptr(s["ws"])
                if self.io16:
                    _lib.check(self.lib.coattn_forward16(
                        s["va"].data_ptr(), s["vb"].data_ptr(), weight.data_ptr(), gw.data_ptr(),
                        None if gate_bias is None else gate_bias.data_ptr(), s["ca"].data_ptr(), s["cb"].data_ptr(), None,
                        None, wsp, nb, m, 1, c, h, w, self.flags, st.cuda_stream), "coattn_forward16")
                    out_a[lo:hi].copy_(s["ca"][:m], non_blocking=True)
                    out_b[lo:hi].copy_(s["cb"][:m], non_blocking=True)
                    continue
                _lib.check(self.lib.coattn_stage_prep_project(s["va"].data_ptr(), s["vb"].data_ptr(), weight.data_ptr(), wsp,
                                                              nb, m, c, h, w, self.flags, st.cuda_stream),
                           "coattn_stage_prep_project")
                _lib.check(self.lib.coattn_stage_attend_gate(
                    pass_a, pass_b, s["ca"].data_ptr(), s["cb"].data_ptr(), None, None, None, gw.data_ptr(),
                    None if gate_bias is None else gate_bias.data_ptr(), wsp, nb, m, c, h, w, self.flags,
                    st.cuda_stream), "coattn_stage_attend_gate")
                if self.host_passthrough:
                    for i in range(m):      # per-sample contiguous D2H copies of the gated half
                        out_a[lo + i, :c].copy_(s["ca"][i, :c], non_blocking=True)
                        out_b[lo + i, :c].copy_(s["cb"][i, :c], non_blocking=True)
                else:
                    out_a[lo:hi].copy_(s["ca"][:m], non_blocking=True)
                    out_b[lo:hi].copy_(s["cb"][:m], non_blocking=True)
        if join:
            for s in self.slots:
                cur.wait_stream(s["stream"])
        return out_a, out_b

    def wait_host(self):
        """Join the host-side passthrough copy of the last call (call before reading the outputs)."""
        if self._worker is not None:
            for th in self._worker:
                th.join()
            self._worker = None
