"""Stage-by-stage GPU diagnostics of the CUDA path against the oracle (run on the B200 box).

    python tests/gpu_diag.py --stage prep|project|attend|gate|forward|time [--n N --h H --w W --sigma S]

Prints error metrics instead of asserting, so one remote run tells as much as possible.
"""
import argparse
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cosnet_b200 import _lib  # noqa: E402
from cosnet_b200.coattention import coattention_forward_raw, workspace_bytes  # noqa: E402
from oracle import coattn_oracle as orc  # noqa: E402


def rel_l2(x, ref):
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.linalg.norm(x - ref) / max(np.linalg.norm(ref), 1e-30))


def segment(ws, name, n, c, h, w, dtype, shape):
    lib = _lib.load()
    off, nb = ctypes.c_int64(), ctypes.c_int64()
    _lib.check(lib.coattn_workspace_segment(name.encode(), n, c, h, w, ctypes.byref(off), ctypes.byref(nb)), "segment")
    base = (ws.data_ptr() + 1023) // 1024 * 1024 - ws.data_ptr()
    raw = ws[base + off.value: base + off.value + nb.value]
    return raw.view(dtype).view(*shape)


def bf16_round(x, dt=torch.bfloat16):
    return torch.from_numpy(np.asarray(x, dtype=np.float32)).to(dt).to(torch.float32).numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", default="forward")
    ap.add_argument("--n", type=int, default=1)
    ap.add_argument("--h", type=int, default=12)
    ap.add_argument("--w", type=int, default=11)
    ap.add_argument("--sigma", type=float, default=0.66)
    ap.add_argument("--bias", type=int, default=1)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--bf16", type=int, default=0)
    ap.add_argument("--single", type=int, default=0)
    ap.add_argument("--unfused-prep", type=int, default=0)
    ap.add_argument("--kmajor", type=int, default=0)
    args = ap.parse_args()
    n, h, w, c = args.n, args.h, args.w, 256
    FL = (1 if args.bf16 else 0) | (4 if args.single else 0) | (64 if args.kmajor else 0)
    odt = torch.bfloat16 if args.bf16 else torch.float16
    L = h * w
    Lp = (L + 255) // 256 * 256
    lib = _lib.load()
    dev = torch.device("cuda:0")
    print(f"[diag] stage={args.stage} n={n} h={h} w={w} L={L} Lp={Lp} sigma={args.sigma} dev={torch.cuda.get_device_name(0)}", flush=True)
    v_a, v_b = orc.synthetic_features(11, n, h, w, args.sigma)
    W, g, b = orc.synthetic_weights(12, bias=bool(args.bias))
    tva, tvb = torch.from_numpy(v_a).to(dev), torch.from_numpy(v_b).to(dev)
    tW, tg = torch.from_numpy(W).to(dev), torch.from_numpy(g).to(dev)
    tb = None if b is None else torch.from_numpy(b).to(dev)
    nbytes = workspace_bytes(n, c, h, w)
    ws = torch.zeros(nbytes + 1024, dtype=torch.uint8, device=dev)
    wsp = (ws.data_ptr() + 1023) // 1024 * 1024
    st = torch.cuda.current_stream().cuda_stream

    if args.stage in ("prep", "project", "attend"):
        _lib.check(lib.coattn_stage_prep(tva.data_ptr(), tvb.data_ptr(), tW.data_ptr(), wsp, nbytes, n, c, h, w, FL, st), "prep")
        torch.cuda.synchronize()
        at = segment(ws, "at", n, c, h, w, odt, (n, Lp, c)).float().cpu().numpy()
        bt = segment(ws, "bt", n, c, h, w, odt, (n, Lp, c)).float().cpu().numpy()
        a16 = segment(ws, "a16", n, c, h, w, odt, (n, c, Lp)).float().cpu().numpy()
        b16 = segment(ws, "b16", n, c, h, w, odt, (n, c, Lp)).float().cpu().numpy()
        w16 = segment(ws, "w16", n, c, h, w, odt, (c, c)).float().cpu().numpy()
        ra = bf16_round(v_a.reshape(n, c, L), odt); rb = bf16_round(v_b.reshape(n, c, L), odt)
        print("[prep] a16 exact:", np.array_equal(a16[:, :, :L], ra), " pad zero:", float(np.abs(a16[:, :, L:]).max(initial=0)))
        print("[prep] b16 exact:", np.array_equal(b16[:, :, :L], rb), " pad zero:", float(np.abs(b16[:, :, L:]).max(initial=0)))
        print("[prep] at  exact:", np.array_equal(at[:, :L], ra.transpose(0, 2, 1)), " pad zero:", float(np.abs(at[:, L:]).max(initial=0)))
        print("[prep] bt  exact:", np.array_equal(bt[:, :L], rb.transpose(0, 2, 1)), " pad zero:", float(np.abs(bt[:, L:]).max(initial=0)))
        print("[prep] w16 exact:", np.array_equal(w16, bf16_round(W, odt)), flush=True)
    if args.stage in ("project", "attend"):
        _lib.check(lib.coattn_stage_project(wsp, nbytes, n, c, h, w, FL, st), "project")
        torch.cuda.synchronize()
        qt = segment(ws, "qt", n, c, h, w, odt, (n, Lp, c)).float().cpu().numpy()
        q_ref = np.matmul(at.astype(np.float64), w16.astype(np.float64).T)
        print("[project] rel_l2(qt, bf16-operand fp64 ref):", rel_l2(qt, q_ref), " max|ref|:", float(np.abs(q_ref).max()))
        print("[project] max abs err:", float(np.abs(qt - q_ref).max()), " pad rows max:", float(np.abs(qt[:, L:]).max(initial=0)), flush=True)
        if rel_l2(qt, q_ref) > 1e-2:
            # help locating layout bugs: per-row / per-column error profile
            err = np.abs(qt - q_ref)[0]
            print("[project] err by row block of 8 (first 16):", err.reshape(-1, 8, c).mean(axis=(1, 2))[:16])
            print("[project] err by col block of 16:", err.reshape(Lp, -1, 16).mean(axis=(0, 2)))
    if args.stage == "attend":
        z = torch.zeros(2, n, c, L, device=dev)
        lse = torch.zeros(2, n, L, device=dev)
        _lib.check(lib.coattn_stage_attend(z.data_ptr(), lse.data_ptr(), wsp, nbytes, n, c, h, w, FL | 64, st), "attend")   # K-major operands from stage_prep/project
        torch.cuda.synchronize()
        z = z.cpu().numpy(); lse = lse.cpu().numpy()
        # reference on the SAME bf16 operands (isolates kernel logic from quantisation)
        q = qt[:, :L].astype(np.float64)                    # [n, L, c]
        A = a16[:, :, :L].astype(np.float64); B = b16[:, :, :L].astype(np.float64)
        s = np.matmul(q, B)
        pa = orc.softmax(s, axis=2); pb = orc.softmax(s, axis=1)
        za = np.matmul(B, pa.transpose(0, 2, 1)); zb = np.matmul(A, pb)
        print("[attend] rel_l2 z_a vs bf16-operand ref:", rel_l2(z[0], za), " z_b:", rel_l2(z[1], zb))
        m = s.max(axis=2); lse_a = m + np.log(np.exp(s - m[:, :, None]).sum(axis=2))
        m = s.max(axis=1); lse_b = m + np.log(np.exp(s - m[:, None, :]).sum(axis=1))
        print("[attend] max abs err lse_a:", float(np.abs(lse[0] - lse_a).max()), " lse_b:", float(np.abs(lse[1] - lse_b).max()))
        full = orc.coattention(v_a, v_b, W, g, b)
        print("[attend] rel_l2 z_a vs fp64 oracle:", rel_l2(z[0], full["z_a"].reshape(n, c, L)),
              " z_b:", rel_l2(z[1], full["z_b"].reshape(n, c, L)), " S std:", float(s.std()), " S absmax:", float(np.abs(s).max()), flush=True)
    if args.stage == "gate":
        zz = torch.randn(2, n, c, L, device=dev)
        cat_a = torch.zeros(n, 2 * c, h, w, device=dev); cat_b = torch.zeros(n, 2 * c, h, w, device=dev)
        _lib.check(lib.coattn_stage_gate(zz.data_ptr(), tva.data_ptr(), tvb.data_ptr(), tg.data_ptr(),
                                         None if tb is None else tb.data_ptr(), cat_a.data_ptr(), cat_b.data_ptr(), n, c, h, w, st), "gate")
        torch.cuda.synchronize()
        for side, (cat, v) in enumerate(((cat_a, tva), (cat_b, tvb))):
            zs = zz[side].view(n, c, h, w)
            t = (zs * tg.view(1, c, 1, 1)).sum(1, keepdim=True) + (0 if tb is None else tb)
            ref = torch.cat([zs * torch.sigmoid(t), v], 1)
            print(f"[gate] side {side} max abs err:", float((cat - ref).abs().max()), " passthrough exact:", bool(torch.equal(cat[:, c:], v)), flush=True)
    if args.stage in ("forward", "time"):
        cat_a, cat_b, z, lse = coattention_forward_raw(tva, tvb, tW, tg, tb, bool(args.bf16), single_cta=bool(args.single))
        torch.cuda.synchronize()
        if args.stage == "forward":
            full = orc.coattention(v_a, v_b, W, g, b)
            print("[forward] rel_l2 cat_a:", rel_l2(cat_a.cpu().numpy(), full["cat_a"]), " cat_b:", rel_l2(cat_b.cpu().numpy(), full["cat_b"]))
            print("[forward] rel_l2 z_a:", rel_l2(z[0].cpu().numpy(), full["z_a"].reshape(n, c, L)), " z_b:", rel_l2(z[1].cpu().numpy(), full["z_b"].reshape(n, c, L)))
            print("[forward] max abs err lse_a:", float(np.abs(lse[0].cpu().numpy() - full["lse_a"]).max()),
                  " lse_b:", float(np.abs(lse[1].cpu().numpy() - full["lse_b"]).max()), flush=True)
        # per-stage timing with CUDA events (fused path: prep, project, attend+gate, passthrough; plus the stand-alone gate)
        names = ["prep", "project", "attend", "-", "total", "gate(unfused)"]
        acc = {k: [] for k in names}
        zt = torch.empty(2, n, c, L, device=dev); lt = torch.empty(2, n, L, device=dev); mk = torch.empty(2, n, L, device=dev)
        ca = torch.empty(n, 2 * c, h, w, device=dev); cb = torch.empty(n, 2 * c, h, w, device=dev)
        bp = None if tb is None else tb.data_ptr()
        for itn in range(args.iters + 3):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            ev[0].record()
            if args.unfused_prep:
                _lib.check(lib.coattn_stage_prep(tva.data_ptr(), tvb.data_ptr(), tW.data_ptr(), wsp, nbytes, n, c, h, w, FL, st), "prep"); ev[1].record()
                _lib.check(lib.coattn_stage_project(wsp, nbytes, n, c, h, w, FL, st), "project"); ev[2].record()
                FL |= 64
            else:
                ev[1].record()
                _lib.check(lib.coattn_stage_prep_project(tva.data_ptr(), tvb.data_ptr(), tW.data_ptr(), wsp, nbytes, n, c, h, w, FL, st), "prep_project"); ev[2].record()
            _lib.check(lib.coattn_stage_attend_gate(tva.data_ptr(), tvb.data_ptr(), ca.data_ptr(), cb.data_ptr(), None, lt.data_ptr(), mk.data_ptr(), tg.data_ptr(), bp, wsp, nbytes, n, c, h, w, FL, st), "attend_gate"); ev[3].record()
            ev[4].record()
            _lib.check(lib.coattn_stage_gate(zt.data_ptr(), tva.data_ptr(), tvb.data_ptr(), tg.data_ptr(), bp, ca.data_ptr(), cb.data_ptr(), n, c, h, w, st), "gate"); ev[5].record()
            torch.cuda.synchronize()
            if itn >= 3:
                for k in range(4):
                    acc[names[k]].append(ev[k].elapsed_time(ev[k + 1]))
                acc["total"].append(ev[0].elapsed_time(ev[4]))
                acc["gate(unfused)"].append(ev[4].elapsed_time(ev[5]))
        flops = n * (6.0 * L * L * c + 2.0 * L * c * c)
        for k in names:
            t = np.median(acc[k])
            extra = ""
            if k == "attend":
                extra = f"  algorithmic {n * 6.0 * L * L * c / t / 1e9:.1f} TFLOP/s, executed {n * 8.0 * L * L * c / t / 1e9:.1f} TFLOP/s"
            if k == "gate(unfused)":
                extra = f"  {2 * n * 16.0 * L * c / t / 1e6:.1f} GB/s"
            if k == "prep":
                extra = f"  {(2 * n * 4.0 * L * c + 4 * n * 2.0 * Lp * c) / t / 1e6:.1f} GB/s"
            if k == "total":
                extra = f"  {n / t * 1e3:.1f} modality-calls/s, algorithmic {flops / t / 1e9:.1f} TFLOP/s"
            print(f"[time] {k:14s} median {t:8.4f} ms  min {min(acc[k]):8.4f} ms{extra}", flush=True)


if __name__ == "__main__":
    main()
