#!/usr/bin/env python
"""Validation + timing of the raster stem kernels against torch fp32 (GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from pointnav_vo_b200 import lib as L

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(got, ref):
    got, ref = got.float(), ref.float()
    return ((got - ref).abs().max() / (ref.pow(2).mean().sqrt() + 1e-12)).item(), int((~torch.isfinite(got)).sum())


def run(B, IH, IW, Cin, stages, time_it=False):
    dev = "cuda"
    torch.manual_seed(0)
    OH, OW = (IH - 1) // 2 + 1, (IW - 1) // 2 + 1
    w = (torch.randn(32, Cin, 7, 7, device=dev) / (Cin * 49) ** 0.5).contiguous()
    Wp = L.load().pnvo_stem_padded_width(IW)
    xp = torch.zeros(B, IH, Wp, 32, device=dev, dtype=torch.float16)
    x = torch.randn(B, IH, IW, 32, device=dev).half()
    x[..., Cin:] = 0
    xp[:, :, 3:3 + IW] = x
    wr = torch.zeros(7 * 4 * 32, 64, dtype=torch.float16, device=dev)
    y = torch.zeros(B, OH, OW, 32, dtype=torch.float16, device=dev)
    stats = torch.zeros(B, 16, 2, device=dev, dtype=torch.float64)
    ops = [L.op_pack_w_stem(w, wr, Cin), L.op_conv_stem(xp, wr, y, stats, B, IH, IW, 16, 2, stages)]
    L.run_ops(ops)
    torch.cuda.synchronize()
    ref = F.conv2d(x[..., :Cin].float().permute(0, 3, 1, 2), w.half().float(), None, 2, 3)
    r, bad = rel(y.permute(0, 3, 1, 2), ref)
    rs = ref.reshape(B, 16, -1)
    sref = torch.stack((rs.sum(-1), rs.pow(2).sum(-1)), -1)
    r2, _ = rel(stats, sref)
    print(f"stem fwd B={B} {IH}x{IW} Cin={Cin} stages={stages}: rel={r:.3e} nonfinite={bad} stats rel={r2:.3e}", flush=True)
    if time_it:
        prog = L.Program([ops[1]])
        for _ in range(3):
            prog.run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            prog.run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * B * OH * OW * 32 * 30 * 49
        print(f"   {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s (30-channel FLOPs)", flush=True)


if __name__ == "__main__":
    run(1, 16, 341, 30, 2)
    run(2, 192, 341, 30, 4)
    run(2, 192, 341, 30, 2)
    for st in (2, 3, 4, 5):
        run(256, 192, 341, 30, st, time_it=True)


def run_wgrad(B, IH, IW, Cin, rows, time_it=False):
    dev = "cuda"
    torch.manual_seed(1)
    OH, OW = (IH - 1) // 2 + 1, (IW - 1) // 2 + 1
    Wp = L.load().pnvo_stem_padded_width(IW)
    xp = torch.zeros(B, IH, Wp, 32, device=dev, dtype=torch.float16)
    x = torch.randn(B, IH, IW, 32, device=dev).half()
    x[..., Cin:] = 0
    xp[:, :, 3:3 + IW] = x
    dy = torch.randn(B, OH, OW, 32, device=dev).half()
    w_ld = 1600
    dw = torch.zeros(32, w_ld, device=dev)
    op = L.op_wgrad_stem(xp, dy, dw, B, IH, IW, w_ld, rows)
    L.run_ops([op])
    torch.cuda.synchronize()
    w = torch.zeros(32, Cin, 7, 7, device=dev, requires_grad=True)
    F.conv2d(x[..., :Cin].float().permute(0, 3, 1, 2), w, None, 2, 3).backward(dy.float().permute(0, 3, 1, 2))
    got = dw[:, :49 * 32].reshape(32, 7, 7, 32)[..., :Cin].permute(0, 3, 1, 2)
    r, bad = rel(got, w.grad)
    print(f"stem wgrad B={B} {IH}x{IW} rows/cta={rows}: rel={r:.3e} nonfinite={bad}", flush=True)
    if time_it:
        prog = L.Program([op])
        for _ in range(3):
            prog.run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            prog.run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"   {ms:.3f} ms  {2.0 * B * OH * OW * 32 * 30 * 49 / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    run_wgrad(1, 16, 341, 30, 8)
    run_wgrad(2, 192, 341, 30, 32)
    for rows in (16, 32, 48, 96):
        run_wgrad(256, 192, 341, 30, rows, time_it=True)
