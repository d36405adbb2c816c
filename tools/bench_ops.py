#!/usr/bin/env python
"""Per-kernel timings of the conv ops at benchmark sizes (B=256), CUDA events, inputs larger than L2 or
L2 flushed between launches.  Used to A/B kernel variants on the GPU box in one call."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pointnav_vo_b200 import lib as L  # noqa: E402
from pointnav_vo_b200.engine import ConvLayer  # noqa: E402

LAYERS = {
    "conv1": (30, 32, 7, 7, 2, 3, 192, 341, 32),
    "layer1": (32, 32, 3, 3, 1, 1, 48, 86, None),
    "layer2.0a": (32, 64, 3, 3, 2, 1, 48, 86, None),
    "layer2": (64, 64, 3, 3, 1, 1, 24, 43, None),
    "layer3": (128, 128, 3, 3, 1, 1, 12, 22, None),
    "layer4": (256, 256, 3, 3, 1, 1, 6, 11, None),
}


def time_op(op, reps=5, flush=None):
    prog = L.Program([op])
    for _ in range(2):
        prog.run()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        prog.run()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--layers", default=",".join(LAYERS))
    ap.add_argument("--flags", default="0,1")
    a = ap.parse_args()
    B = a.batch
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name in a.layers.split(","):
        Cin, Cout, R, S, st, pad, IH, IW, cpad = LAYERS[name]
        c = ConvLayer("w", Cin, Cout, R, S, st, pad, IH, IW, need_dgrad=(name != "conv1"), cin_pad=cpad)
        c.alloc(dev, True)
        w = torch.randn(Cout, Cin, R, S, device=dev) * 0.05
        x = torch.randn(B, IH, IW, c.cin_pad, device=dev).half()
        y = torch.empty(B, c.OH, c.OW, c.cout_pad, dtype=torch.float16, device=dev)
        dy = torch.randn(B, c.OH, c.OW, c.cout_pad, device=dev).half()
        gx = torch.empty_like(x)
        stats = torch.zeros(B, 16, 2, device=dev, dtype=torch.float64)
        L.run_ops([c.op_pack(w)])
        gf = c.flops(B) / 1e9
        for flag in [int(f) for f in a.flags.split(",")]:
            res = []
            ops = {"fwd": c.op_fwd(x, y, B, stats, c.cout_pad // 16, 16), "wgrad": c.op_wgrad(x, dy, B)}
            if c.need_dgrad:
                ops["dgrad"] = c.op_dgrad(dy, gx, B)
            for k, op in ops.items():
                op.i[19] = flag
                ms = time_op(op, flush=flush)
                res.append(f"{k} {ms:7.3f} ms {gf / ms:7.1f} TFLOP/s")
            print(f"{name:10s} flag={flag}  " + " | ".join(res), flush=True)


if __name__ == "__main__":
    main()
