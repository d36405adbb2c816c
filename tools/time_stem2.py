#!/usr/bin/env python
"""Times the two stem forward kernels at the bench size (B=256, 192x341, 30 channels).  GPU only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pointnav_vo_b200 import lib as L  # noqa: E402

B, IH, IW, Cin = 256, 192, 341, 30
dev = "cuda"
OH, OW = (IH - 1) // 2 + 1, (IW - 1) // 2 + 1
w = torch.randn(32, Cin, 7, 7, device=dev) * 0.03
Wp = L.load().pnvo_stem_padded_width(IW)
xp = torch.zeros(B, IH, Wp, 32, device=dev, dtype=torch.float16)
xp[:, :, 3:3 + IW, :Cin] = torch.randn(B, IH, IW, Cin, device=dev).half()
y = torch.zeros(B, OH, OW, 32, dtype=torch.float16, device=dev)
stats = torch.zeros(B, 16, 2, device=dev, dtype=torch.float64)
wr = torch.zeros(4 * 7 * 32, 64, dtype=torch.float16, device=dev)
for name, pack, op in (("stem v1 (N=32 raster)", L.op_pack_w_stem(w, wr, Cin), L.op_conv_stem(xp, wr, y, stats, B, IH, IW, 16, 2, 2)),
                       ("stem v2 (pixels as N)", L.op_pack_w_stem2(w, wr, Cin), L.op_conv_stem2(xp, wr, y, stats, B, IH, IW, 16, 2))):
    L.run_ops([pack])
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
    print(f"{name}: {ms:.3f} ms  {2.0 * B * OH * OW * 32 * Cin * 49 / ms / 1e9:.1f} TFLOP/s")

dy = torch.randn(B, OH, OW, 32, device=dev).half()
dw = torch.zeros(32, 1600, device=dev)
for name, op in (("stem wgrad v1", L.op_wgrad_stem(xp, dy, dw, B, IH, IW, 1600, 48)),
                 ("stem wgrad v2 (4 dy rows as N)", L.op_wgrad_stem2(xp, dy, dw, B, IH, IW, 1600))):
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
    print(f"{name}: {ms:.3f} ms  {2.0 * B * OH * OW * 32 * Cin * 49 / ms / 1e9:.1f} TFLOP/s")
