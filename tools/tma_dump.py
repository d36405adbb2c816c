import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pointnav_vo_b200 import lib as L
lib = L.load()
lib.pnvo_debug_tma_dump.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
lib.pnvo_debug_tma_dump.restype = ctypes.c_int
B, IH, IW = 2, 8, 341
# element value encodes (pixel iw, chunk): fp16 value = iw + chunk/8 -> use int16 view instead: store raw int16 = iw*4 + chunk
x = torch.zeros(B, IH, IW, 32, dtype=torch.int16)
iw = torch.arange(IW).view(1, 1, IW, 1)
ch = (torch.arange(32) // 8).view(1, 1, 1, 32)
x[:] = (iw * 4 + ch + 1).to(torch.int16)
x[1] += 2000
xd = x.cuda().view(torch.float16)
box_px = 174
n16 = (2 * box_px * 128) // 16 + 64
out = torch.zeros(n16 * 16, dtype=torch.uint8, device="cuda")
rc = lib.pnvo_debug_tma_dump(xd.data_ptr(), B, IH, IW, box_px, 3, 0, -3, out.data_ptr(), n16, None)
torch.cuda.synchronize()
print("rc", rc, lib.pnvo_last_error())
o = out.cpu().numpy().view(np.int16).reshape(-1, 8)  # one row per 16-byte chunk
# for each chunk print decoded (iw, ch) of first element; all 8 elements should be equal
for i in list(range(0, 48)) + list(range(box_px * 4 - 8, box_px * 4 + 24)) + list(range(box_px * 8 - 8, box_px * 8 + 16)) + list(range(n16 - 72, n16 - 56)):
    v = o[i]
    tag = "uniform" if (v == v[0]).all() else "MIXED"
    val = int(v[0])
    if val == -1:
        desc = "untouched"
    elif val == 0:
        desc = "zero(OOB)"
    else:
        desc = f"iw={(val - 1) // 4} ch={(val - 1) % 4}"
    print(i, i * 16, desc, tag)
