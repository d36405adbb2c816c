#!/usr/bin/env python
"""Measures the tcgen05.mma issue rate (cycles per M=128 x N x K=16 fp16 MMA) for the operand shapes the conv kernels
use: N, swizzle mode, K-major vs MN-major, and A start addresses shifted off the swizzle atom.  GPU only."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pointnav_vo_b200 import lib as L  # noqa: E402


def main():
    lib = L.load()
    fn = lib.pnvo_debug_mma_rate
    fn.argtypes = [ctypes.c_int] * 7 + [ctypes.c_void_p, ctypes.c_void_p]
    fn.restype = ctypes.c_int
    out = torch.zeros(148, dtype=torch.int64, device="cuda")
    n_mma = 4096
    print("mode       N  row  shift step  cycles/MMA (1 CTA/SM x 148)   MAC/clk/SM")
    for mn in (0, 1):
        for row in (64, 128):
            for N in (32, 64, 128, 176, 256):
                for shift, step in ((0, 0), (row, 0), (3 * row, 0), (0, row), (0, 1024)):
                    if mn and N * 2 > row and N % (row // 2):
                        continue
                    for n_ctas in (148,):
                        L.check(fn(N, row, shift, step, mn, n_mma, n_ctas, out.data_ptr(), None))
                        torch.cuda.synchronize()
                        cyc = out[:n_ctas].float().mean().item() / n_mma
                        print(f"{'MN' if mn else 'K '}-major {N:4d} {row:4d} {shift:5d} {step:4d}   {cyc:8.1f}"
                              f"                      {128 * N * 16 / cyc:8.0f}")


if __name__ == "__main__":
    main()
