// TMA tensor-map construction (host) and bulk-tensor load wrappers (device).
// Maps are built with the driver's cuTensorMapEncode{Tiled,Im2col}, fetched through
// cudaGetDriverEntryPoint (no link-time dependency on libcuda), and cached per process keyed by
// (address, geometry) -- the only global state of the library besides the last-error string.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace pnvo {

struct ConvTmaps {
  CUtensorMap a;     // activations: im2col map (fprop / wgrad A operand)
  CUtensorMap b;     // weights (fprop) or output gradient (wgrad): 2-D tiled map
  CUtensorMap a_lo;  // split-fp16 forward: the residual planes x - fp16(x), w - fp16(w) (same geometry as a / b)
  CUtensorMap b_lo;
};

// NHWC fp16 activations [N, H, W, C]; a load fetches `pixels` consecutive output positions (w fastest,
// wrapping over h and n inside the padded bounding box) x `channels` channels of one filter tap.
// pad = low-side padding in H (pad_w_lo in W, default pad); pad_h_hi / pad_w_hi = high-side padding (default: symmetric)
int tmap_im2col(CUtensorMap* out, const void* x, int N, int H, int W, int C, int R, int S, int stride, int pad,
                int channels, int pixels, int row_pitch_px = 0, int pad_w_lo = -1, int pad_h_hi = -1, int pad_w_hi = -1);
// NHWC fp16 activations [N, H, W, C]: plain tiled boxes of box_w pixels x C channels of one image row,
// SWIZZLE_128B (C*2 <= 128 bytes), out-of-range pixels / rows zero-filled
int tmap_tiled4d(CUtensorMap* out, const void* x, int N, int H, int W, int C, int box_w, int swizzle_bytes = 128,
                 int box_h = 1, int box_c = 0);  // box_c: channels per box (0 = all C)
// row-major fp16 matrix [rows, ld] (cols valid columns); box = box_rows x box_cols
int tmap_tiled2d(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                 int box_cols);

#ifdef __CUDACC__
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c, int w,
                                                   int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c, int w, int h, int n) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n)
      : "memory");
}
// L2 prefetch of the same box (no shared-memory destination, no barrier): the later tma_load_4d then hits L2
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* m, int c, int w, int h, int n) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c), "r"(w), "r"(h), "r"(n)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
// UMMA smem descriptor with an explicit swizzle mode: 128-byte rows (layout 2) or 64-byte rows (layout 4)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, int row_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(row_bytes == 128 ? 2 : 4) << 61;  // SWIZZLE_128B : SWIZZLE_64B
  return d;
}
#endif

}  // namespace pnvo
