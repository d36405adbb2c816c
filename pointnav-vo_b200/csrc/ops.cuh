// Internal argument structs + launchers shared by the op dispatcher (api.cu) and the kernels.
#pragma once
#include "common.cuh"

namespace pnvo {

struct ConvArgs {
  const __half* x;    // input NHWC [B, IH, IW, Cin] fp16
  const __half* w;    // packed weights [n_total][w_ld] fp16, K-major, zero padded
  void* y;            // output [M][ldo] fp16 (or fp32 when out_fp32)
  __half* y_lo;       // split-fp16 mode, optional: the output as value + residual fp16 planes (y = fp16(acc), y_lo =
                      // fp16(acc - y)) instead of fp32 -- same bytes, but the backward pass can read the value plane alone
  const __half* add;  // optional [M][ldo] fp16 added before the store (dgrad accumulation)
  const float* x_c;   // optional, direct one-channel stem only (conv_direct.cu): the image as a compact fp32 plane [B, IH, IW]
  const __half* x_lo; // split-fp16 mode (both or neither): residual planes x - fp16(x), w - fp16(w) in the same layouts;
  const __half* w_lo; //   the kernel accumulates x*w + x_lo*w + x*w_lo (3 MMAs per product, ~fp32 operand precision)
  double* stats;      // optional GroupNorm partial sums [B][G][2] (sum, sum of squares), pre-zeroed; fp64 accumulators:
                      // the ORDER of the atomics then no longer shows in the fp32 mean / rstd (reproducible forward)
  int B, IH, IW, Cin;
  int OH, OW;
  int R, S, mul, pad, pad_w, div;  // t = o*mul - pad + r; tap valid iff t >= 0, t % div == 0, t/div < I
  int w_ld;
  int n_total, n_store, ldo;
  int cpg, G;
  int out_fp32;
  int force_generic;   // 0: best kernel; 1: cp.async im2col producer; 2: TMA im2col producer (A/B testing)
  int cin_real;        // input channels that are not zero padding (0 = unknown): 1 selects the direct fp32 stem kernel
  // strided output (parity-class data gradient of a stride-2 convolution): output pixel (b, oh, ow) of the GEMM is stored
  // at (b, oh * o_mul + o_off_h, ow * o_mul + o_off_w) of a [B, o_H, o_W] tensor (skipped when outside); 0 = dense
  int o_mul, o_off_h, o_off_w, o_H, o_W;
  // asym != 0: padding `pad` / `pad_w` on the low side and pad_hi_h / pad_hi_w on the high side (default: symmetric)
  int asym, pad_hi_h, pad_hi_w;
  int tma, chunk_k;    // derived: TMA producer on/off, K elements per pipeline stage (64 or 32)
  // derived by conv_plan
  int cin_log2, cmask, M, K, nkb, N, tmem_cols, stages, lookahead, smem_bytes, grid_x, grid_y;
};
int conv_plan(ConvArgs& a);
// conv_raster.cu: persistent no-im2col kernel for 3x3 / stride 1 / pad 1 with 32 or 64 channels
int conv_raster_supported(const ConvArgs& a);
int conv_raster_launch(const ConvArgs& a, cudaStream_t st);
// conv_raster128.cu: the same raster for 128 input channels, weights streamed through a ring shared by all M tiles
int conv_raster128_supported(const ConvArgs& a);
int conv_raster128_launch(const ConvArgs& a, cudaStream_t st);
// conv_direct.cu: one-channel 7x7 / stride 2 stem (depth-only policy encoder) on the fp32 CUDA cores
int conv_direct1_supported(const ConvArgs& a);
int conv_direct1_launch(const ConvArgs& a, cudaStream_t st);
int conv_launch(ConvArgs a, cudaStream_t st);

struct WgradArgs {
  const __half* x;   // conv input NHWC [B, IH, IW, Cin] fp16
  const float* x_c;  // optional, direct one-channel stem only: the image as a compact fp32 plane [B, IH, IW]
  const __half* dy;  // gradient of the conv output [M][ld_dy] fp16 (M = B*OH*OW)
  float* dw;         // packed fp32 gradient [n_total][w_ld], accumulated with atomics (pre-zeroed)
  int B, IH, IW, Cin;
  int OH, OW;
  int R, S, mul, pad, pad_w;
  int w_ld, n_total, ld_dy;
  int x_row_pitch;     // pixels between input rows (0 = IW); TMA path only
  int force_generic;
  int cin_real;        // as ConvArgs::cin_real
  int tma, chunk_k;
  // derived
  int cin_log2, cmask, M, K, n_mtiles, mt, N, n_ntiles, tmem_cols, stages, lookahead, smem_bytes, grid_x, grid_y, grid_z, chunks_per_split;
};
int conv_stem_fwd_launch(const __half* x, const __half* wr, void* y, double* stats, int B, int IH, int IW, int G, int cpg,
                         int stages, cudaStream_t st);
int stem_padded_width(int IW);
int conv_stem_wgrad_launch(const __half* x, const __half* dy, float* dw, int w_ld, int B, int IH, int IW,
                           int rows_per_cta, cudaStream_t st);
int pack_w_stem_launch(const float* w, int Cin, __half* wr, cudaStream_t st);
// conv_stem2.cu: pixels-as-N formulation of the stem (full-rate MMAs, resident weights, persistent)
int conv_stem2_supported(int IH, int IW);
// x_lo (split-fp16 mode): residual plane of x, staged row by row next to x against the same weights; add: fp16
// [B,OH,OW,32] added in the epilogue (the separately computed w_lo * x product); out_fp32: fp32 raw output
int conv_stem2_fwd_launch(const __half* x, const __half* wr, void* y, double* stats, int B, int IH, int IW, int G, int cpg,
                          cudaStream_t st, const __half* x_lo = nullptr, const __half* add = nullptr, int out_fp32 = 0,
                          const float* bias5 = nullptr, __half* y_lo = nullptr);
int pack_w_stem2_launch(const float* w, int Cin, __half* wr, int lo, cudaStream_t st);
int conv_stem_wgrad2_supported(int IH, int IW);
int conv_stem_wgrad2_launch(const __half* x, const __half* dy, float* dw, int w_ld, int B, int IH, int IW, cudaStream_t st);
int wgrad_plan(WgradArgs& a);
// conv_wgrad_raster.cu: persistent no-im2col weight gradient for 3x3 / stride 1 / pad 1 with 32 or 64 channels
int wgrad_raster_supported(const WgradArgs& a);
int wgrad_raster_launch(const WgradArgs& a, cudaStream_t st);
int wgrad_direct1_supported(const WgradArgs& a);
int wgrad_direct1_launch(const WgradArgs& a, cudaStream_t st);
int wgrad_launch(WgradArgs a, cudaStream_t st);

}  // namespace pnvo
