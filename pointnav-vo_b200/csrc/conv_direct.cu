// Direct fp32 convolution for ONE-channel 7x7 / stride 2 / pad 3 stems: the depth-only actor-critic encoder
// (resnet_policy.py:146-174 -> resnet.py:156-164 conv1 with in_channels = 1 after avg_pool2d(2); ddppo_pointnav.yaml ships
// SENSORS = ["DEPTH_SENSOR"]).
//
// With one input channel the convolution has K = 49: as an implicit GEMM its operands are padded 8x (8-channel pixels) and
// N = 32 runs the tensor pipe at 40 %, and the generic kernel's im2col gather moves 16 bytes per useful fp16 -- measured
// 38.9 ms per 8192-frame PPO minibatch (21 TFLOP/s on padded FLOPs) and 10.2 ms for the weight gradient.  The work is
// 52 G fused multiply-adds per minibatch: ~3 ms on the fp32 CUDA cores, exact fp32 products (better than the split-fp16
// representation), HBM traffic = the output.
//
//   forward : block = (sample, 16 x 32 output tile), 256 threads; thread (warp w, lane l) owns output pixels (w, l) and (w + 8, l)
//             with all 32 output channels in registers.  The 37 x 69 input tile is staged in shared memory split by column
//             parity (stride-2 taps of consecutive lanes then hit consecutive words: conflict-free); weights (value +
//             residual planes summed to fp32) sit in shared memory and are read as broadcast 128-bit words.
//             Epilogue as the tensor-core kernels': GroupNorm partial sums (fp64 atomics per sample and group), output as
//             value + residual fp16 planes / fp32 / fp16.
//   wgrad   : persistent blocks of 7 warps; warp r <-> filter row r, lane <-> output channel; acc[7] (the 7 taps of the row)
//             stays in registers over all tiles of the block; per output pixel 1 dy load + 2 new input values (sliding
//             window over the parity-split tile) + 7 FMAs; one atomic flush per block.
#include <algorithm>

#include "common.cuh"
#include "ops.cuh"

namespace pnvo {

static constexpr int kD1TileH = 16, kD1TileW = 32;
static constexpr int kD1InH = 2 * kD1TileH + 5;        // 37 input rows
static constexpr int kD1InWHalf = kD1TileW + 3;        // 35 even + 35 odd columns (69 input columns -> 35 + 34)
static constexpr int kD1RowWords = 2 * kD1InWHalf + 2; // 72 words per staged row (even plane, then odd plane)

struct Direct1Args {
  const __half* x;
  const __half* x_lo;   // nullable
  const float* x_c;     // nullable: the image as a compact fp32 plane [B, IH, IW] (4 instead of 2 x 16 bytes per pixel read)
  const __half* w;
  const __half* w_lo;   // nullable
  void* y;
  __half* y_lo;         // nullable
  const __half* dy;     // wgrad
  float* dw;            // wgrad
  double* stats;
  int B, IH, IW, OH, OW;
  int cpad;             // input pixel stride in halves (channel 0 is the image)
  int w_ld;             // packed weight row length; element (n, tap) at n * w_ld + tap * cpad
  int ldo;              // output pixel stride in elements (32)
  int ld_dy;
  int out_fp32, cpg, G;
  int tiles_h, tiles_w;
};

// 32 channels -> 32 / CPG groups, (sum, sumsq) added at gs[2g], gs[2g + 1]
template <int CPG>
__device__ __forceinline__ void d1_group_sums(const float* v, float* gs) {
#pragma unroll
  for (int g = 0; g < 32 / CPG; ++g) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < CPG; ++c) {
      const float x = v[g * CPG + c];
      s1 += x;
      s2 = fmaf(x, x, s2);
    }
    gs[2 * g] += s1;
    gs[2 * g + 1] += s2;
  }
}

// stage the input tile of output tile (b, th, tw): s_x[row][parity][j] = x[2*oh0 - 3 + row][2*ow0 - 3 + 2j + parity]
__device__ __forceinline__ void d1_stage_input(const Direct1Args& a, int b, int oh0, int ow0, float* s_x) {
  const int ih0 = 2 * oh0 - 3, iw0 = 2 * ow0 - 3;
  for (int i = threadIdx.x; i < kD1InH * 2 * kD1InWHalf; i += blockDim.x) {
    const int row = i / (2 * kD1InWHalf), col = i - row * (2 * kD1InWHalf);   // col = tile column 0..69
    const int ih = ih0 + row, iw = iw0 + col;
    float v = 0.f;
    if (ih >= 0 && ih < a.IH && iw >= 0 && iw < a.IW) {
      const int64_t px = (static_cast<int64_t>(b) * a.IH + ih) * a.IW + iw;
      if (a.x_c) {
        v = __ldg(a.x_c + px);
      } else {
        v = __half2float(__ldg(a.x + px * a.cpad));
        if (a.x_lo) v += __half2float(__ldg(a.x_lo + px * a.cpad));
      }
    }
    s_x[row * kD1RowWords + (col & 1) * (kD1InWHalf + 1) + (col >> 1)] = v;
  }
}

__global__ void __launch_bounds__(256) conv_direct1_fwd_kernel(const Direct1Args a) {
  __shared__ __align__(16) float s_w[49 * 32];
  __shared__ float s_x[kD1InH * kD1RowWords];
  __shared__ double s_red[8][32];
  const int tw = blockIdx.x % a.tiles_w, th = (blockIdx.x / a.tiles_w) % a.tiles_h, b = blockIdx.x / (a.tiles_w * a.tiles_h);
  const int oh0 = th * kD1TileH, ow0 = tw * kD1TileW;
  for (int i = threadIdx.x; i < 49 * 32; i += blockDim.x) {
    const int tap = i >> 5, n = i & 31;
    const int64_t o = static_cast<int64_t>(n) * a.w_ld + tap * a.cpad;
    float v = __half2float(a.w[o]);
    if (a.w_lo) v += __half2float(a.w_lo[o]);
    s_w[i] = v;
  }
  d1_stage_input(a, b, oh0, ow0, s_x);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[2][32];
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int n = 0; n < 32; ++n) acc[q][n] = 0.f;
#pragma unroll 1
  for (int r = 0; r < 7; ++r) {
    float xv[2][7];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float* row = s_x + (2 * (warp + 8 * q) + r) * kD1RowWords;
#pragma unroll
      for (int s = 0; s < 7; ++s) xv[q][s] = row[(s & 1) * (kD1InWHalf + 1) + lane + (s >> 1)];
    }
#pragma unroll
    for (int s = 0; s < 7; ++s) {
      const float4* wp = reinterpret_cast<const float4*>(s_w + (r * 7 + s) * 32);
#pragma unroll
      for (int n4 = 0; n4 < 8; ++n4) {
        const float4 w4 = wp[n4];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          acc[q][4 * n4] = fmaf(xv[q][s], w4.x, acc[q][4 * n4]);
          acc[q][4 * n4 + 1] = fmaf(xv[q][s], w4.y, acc[q][4 * n4 + 1]);
          acc[q][4 * n4 + 2] = fmaf(xv[q][s], w4.z, acc[q][4 * n4 + 2]);
          acc[q][4 * n4 + 3] = fmaf(xv[q][s], w4.w, acc[q][4 * n4 + 3]);
        }
      }
    }
  }
  // ---- epilogue: stores + GroupNorm partial sums
  float gs[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) gs[i] = 0.f;
  const int ow = ow0 + lane;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int oh = oh0 + warp + 8 * q;
    const bool valid = oh < a.OH && ow < a.OW;
    if (!valid) continue;
    const int64_t gofs = ((static_cast<int64_t>(b) * a.OH + oh) * a.OW + ow) * a.ldo;
    if (a.stats) {
      if (a.cpg == 2) d1_group_sums<2>(acc[q], gs);
      else if (a.cpg == 4) d1_group_sums<4>(acc[q], gs);
      else if (a.cpg == 8) d1_group_sums<8>(acc[q], gs);
      else if (a.cpg == 16) d1_group_sums<16>(acc[q], gs);
      else d1_group_sums<32>(acc[q], gs);
    }
    if (a.y_lo) {
      uint4* yh = reinterpret_cast<uint4*>(static_cast<__half*>(a.y) + gofs);
      uint4* yl = reinterpret_cast<uint4*>(a.y_lo + gofs);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint4 hi, lo;
        split8(acc[q] + 8 * k, hi, lo);
        yh[k] = hi;
        yl[k] = lo;
      }
    } else if (a.out_fp32) {
      float4* yp = reinterpret_cast<float4*>(static_cast<float*>(a.y) + gofs);
#pragma unroll
      for (int k = 0; k < 8; ++k) yp[k] = make_float4(acc[q][4 * k], acc[q][4 * k + 1], acc[q][4 * k + 2], acc[q][4 * k + 3]);
    } else {
      __half* yp = static_cast<__half*>(a.y) + gofs;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint4 uu;
        __half2* h2 = reinterpret_cast<__half2*>(&uu);
#pragma unroll
        for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(acc[q][8 * k + 2 * e], acc[q][8 * k + 2 * e + 1]);
        *reinterpret_cast<uint4*>(yp + 8 * k) = uu;
      }
    }
  }
  if (a.stats) {
    // reduce-scatter over the 32 lanes (lane i ends with the warp total of value i), then over the 8 warps
    int off = 16;
#pragma unroll
    for (int cnt = 16; cnt >= 1; cnt >>= 1) {
      const bool up = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < cnt; ++i) {
        const float send = up ? gs[i] : gs[i + cnt];
        const float recv = __shfl_xor_sync(0xffffffffu, send, off);
        gs[i] = (up ? gs[i + cnt] : gs[i]) + recv;
      }
      off >>= 1;
    }
    s_red[warp][lane] = static_cast<double>(gs[0]);
    __syncthreads();
    if (threadIdx.x < 2 * a.G) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += s_red[w][threadIdx.x];
      atomicAdd(a.stats + static_cast<int64_t>(b) * a.G * 2 + threadIdx.x, t);
    }
  }
}

// dw[n][tap] += sum over all output pixels of dy[.., n] * x[2 oh + r - 3][2 ow + s - 3]
__global__ void __launch_bounds__(224) conv_direct1_wgrad_kernel(const Direct1Args a, int n_tiles_total) {
  __shared__ float s_x[kD1InH * kD1RowWords];
  __shared__ __align__(16) __half s_dy[kD1TileH * kD1TileW * 32];
  const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;   // filter row, output channel
  float acc[7];
#pragma unroll
  for (int s = 0; s < 7; ++s) acc[s] = 0.f;
  for (int t = blockIdx.x; t < n_tiles_total; t += gridDim.x) {
    const int tw = t % a.tiles_w, th = (t / a.tiles_w) % a.tiles_h, b = t / (a.tiles_w * a.tiles_h);
    const int oh0 = th * kD1TileH, ow0 = tw * kD1TileW;
    __syncthreads();   // previous tile fully consumed
    d1_stage_input(a, b, oh0, ow0, s_x);
    // dy tile: 16 x 32 pixels x 32 channels fp16, 16-byte vectors; outside the image -> zero
    for (int i = threadIdx.x; i < kD1TileH * kD1TileW * 4; i += blockDim.x) {
      const int q = i & 3, px = i >> 2;
      const int oh = oh0 + px / kD1TileW, ow = ow0 + px % kD1TileW;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (oh < a.OH && ow < a.OW)
        v = __ldg(reinterpret_cast<const uint4*>(a.dy + ((static_cast<int64_t>(b) * a.OH + oh) * a.OW + ow) * a.ld_dy) + q);
      reinterpret_cast<uint4*>(s_dy)[i] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int oh = 0; oh < kD1TileH; ++oh) {
      const float* row = s_x + (2 * oh + r) * kD1RowWords;
      const float* ev = row;                       // even columns: taps s = 0, 2, 4, 6 of output ow -> ev[ow + s / 2]
      const float* od = row + kD1InWHalf + 1;      // odd columns:  taps s = 1, 3, 5       -> od[ow + (s - 1) / 2]
      float e0 = ev[0], e1 = ev[1], e2 = ev[2], o0 = od[0], o1 = od[1];
      const __half* dyr = s_dy + oh * kD1TileW * 32 + lane;
#pragma unroll 4
      for (int ow = 0; ow < kD1TileW; ++ow) {
        const float e3 = ev[ow + 3], o2 = od[ow + 2];
        const float d = __half2float(dyr[ow * 32]);
        acc[0] = fmaf(d, e0, acc[0]);
        acc[1] = fmaf(d, o0, acc[1]);
        acc[2] = fmaf(d, e1, acc[2]);
        acc[3] = fmaf(d, o1, acc[3]);
        acc[4] = fmaf(d, e2, acc[4]);
        acc[5] = fmaf(d, o2, acc[5]);
        acc[6] = fmaf(d, e3, acc[6]);
        e0 = e1; e1 = e2; e2 = e3; o0 = o1; o1 = o2;
      }
    }
  }
#pragma unroll
  for (int s = 0; s < 7; ++s) atomicAdd(a.dw + static_cast<int64_t>(lane) * a.w_ld + (r * 7 + s) * a.cpad, acc[s]);
}

static bool direct1_geometry(int R, int S, int mul, int pad, int pad_w, int div, int n_total, int IH, int IW, int OH, int OW) {
  return R == 7 && S == 7 && mul == 2 && pad == 3 && pad_w == 3 && div == 1 && n_total == 32 && OH == (IH - 1) / 2 + 1 &&
         OW == (IW - 1) / 2 + 1;
}

int conv_direct1_supported(const ConvArgs& a) {
  if (a.cin_real != 1 || a.add) return 0;
  if (!direct1_geometry(a.R, a.S, a.mul, a.pad, a.pad_w, a.div, a.n_total, a.IH, a.IW, a.OH, a.OW)) return 0;
  if (a.n_store != 32 || a.ldo != 32) return 0;
  if (a.stats && (a.G * a.cpg != 32 || a.G > 16 || !(a.cpg == 2 || a.cpg == 4 || a.cpg == 8 || a.cpg == 16 || a.cpg == 32))) return 0;
  if (a.y_lo && a.out_fp32) return 0;
  return 1;
}

int conv_direct1_launch(const ConvArgs& c, cudaStream_t st) {
  PNVO_REQUIRE(conv_direct1_supported(c), "conv_direct1: unsupported geometry");
  if (c.B <= 0) return 0;
  Direct1Args a{};
  a.x = c.x; a.x_lo = c.x_lo; a.x_c = c.x_c; a.w = c.w; a.w_lo = c.w_lo; a.y = c.y; a.y_lo = c.y_lo; a.stats = c.stats;
  a.B = c.B; a.IH = c.IH; a.IW = c.IW; a.OH = c.OH; a.OW = c.OW; a.cpad = c.Cin; a.w_ld = c.w_ld; a.ldo = c.ldo;
  a.out_fp32 = c.out_fp32; a.cpg = c.cpg; a.G = c.G;
  a.tiles_h = ceil_div(c.OH, kD1TileH);
  a.tiles_w = ceil_div(c.OW, kD1TileW);
  const int64_t blocks = static_cast<int64_t>(c.B) * a.tiles_h * a.tiles_w;
  PNVO_REQUIRE(blocks < (1ll << 31), "conv_direct1: too many tiles");
  conv_direct1_fwd_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(a);
  count_launch();
  return check_launch("conv_direct1_fwd");
}

int wgrad_direct1_supported(const WgradArgs& a) {
  if (a.cin_real != 1 || a.x_row_pitch != 0) return 0;
  const int OH = (a.IH - 1) / 2 + 1, OW = (a.IW - 1) / 2 + 1;
  return direct1_geometry(a.R, a.S, a.mul, a.pad, a.pad_w, 1, a.n_total, a.IH, a.IW, a.OH, a.OW) && a.OH == OH && a.OW == OW &&
         a.ld_dy % 8 == 0 && a.ld_dy >= 32;
}

int wgrad_direct1_launch(const WgradArgs& w, cudaStream_t st) {
  PNVO_REQUIRE(wgrad_direct1_supported(w), "wgrad_direct1: unsupported geometry");
  if (w.B <= 0) return 0;
  Direct1Args a{};
  a.x = w.x; a.x_c = w.x_c; a.dy = w.dy; a.dw = w.dw;
  a.B = w.B; a.IH = w.IH; a.IW = w.IW; a.OH = w.OH; a.OW = w.OW; a.cpad = w.Cin; a.w_ld = w.w_ld; a.ld_dy = w.ld_dy;
  a.tiles_h = ceil_div(w.OH, kD1TileH);
  a.tiles_w = ceil_div(w.OW, kD1TileW);
  const int64_t tiles = static_cast<int64_t>(w.B) * a.tiles_h * a.tiles_w;
  PNVO_REQUIRE(tiles < (1ll << 31), "wgrad_direct1: too many tiles");
  const int grid = static_cast<int>(std::min<int64_t>(tiles, 148 * 4));
  conv_direct1_wgrad_kernel<<<grid, 224, 0, st>>>(a, static_cast<int>(tiles));
  count_launch();
  return check_launch("conv_direct1_wgrad");
}

}  // namespace pnvo
