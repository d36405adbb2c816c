// Exact-input formulation of the stem convolution in split precision (vo_cnn.py:110-176 + resnet.py:156-164).
//
// The split-fp16 stem needs the NORMALISED input (x - mean) / std to ~fp32 precision (the first layer is the
// precision-critical one: a single fp16 rounding of its input alone costs up to 3e-3 at the network output), which the
// generic path buys with a residual plane x_lo: 2.1 GB more HBM traffic and a third tensor-core product.  But the raw
// observations are EXACTLY representable in fp16 -- rgb bytes, one-hot depth bins, fp16 depth -- so the normalisation is
// folded into the weights instead:
//
//   n_c = a_c x_c + b_c      x_c = (byte - 128) / 256 (rgb), the raw value otherwise
//   conv(n, W)[co, oh, ow] = sum_{taps inside the image} (a_c W)[co, c, r, s] x_c  +  sum_{taps inside} W[co, c, r, s] b_c
//
// The first term is the stem kernel on the zero-padded exact tensor with weights W' = a_c W (value + residual planes: two
// products instead of three, no x_lo plane); the second depends only on WHICH taps fall inside the image, i.e. on the
// border class of (oh, ow): five classes per axis (first two / interior / last two outputs), a [5][5][32] table added in the
// epilogue.  The top-down channels are fp32 ratios: their fp16 value and residual ride in two channels (the residual in
// the two spare channels 30 / 31, with the same weights).
// Backward: dW[co,c,r,s] = a_c G[co,c,r,s] + b_c D[co,r,s], G = the stem weight-gradient kernel on the exact tensor,
// D = sum of dy over the outputs whose tap (r,s) lies inside the image = a sum of per-border-class sums S[5][5][co].
#include "common.cuh"
#include "elem.cuh"

namespace pnvo {

// xp (device, 6 x 32 floats): [0] scale / [1] shift that turn the reference-unit channel values (rgb/255, depth, one-hot,
// top-down) into the exact stored values, [2] a_c, [3] b_c, [4] source channel of the weights (as float)
__global__ void stem_exact_prep_kernel(const float* __restrict__ scale, const float* __restrict__ shift, int use_rgb,
                                       int use_depth, int n_dd, int use_td, float* __restrict__ xp) {
  const int c = threadIdx.x;
  if (c >= 32) return;
  const int cf = 3 * use_rgb + use_depth + n_dd + use_td;
  const int C = 2 * cf, n_lo = use_td ? 2 : 0;
  float xs = 0.f, xh = 0.f, a = 0.f, b = 0.f, src = 0.f;
  if (c < C) {
    const float sc = scale ? scale[c] : 1.f, sh = shift ? shift[c] : 0.f;
    const bool rgb = use_rgb && (c % cf) < 3;
    src = static_cast<float>(c);
    if (rgb) {
      // stored value (byte - M) / 256: exact in fp16, |x| <= 1.  (Storing byte - M itself would make a_c = 1 / (255 std)
      // and push the RESIDUAL plane of W' = a_c W into fp16 subnormals: measured 1.5e-4 instead of 8e-6 at the output.)
      // M is a CONSTANT (not round(255 mean)): the stored tensor then does not depend on the batch statistics, so the
      // statistics pass and the assembly are one kernel (raw_input.cu: raw_assemble with stats); the remainder
      // b = (128 / 255 - mean) / std is absorbed by the border-class bias
      const float M = 128.f;
      xs = 255.f / 256.f; xh = -M / 256.f;
      a = sc * (256.f / 255.f);
      b = sh + M * sc / 255.f;
    } else {
      xs = 1.f; xh = 0.f; a = sc; b = sh;
    }
  } else if (c < C + n_lo) {
    const int td = (c - C) * cf + cf - 1;          // top-down channel of frame c - C
    xs = 1.f; xh = 0.f;
    a = scale ? scale[td] : 1.f;
    b = 0.f;
    src = static_cast<float>(td);
  }
  xp[c] = xs; xp[32 + c] = xh; xp[64 + c] = a; xp[96 + c] = b; xp[128 + c] = src; xp[160 + c] = 0.f;
}

__device__ __forceinline__ void tap_range(int o, int I, int& lo, int& hi) {  // taps of output o that fall inside [0, I)
  lo = max(0, 3 - 2 * o);
  hi = min(6, I + 2 - 2 * o);
}
__device__ __forceinline__ int class_output(int k, int O) { return k < 2 ? k : (k == 2 ? 2 : O - 5 + k); }
__device__ __forceinline__ int output_class(int o, int O) { return o < 2 ? o : (o <= O - 3 ? 2 : 3 + o - (O - 2)); }

// W' = a_c W[:, src_c] in the stem kernel's layout ([tap pair][descending filter rows by parity][cout][64]), value and
// residual planes
__global__ void stem_exact_pack_kernel(const float* __restrict__ w, int Cin, const float* __restrict__ xp,
                                       __half* __restrict__ wr, __half* __restrict__ wr_lo) {
  const int total = 32 * 32 * 49;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int s = i % 7, r = (i / 7) % 7, c = (i / 49) % 32, n = i / (49 * 32);
    const int pos = (r & 1) ? 4 + (5 - r) / 2 : (6 - r) / 2;
    const float a = xp[64 + c];
    const int src = static_cast<int>(xp[128 + c]);
    const float x = (a != 0.f && src < Cin) ? a * w[((n * Cin + src) * 7 + r) * 7 + s] : 0.f;
    const __half h = __float2half_rn(x);
    const int o = (((s >> 1) * 7 + pos) * 32 + n) * 64 + (s & 1) * 32 + c;
    wr[o] = h;
    wr_lo[o] = __float2half_rn(x - __half2float(h));
  }
}

// bias5[rc][sc][n] = sum over the taps (r, s) valid for row class rc / column class sc of sum_c W[n,c,r,s] b_c
__global__ void __launch_bounds__(64) stem_exact_bias_kernel(const float* __restrict__ w, int Cin, const float* __restrict__ xp,
                                                             int IH, int IW, float* __restrict__ bias5) {
  __shared__ double s_wb[49];
  const int n = blockIdx.x;
  const int OH = (IH - 1) / 2 + 1, OW = (IW - 1) / 2 + 1;
  if (threadIdx.x < 49) {
    double t = 0.0;
    for (int c = 0; c < Cin; ++c) t += static_cast<double>(w[(n * Cin + c) * 49 + threadIdx.x]) * static_cast<double>(xp[96 + c]);
    s_wb[threadIdx.x] = t;
  }
  __syncthreads();
  if (threadIdx.x < 25) {
    const int rc = threadIdx.x / 5, sc = threadIdx.x % 5;
    int rl, rh, sl, sh;
    tap_range(class_output(rc, OH), IH, rl, rh);
    tap_range(class_output(sc, OW), IW, sl, sh);
    double t = 0.0;
    for (int r = rl; r <= rh; ++r)
      for (int s = sl; s <= sh; ++s) t += s_wb[r * 7 + s];
    bias5[(rc * 5 + sc) * 32 + n] = static_cast<float>(t);
  }
}

// S[rc][sc][n] += sum of dy over the outputs of row class rc / column class sc (all samples); dy [B, OH, OW, 32] fp16.
// One block per (output row, 8 samples); a thread owns an 8-channel chunk (one 16-byte load per pixel) of every 64th pixel.
static constexpr int kDySumImgs = 8;
__global__ void __launch_bounds__(256) stem_dy_sums_kernel(const __half* __restrict__ dy, int B, int OH, int OW,
                                                           float* __restrict__ S) {
  __shared__ float s_acc[8][5][32];
  const int oh = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = threadIdx.x & 3, pl = threadIdx.x >> 2;
  const int b0 = blockIdx.y * kDySumImgs, b1 = min(B, b0 + kDySumImgs);
  float acc[5][8];
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;
  for (int b = b0; b < b1; ++b) {
    const uint4* row = reinterpret_cast<const uint4*>(dy + (static_cast<int64_t>(b) * OH + oh) * OW * 32) + q;
    for (int ow = pl; ow < OW; ow += 64) {
      const uint4 u = __ldg(row + static_cast<int64_t>(ow) * 4);
      const __half2* h2 = reinterpret_cast<const __half2*>(&u);
      const int sc = output_class(ow, OW);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h2[e]);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          acc[k][2 * e] += (k == sc) ? f.x : 0.f;
          acc[k][2 * e + 1] += (k == sc) ? f.y : 0.f;
        }
      }
    }
  }
  // lanes with the same channel chunk (lane & 3) hold different pixels: fold them, then the 8 warps through shared memory
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = acc[k][e];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 4) s_acc[warp][k][lane * 8 + e] = v;
    }
  __syncthreads();
  if (threadIdx.x < 160) {
    const int k = threadIdx.x >> 5;
    float t = 0.f;
    for (int w8 = 0; w8 < 8; ++w8) t += s_acc[w8][k][lane];
    atomicAdd(S + (output_class(oh, OH) * 5 + k) * 32 + lane, t);
  }
}

// grad[n][c][r][s] (OIHW fp32) = a_c (G[n][(r,s), c] (+ G of the residual channel for a top-down channel)) + b_c D[n][r][s]
__global__ void __launch_bounds__(256) stem_exact_unpack_kernel(const float* __restrict__ dwp, int w_ld,
                                                                const float* __restrict__ S, const float* __restrict__ xp,
                                                                int Cin, int IH, int IW, float* __restrict__ grad) {
  __shared__ float s_D[49];
  const int n = blockIdx.x;
  const int OH = (IH - 1) / 2 + 1, OW = (IW - 1) / 2 + 1;
  if (threadIdx.x < 49) {
    const int r = threadIdx.x / 7, s = threadIdx.x % 7;
    float t = 0.f;
    for (int rc = 0; rc < 5; ++rc) {
      int rl, rh;
      tap_range(class_output(rc, OH), IH, rl, rh);
      if (r < rl || r > rh) continue;
      for (int sc = 0; sc < 5; ++sc) {
        int sl, sh;
        tap_range(class_output(sc, OW), IW, sl, sh);
        if (s < sl || s > sh) continue;
        t += S[(rc * 5 + sc) * 32 + n];
      }
    }
    s_D[threadIdx.x] = t;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cin * 49; i += blockDim.x) {
    const int c = i / 49, rs = i % 49;
    float g = dwp[static_cast<int64_t>(n) * w_ld + rs * 32 + c];
    for (int cl = Cin; cl < 32; ++cl)  // residual channels that share this channel's weights
      if (xp[64 + cl] != 0.f && static_cast<int>(xp[128 + cl]) == c) g += dwp[static_cast<int64_t>(n) * w_ld + rs * 32 + cl];
    grad[(static_cast<int64_t>(n) * Cin + c) * 49 + rs] = xp[64 + c] * g + xp[96 + c] * s_D[rs];
  }
}

int stem_exact_op(int code, const int32_t* i, const float* f, void* const* p, cudaStream_t st) {
  (void)f;
  switch (code) {
    case PNVO_OP_STEM_EXACT_PREP:
      // p0 = scale [C] (nullable), p1 = shift [C] (nullable), p2 = xp [6][32]; i0 = use_rgb, i1 = use_depth, i2 = n_dd, i3 = use_td
      PNVO_REQUIRE(p[2] && 2 * (3 * i[0] + i[1] + i[2] + i[3]) + (i[3] ? 2 : 0) <= 32, "stem_exact_prep: bad channel layout");
      stem_exact_prep_kernel<<<1, 32, 0, st>>>(static_cast<const float*>(p[0]), static_cast<const float*>(p[1]), i[0], i[1],
                                              i[2], i[3], static_cast<float*>(p[2]));
      count_launch();
      return check_launch("stem_exact_prep");
    case PNVO_OP_STEM_EXACT_PACK:
      // p0 = w OIHW fp32 [32][Cin][7][7], p1 = xp, p2 = packed value plane, p3 = packed residual plane, p4 = bias5 [5][5][32]
      // i0 = Cin, i1 = IH, i2 = IW
      PNVO_REQUIRE(p[0] && p[1] && p[2] && p[3] && p[4] && i[0] <= 32 && i[1] >= 7 && i[2] >= 7, "stem_exact_pack: bad arguments");
      stem_exact_pack_kernel<<<ceil_div(32 * 32 * 49, 256), 256, 0, st>>>(static_cast<const float*>(p[0]), i[0],
                                                                          static_cast<const float*>(p[1]),
                                                                          static_cast<__half*>(p[2]), static_cast<__half*>(p[3]));
      stem_exact_bias_kernel<<<32, 64, 0, st>>>(static_cast<const float*>(p[0]), i[0], static_cast<const float*>(p[1]), i[1],
                                                i[2], static_cast<float*>(p[4]));
      count_launch(2);
      return check_launch("stem_exact_pack");
    case PNVO_OP_STEM_DY_SUMS: {
      // p0 = dy [B, OH, OW, 32] fp16, p1 = S [5][5][32] fp32 (pre-zeroed, accumulated); i0 = B, i1 = OH, i2 = OW
      PNVO_REQUIRE(p[0] && p[1] && i[1] >= 4 && i[2] >= 4, "stem_dy_sums: bad arguments");
      if (i[0] <= 0) return 0;
      stem_dy_sums_kernel<<<dim3(i[1], ceil_div(i[0], kDySumImgs)), 256, 0, st>>>(static_cast<const __half*>(p[0]), i[0], i[1],
                                                                                 i[2], static_cast<float*>(p[1]));
      count_launch();
      return check_launch("stem_dy_sums");
    }
    case PNVO_OP_STEM_EXACT_UNPACK:
      // p0 = packed fp32 dW' [32][w_ld], p1 = S, p2 = xp, p3 = grad OIHW fp32 [32][Cin][7][7]; i0 = w_ld, i1 = Cin, i2 = IH, i3 = IW
      PNVO_REQUIRE(p[0] && p[1] && p[2] && p[3] && i[0] >= 49 * 32 && i[1] <= 32, "stem_exact_unpack: bad arguments");
      stem_exact_unpack_kernel<<<32, 256, 0, st>>>(static_cast<const float*>(p[0]), i[0], static_cast<const float*>(p[1]),
                                                   static_cast<const float*>(p[2]), i[1], i[2], i[3], static_cast<float*>(p[3]));
      count_launch();
      return check_launch("stem_exact_unpack");
    default:
      set_error("stem_exact_op: bad opcode %d", code);
      return -3;
  }
}

}  // namespace pnvo
