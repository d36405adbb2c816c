// Stem convolution (7x7 / stride 2 / pad 3, Cin_pad = 32, Cout = 32: resnet.py:156-164) without im2col
// expansion.
//
// With NHWC fp16 and 32 channels a pixel is 64 bytes, so a stride-2 step along W is 128 bytes -- exactly the
// row pitch of a SWIZZLE_128B UMMA operand.  An input row staged ONCE in shared memory (as TMA wrote it) is
// therefore already the A operand of every filter column pair: for output pixels ow0..ow0+127 and taps
// (s, s+1), s even, the operand rows are the 128-byte pixel pairs starting at raster pixel 2*ow + s, i.e. the
// SAME bytes viewed through a descriptor whose start address is shifted by s/2 rows.  One CTA computes one
// output row (b, oh): 7 pipeline stages (filter rows r), each = one input row (TMA tiled, zero-filled halo)
// + the [4 pairs][32][64] weight slice of that filter row; 2 tiles x 4 pairs x 4 k-steps tcgen05.mma per stage.
// L2 -> smem traffic drops from 12.25x (im2col) to 3.5x the input.
#include "common.cuh"
#include "ops.cuh"
#include "tmap.cuh"

namespace pnvo {

struct StemArgs {
  void* y;        // [B, OH, OW, 32] fp16
  double* stats;  // [B][G][2]
  int B, IH, IW, OH, OW;
  int G, cpg;
  int stages;
  int xrow_bytes;   // smem bytes reserved per staged input row (multiple of 1024)
  int box_px;       // pixels per TMA box (two boxes per row)
  int ow1;          // first output pixel of the second tile (OW - 128), or -1 when OW <= 128
};

static constexpr int kStemW = 16384;  // weight slice of one filter row: 4 pairs x 32 cout x 128 B

template <int CPG>
__device__ __forceinline__ void group_sums(const float* v, float* out) {
#pragma unroll
  for (int g = 0; g < 32 / CPG; ++g) {
    float a = 0.f, q = 0.f;
#pragma unroll
    for (int c = 0; c < CPG; ++c) {
      const float x = v[g * CPG + c];
      a += x;
      q = fmaf(x, x, q);
    }
    out[2 * g] = a;
    out[2 * g + 1] = q;
  }
}

__global__ void __launch_bounds__(192) conv_stem_fwd_kernel(const StemArgs p, const __grid_constant__ ConvTmaps tm) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_full[8];
  __shared__ __align__(8) uint64_t s_empty[8];
  __shared__ __align__(8) uint64_t s_accum;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stages = p.stages;
  const int b = blockIdx.x / p.OH, oh = blockIdx.x - b * p.OH;
  const uint32_t stage_bytes = static_cast<uint32_t>(p.xrow_bytes) + kStemW;
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;
  const int n_tiles = p.ow1 >= 0 ? 2 : 1;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&s_full[s]), 1);
      mbar_init(smem_u32(&s_empty[s]), 1);
    }
    mbar_init(smem_u32(&s_accum), 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(&s_tmem), 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;

  if (warp == 5) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      tma_prefetch_desc(&tm.a);
      tma_prefetch_desc(&tm.b);
      const uint32_t box_bytes = static_cast<uint32_t>(p.box_px) * 128;  // box_px = pixel PAIRS per row
      for (int r = 0; r < 7; ++r) {
        const int s = r % stages;
        if (r >= stages) mbar_wait(smem_u32(&s_empty[s]), ((r / stages) & 1) ^ 1);
        const uint32_t bar = smem_u32(&s_full[s]);
        const uint32_t sX = smem_base + s * stage_bytes;
        const int ih = 2 * oh - 3 + r;
        mbar_arrive_expect_tx(bar, box_bytes + kStemW);
        tma_load_4d(sX, &tm.a, bar, 0, 0, ih, b);  // whole W-padded input row as 128-byte pixel pairs (rows ih < 0 / >= IH: zeros)
        tma_load_2d(sX + p.xrow_bytes, &tm.b, bar, 0, r * 128);       // [4 pairs x 32 cout][64] of filter row r
      }
    }
  } else if (warp == 4) {
    // ================================ MMA issuer ================================
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16(128, 32, 0, 0);
      const uint64_t d0 = umma_desc(0, 16, 1024, 128);
      const uint32_t hi = static_cast<uint32_t>(d0 >> 32), lo0 = static_cast<uint32_t>(d0);
      for (int r = 0; r < 7; ++r) {
        const int s = r % stages;
        mbar_wait(smem_u32(&s_full[s]), (r / stages) & 1);
        tc_fence_after();
        const uint32_t sX = smem_base + s * stage_bytes;
        const uint32_t x_lo = lo0 + (sX >> 4), w_lo = lo0 + ((sX + p.xrow_bytes) >> 4);
        for (int t = 0; t < n_tiles; ++t) {
          const int ow0 = t == 0 ? 0 : p.ow1;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // taps (2j, 2j+1): operand row m = raster pixels 2*(ow0+m)+2j, +1 -> start shifted by (ow0 + j) rows
#pragma unroll
            for (int q = 0; q < 4; ++q)
              tc_mma_f16_lohi(tmem_base + t * 32, x_lo + static_cast<uint32_t>(ow0 + j) * 8 + q * 2, w_lo + j * 256 + q * 2, hi,
                              idesc, (r | j | q) != 0 ? 1u : 0u);
          }
        }
        tc_commit(smem_u32(&s_empty[s]));
      }
      tc_commit(smem_u32(&s_accum));
    }
    __syncwarp();
    tc_fence_before();
  } else {
    // ================================ epilogue (warps 0-3) ================================
    mbar_wait(smem_u32(&s_accum), 0);
    tc_fence_after();
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    for (int t = 0; t < n_tiles; ++t) {
      float v[32];
      tmem_ld32(t_row + t * 32, v);
      tmem_ld_wait();
      const int ow = (t == 0 ? 0 : p.ow1) + tid;
      // tile 1 re-computes pixels [ow1, 128): only its rows >= 128 are new
      const bool valid = (t == 0) ? (ow < p.OW) : (ow >= 128 && ow < p.OW);
      if (p.stats) {
        // all rows of this CTA belong to sample b: plain butterfly over the warp, zeros for masked rows
        float sgrp[32];
        const int ng = 32 / p.cpg;  // cpg in {2,4,8,16,32}
#pragma unroll
        for (int i = 0; i < 32; ++i) sgrp[i] = 0.f;
        if (valid) {
          switch (p.cpg) {
            case 2: group_sums<2>(v, sgrp); break;
            case 4: group_sums<4>(v, sgrp); break;
            case 8: group_sums<8>(v, sgrp); break;
            case 16: group_sums<16>(v, sgrp); break;
            default: group_sums<32>(v, sgrp); break;
          }
        }
        // reduce-scatter over the 32 lanes: lane i ends with the total of value i
        int off = 16;
#pragma unroll
        for (int cnt = 16; cnt >= 1; cnt >>= 1) {
          const bool up = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < cnt; ++i) {
            const float send = up ? sgrp[i] : sgrp[i + cnt];
            const float recv = __shfl_xor_sync(0xffffffffu, send, off);
            sgrp[i] = (up ? sgrp[i + cnt] : sgrp[i]) + recv;
          }
          off >>= 1;
        }
        if (lane < 2 * ng) atomicAdd(p.stats + static_cast<int64_t>(b) * p.G * 2 + lane, static_cast<double>(sgrp[0]));
      }
      if (valid) {
        __half* yp = reinterpret_cast<__half*>(p.y) + (static_cast<int64_t>(b) * p.OH * p.OW + static_cast<int64_t>(oh) * p.OW + ow) * 32;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(v[q * 8 + 2 * e], v[q * 8 + 2 * e + 1]);
          *reinterpret_cast<uint4*>(yp + q * 8) = u;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

// OIHW fp32 [32][Cin][7][7] -> [r][pair j][cout][64] fp16, column = (s & 1) * 32 + c, tap s = 2j + (s & 1); the
// 8th tap (j = 3, odd half) and channels >= Cin stay zero (buffer zero-initialised once).
__global__ void pack_w_stem_kernel(const float* __restrict__ w, int Cin, __half* __restrict__ wr) {
  const int total = 32 * Cin * 49;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int s = i % 7, r = (i / 7) % 7, c = (i / 49) % Cin, n = i / (49 * Cin);
    wr[((r * 4 + (s >> 1)) * 32 + n) * 64 + (s & 1) * 32 + c] = __float2half_rn(w[i]);
  }
}

int pack_w_stem_launch(const float* w, int Cin, __half* wr, cudaStream_t st) {
  PNVO_REQUIRE(w && wr && Cin <= 32, "pack_w_stem: bad arguments");
  pack_w_stem_kernel<<<ceil_div(32 * Cin * 49, 256), 256, 0, st>>>(w, Cin, wr);
  count_launch();
  return check_launch("pack_w_stem");
}

// x: W-padded input [B, IH, Wp, 32] fp16 with Wp = stem_padded_width(IW): 3 zero pixels on the left (so that raster
// pixel p' = iw + 3 and tap pairs start on even p'), zeros on the right up to an even pixel count.
int stem_padded_width(int IW) {
  const int OW = (IW + 6 - 7) / 2 + 1;
  const int npx = 2 * (std::max(OW, 128) - 1) + 8;
  return std::max((IW + 3 + 1) & ~1, (npx + 1) & ~1);
}

int conv_stem_fwd_launch(const __half* x, const __half* wr, void* y, double* stats, int B, int IH, int IW, int G, int cpg,
                         int stages, cudaStream_t st) {
  PNVO_REQUIRE(x && wr && y, "conv_stem: null pointer");
  StemArgs a{};
  a.y = y; a.stats = stats; a.B = B; a.IH = IH; a.IW = IW;
  a.OH = (IH + 6 - 7) / 2 + 1;
  a.OW = (IW + 6 - 7) / 2 + 1;
  PNVO_REQUIRE(a.OW <= 256, "conv_stem: output width %d > 256", a.OW);
  PNVO_REQUIRE(!stats || (cpg >= 2 && cpg <= 32 && (cpg & (cpg - 1)) == 0 && G * cpg == 32), "conv_stem: bad group config");
  a.G = G; a.cpg = cpg;
  a.ow1 = a.OW > 128 ? a.OW - 128 : -1;
  const int Wp = stem_padded_width(IW);
  a.box_px = Wp / 2;                             // pixel pairs per staged row (one TMA box)
  PNVO_REQUIRE(a.box_px <= 256, "conv_stem: row too wide for one TMA box");
  a.xrow_bytes = (a.box_px * 128 + 1023) & ~1023;
  a.stages = std::max(2, std::min(stages, 7));
  const int smem_bytes = a.stages * (a.xrow_bytes + kStemW) + 1024;
  PNVO_REQUIRE(smem_bytes <= 220 * 1024, "conv_stem: %d bytes of shared memory", smem_bytes);
  alignas(64) ConvTmaps tm;
  memset(&tm, 0, sizeof(tm));
  // pixel pairs as the innermost 128-byte dimension: [B, IH, Wp/2, 64]
  if (tmap_tiled4d(&tm.a, x, B, IH, Wp / 2, 64, a.box_px)) return -1;
  if (tmap_tiled2d(&tm.b, wr, 7 * 4 * 32, 64, 64, 128, 64)) return -1;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(conv_stem_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    attr = true;
  }
  conv_stem_fwd_kernel<<<B * a.OH, 192, smem_bytes, st>>>(a, tm);
  count_launch();
  return check_launch("conv_stem_fwd");
}

// ---------------------------------------------------------------------------------------------------------
// Stem weight gradient on the same raster: dW[n, (r, s, c)] = sum_{b,oh,ow} dy[b,oh,ow,n] * x[b, 2oh-3+r, 2ow-3+s, c].
// GEMM-K = output pixels along a row.  A (MN-major, M = 128 = two adjacent tap pairs x 64) is again a shifted
// view of the staged input row: block 0 starts at pixel pair (ow + 2h), block 1 one 128-byte row later
// (LBO = 128 B).  B = the dy row segment (MN-major, SWIZZLE_64B, N = 32).  A CTA owns (sample, half row, range of
// output rows): input rows live in a 9-slot ring (each is used by 3-4 consecutive output rows, two new rows per
// step), all 14 = 7 r x 2 accumulators stay in TMEM (448 columns) for the whole range, one atomic flush at the end.
// ---------------------------------------------------------------------------------------------------------
struct StemWgradArgs {
  float* dw;  // packed fp32 [32][w_ld], kflat = (r*7 + s)*32 + c
  int w_ld;
  int B, IH, OH, OW;
  int rows_per_cta, n_chunks;
};
static constexpr int kSwRow = 13312;   // 104 pixel pairs x 128 B
static constexpr int kSwDy = 6144;     // 96 pixels x 64 B
static constexpr int kSwHalf0 = 96;    // output pixels [0, 96) and [96, 176)

__global__ void __launch_bounds__(192) conv_stem_wgrad_kernel(const StemWgradArgs p,
                                                              const __grid_constant__ ConvTmaps tm) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_full[2];
  __shared__ __align__(8) uint64_t s_empty[2];
  __shared__ __align__(8) uint64_t s_accum;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t s_dy = smem_base + 9 * kSwRow;
  int bid = blockIdx.x;
  const int chunk = bid % p.n_chunks; bid /= p.n_chunks;
  const int half = bid & 1;
  const int b = bid >> 1;
  const int oh_begin = chunk * p.rows_per_cta;
  const int oh_end = min(p.OH, oh_begin + p.rows_per_cta);
  const int n_rows = oh_end - oh_begin;
  const int ow_base = half ? kSwHalf0 : 0;
  const int ksteps = half ? (ceil_div(p.OW - kSwHalf0, 16)) : (kSwHalf0 / 16);

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&s_full[s]), 1);
      mbar_init(smem_u32(&s_empty[s]), 1);
    }
    mbar_init(smem_u32(&s_accum), 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(&s_tmem), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;

  if (warp == 5) {
    if (elect_one()) {
      // ================================ TMA producer ================================
      tma_prefetch_desc(&tm.a);
      tma_prefetch_desc(&tm.b);
      for (int t = 0; t < n_rows; ++t) {
        const int oh = oh_begin + t, pslot = t & 1;
        if (t >= 2) mbar_wait(smem_u32(&s_empty[pslot]), ((t >> 1) & 1) ^ 1);
        const uint32_t bar = smem_u32(&s_full[pslot]);
        const int r_first = (t == 0) ? 0 : 5;  // first step: all 7 rows of the window; later: the two new ones
        mbar_arrive_expect_tx(bar, static_cast<uint32_t>(7 - r_first) * kSwRow + kSwDy);
        for (int r = r_first; r < 7; ++r) {
          const int ih = 2 * oh - 3 + r;
          const int slot = (2 * oh + r) % 9;  // == (ih + 3) mod 9
          tma_load_4d(smem_base + slot * kSwRow, &tm.a, bar, 0, ow_base, ih, b);
        }
        tma_load_4d(s_dy + pslot * kSwDy, &tm.b, bar, 0, ow_base, oh, b);
      }
    }
  } else if (warp == 4) {
    if (elect_one()) {
      // ================================ MMA issuer ================================
      const uint32_t idesc = umma_idesc_f16(128, 32, 1, 1);
      const uint64_t da0 = umma_desc(0, 128, 1024, 128), db0 = umma_desc(0, 64, 512, 64);
      const uint32_t a_hi = static_cast<uint32_t>(da0 >> 32), a_lo0 = static_cast<uint32_t>(da0);
      const uint32_t b_hi = static_cast<uint32_t>(db0 >> 32), b_lo0 = static_cast<uint32_t>(db0);
      for (int t = 0; t < n_rows; ++t) {
        const int oh = oh_begin + t, pslot = t & 1;
        mbar_wait(smem_u32(&s_full[pslot]), (t >> 1) & 1);
        tc_fence_after();
        const uint32_t b_lo = b_lo0 + ((s_dy + pslot * kSwDy) >> 4);
        for (int r = 0; r < 7; ++r) {
          const uint32_t x_lo = a_lo0 + ((smem_base + static_cast<uint32_t>((2 * oh + r) % 9) * kSwRow) >> 4);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            for (int q = 0; q < ksteps; ++q) {
              // K step q = output pixels ow_base + 16q .. +15 -> staged pair rows (16q + 2h) .. ; block 1 one row later
              tc_mma_f16_parts(tmem_base + static_cast<uint32_t>((r * 2 + h) * 32), x_lo + static_cast<uint32_t>(16 * q + 2 * h) * 8,
                               a_hi, b_lo + static_cast<uint32_t>(q) * 64, b_hi, idesc, (t | q) != 0 ? 1u : 0u);
            }
          }
        }
        tc_commit(smem_u32(&s_empty[pslot]));
      }
      tc_commit(smem_u32(&s_accum));
    }
    __syncwarp();
    tc_fence_before();
  } else {
    // ================================ epilogue ================================
    mbar_wait(smem_u32(&s_accum), 0);
    tc_fence_after();
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const int jj = tid >> 6, within = tid & 63;
    for (int r = 0; r < 7; ++r) {
      for (int h = 0; h < 2; ++h) {
        float v[32];
        tmem_ld32(t_row + (r * 2 + h) * 32, v);
        tmem_ld_wait();
        const int s = 2 * (2 * h + jj) + (within >> 5);
        if (s < 7) {
          const int kf = (r * 7 + s) * 32 + (within & 31);
#pragma unroll
          for (int n = 0; n < 32; ++n) atomicAdd(p.dw + static_cast<int64_t>(n) * p.w_ld + kf, v[n]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int conv_stem_wgrad_launch(const __half* x, const __half* dy, float* dw, int w_ld, int B, int IH, int IW,
                           int rows_per_cta, cudaStream_t st) {
  PNVO_REQUIRE(x && dy && dw, "conv_stem_wgrad: null pointer");
  StemWgradArgs a{};
  a.dw = dw; a.w_ld = w_ld; a.B = B; a.IH = IH;
  a.OH = (IH + 6 - 7) / 2 + 1;
  a.OW = (IW + 6 - 7) / 2 + 1;
  PNVO_REQUIRE(a.OW > kSwHalf0 && a.OW <= 176, "conv_stem_wgrad: output width %d not in (96, 176]", a.OW);
  PNVO_REQUIRE(w_ld >= 49 * 32, "conv_stem_wgrad: w_ld too small");
  a.rows_per_cta = std::max(4, rows_per_cta);
  a.n_chunks = ceil_div(a.OH, a.rows_per_cta);
  const int Wp = stem_padded_width(IW);
  alignas(64) ConvTmaps tm;
  memset(&tm, 0, sizeof(tm));
  if (tmap_tiled4d(&tm.a, x, B, IH, Wp / 2, 64, kSwRow / 128)) return -1;   // 104 pixel pairs per box
  if (tmap_tiled4d(&tm.b, dy, B, a.OH, a.OW, 32, kSwHalf0, 64)) return -1;   // 96 pixels x 32 channels, SWIZZLE_64B
  const int smem_bytes = 9 * kSwRow + 2 * kSwDy + 1024;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(conv_stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    attr = true;
  }
  conv_stem_wgrad_kernel<<<B * 2 * a.n_chunks, 192, smem_bytes, st>>>(a, tm);
  count_launch();
  return check_launch("conv_stem_wgrad");
}

}  // namespace pnvo

// ---------------------------------------------------------------------------------------------------------
// debug / measurement: issue rate of tcgen05.mma (M = 128, K = 16, fp16) as a function of N, the swizzle mode and the
// byte shift of the A descriptor's start address relative to the swizzle atom (the raster kernels shift it by whole
// pixels).  One CTA per SM issues `n_mma` MMAs back to back on fixed shared-memory operands and reports cycles / MMA.
// ---------------------------------------------------------------------------------------------------------
namespace pnvo {
__global__ void __launch_bounds__(128) mma_rate_kernel(int N, int row_bytes, int a_shift, int a_step, int mn_major,
                                                       int n_mma, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  for (uint32_t off = tid * 16; off < 96 * 1024; off += blockDim.x * 16)
    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(base + off), "r"(0x3c003c00u) : "memory");
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(smem_u32(&s_bar), 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&s_tmem), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16(128, N, mn_major, mn_major);
      const uint32_t sA = base + a_shift, sB = base + 64 * 1024;
      const uint32_t lbo = mn_major ? row_bytes : 16;
      const long long t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        const uint64_t adesc = umma_desc(sA + static_cast<uint32_t>(i & 7) * a_step, lbo, 8 * row_bytes, row_bytes);
        const uint64_t bdesc = umma_desc(sB, lbo, 8 * row_bytes, row_bytes);
        tc_mma_f16(tmem_base + ((i & 1) ? 256 : 0), adesc, bdesc, idesc, i > 1 ? 1u : 0u);
      }
      tc_commit(smem_u32(&s_bar));
      mbar_wait(smem_u32(&s_bar), 0);
      const long long t1 = clock64();
      out[blockIdx.x] = t1 - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}
}  // namespace pnvo

extern "C" int pnvo_debug_mma_rate(int N, int row_bytes, int a_shift, int a_step, int mn_major, int n_mma, int n_ctas,
                                   void* out_cycles, void* stream) {
  using namespace pnvo;
  PNVO_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && (row_bytes == 64 || row_bytes == 128) && out_cycles, "mma_rate: bad args");
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  mma_rate_kernel<<<n_ctas, 128, 97 * 1024, static_cast<cudaStream_t>(stream)>>>(N, row_bytes, a_shift, a_step, mn_major,
                                                                                  n_mma, static_cast<long long*>(out_cycles));
  return check_launch("mma_rate");
}
