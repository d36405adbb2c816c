// Stem convolution (7x7 / stride 2 / pad 3, Cin_pad = 32, Cout = 32: resnet.py:156-164), second formulation:
// output PIXELS are the UMMA N dimension and (4 output rows x 32 channels) the M dimension.
//
// conv_stem.cu computes D[pixel, cout] with N = 32: every tcgen05.mma re-reads a 4 KB pixel tile from shared memory
// for 128 x 32 x 16 MACs and runs at 40 clk (measured, tools/mma_rate.py: 32 + N/4 clk for N <= 128, N/2 above), i.e.
// 40 % of the tensor-pipe rate, and the 112 KB of stem weights are re-streamed from L2 for every output row.  Here
//
//   D[(q, cout), ow] += W_r(q)[cout, (s, c)] * x[h, 2 ow + s, c]        q = 0..3 <-> output row oh0 + q, r(q) = h + 3 - 2 (oh0 + q)
//
// for every input row h that touches the group of 4 output rows (13 rows): the B operand is the staged input row
// seen as 128-byte pixel pairs (tap pairs = descriptor shifts by one pair, as before), N = 176 >= OW columns at the
// full MMA rate (88 clk for 128 x 176 x 16); the A operand is a window of 4 consecutive entries of the filter rows
// stored in descending order of the same parity ([6,4,2,0] / [5,3,1]), so the 4 blocks of M are contiguous; blocks
// whose tap falls outside 0..6 are switched off with tcgen05.mma's disable_output_lane mask.  Persistent CTA: all
// weights resident (112 KB), input rows streamed ONCE per group through a TMA ring, accumulator double-buffered in
// TMEM (2 x 176 columns) so the epilogue of group g overlaps the MMAs of group g+1.
// Epilogue: thread = (q, cout), 32 lanes of a warp = the 32 channels of one pixel -> 64-byte coalesced fp16 stores;
// GroupNorm partial sums per thread over the row, folded across the group's lanes by shuffles, one atomic per value.
#include "common.cuh"
#include "ops.cuh"
#include "tmap.cuh"

namespace pnvo {

struct Stem2Args {
  void* y;        // [B, OH, OW, 32] fp16 (fp32 when out_fp32)
  __half* y_lo;   // optional residual plane of the output (y then holds fp16(acc), y_lo fp16(acc - y))
  const __half* add;  // optional [B, OH, OW, 32] fp16 added before the store (split mode: the w_lo * x product)
  const float* bias5; // optional [5][5][32] fp32: bias per (row class, column class, cout) of the exact-input stem (stem_exact.cu)
  int out_fp32;
  int x_planes;   // 1, or 2 in split mode: every input row is staged twice (x through tm.a, x_lo through tm.a_lo)
  double* stats;  // [B][G][2]
  int B, IH, OH, OW;
  int G, cpg;
  int n_cols;         // UMMA N: OW rounded up to 16
  int xrow_bytes;     // shared-memory bytes of one staged input row (multiple of 1024)
  int stages;         // ring depth
  int groups_per_img, n_groups;
};

static constexpr int kS2Plane = 7 * 32 * 128;   // one tap-pair plane: 7 filter rows x 32 cout x 128 B
static constexpr int kS2W = 4 * kS2Plane;       // 114688 B

__device__ __forceinline__ void tc_mma_f16_masked(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                                  uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, {%6, %7, %8, %9}, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate), "r"(m0), "r"(m1), "r"(m2), "r"(m3)
      : "memory");
}

// order in which the 13 input rows of a group are processed: d = 6 first (all four blocks valid -> it initialises
// every accumulator lane with accumulate = 0), then the rest
__device__ __forceinline__ int stem2_row(int k) { return k == 0 ? 6 : (k <= 6 ? k - 1 : k); }

__global__ void __launch_bounds__(192) conv_stem2_fwd_kernel(const Stem2Args p, const __grid_constant__ ConvTmaps tm) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_wfull;
  __shared__ __align__(8) uint64_t s_xfull[4];
  __shared__ __align__(8) uint64_t s_xempty[4];
  __shared__ __align__(8) uint64_t s_accfull[2];
  __shared__ __align__(8) uint64_t s_accempty[2];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;
  // [x stage 0][weights][x stages 1..]: weight windows that start before the first plane / run past the last one
  // (their lanes are masked) still read inside the allocation
  const uint32_t sW = smem_base + p.xrow_bytes;
  const int stages = p.stages;
  auto stage_addr = [&](int s) -> uint32_t {
    return s == 0 ? smem_base : sW + kS2W + static_cast<uint32_t>(s - 1) * p.xrow_bytes;
  };
  const uint32_t sStage = sW + kS2W + static_cast<uint32_t>(stages - 1) * p.xrow_bytes;  // epilogue tiles: 4 warps x 4 KB

  if (tid == 0) {
    mbar_init(smem_u32(&s_wfull), 1);
    for (int s = 0; s < 4; ++s) {
      mbar_init(smem_u32(&s_xfull[s]), 1);
      mbar_init(smem_u32(&s_xempty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&s_accfull[s]), 1);
      mbar_init(smem_u32(&s_accempty[s]), 4);
    }
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
    // ================================ TMA producer ================================
    if (elect_one()) {
      tma_prefetch_desc(&tm.a);
      tma_prefetch_desc(&tm.b);
      if (p.x_planes > 1) tma_prefetch_desc(&tm.a_lo);
      const uint32_t wbar = smem_u32(&s_wfull);
      mbar_arrive_expect_tx(wbar, kS2W);
      for (int j = 0; j < 4; ++j) tma_load_2d(sW + j * kS2Plane, &tm.b, wbar, 0, j * 224);
      const uint32_t x_tx = static_cast<uint32_t>(p.n_cols + 8) * 128;
      int ctr = 0;
      for (int g = blockIdx.x; g < p.n_groups; g += gridDim.x) {
        const int b = g / p.groups_per_img;
        const int oh0 = (g - b * p.groups_per_img) * 4;
        for (int k = 0; k < 13; ++k) {
          const int d = stem2_row(k);
          const int h = 2 * oh0 - 3 + d;
          if (k != 0 && (h < 0 || h >= p.IH)) continue;  // all-zero row: nothing to add
          for (int pl = 0; pl < p.x_planes; ++pl) {
            const int s = ctr % stages;
            if (ctr >= stages) mbar_wait(smem_u32(&s_xempty[s]), ((ctr / stages) & 1) ^ 1);
            const uint32_t bar = smem_u32(&s_xfull[s]);
            mbar_arrive_expect_tx(bar, x_tx);
            // whole W-padded row as pixel pairs; outside the image: zeros
            tma_load_4d(stage_addr(s), pl ? &tm.a_lo : &tm.a, bar, 0, 0, h, b);
            ++ctr;
          }
        }
      }
    }
  } else if (warp == 4) {
    // ================================ MMA issuer ================================
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16(128, p.n_cols, 0, 0);
      const uint64_t d0 = umma_desc(0, 16, 1024, 128);
      const uint32_t hi = static_cast<uint32_t>(d0 >> 32), lo0 = static_cast<uint32_t>(d0);
      mbar_wait(smem_u32(&s_wfull), 0);
      int ctr = 0, i = 0;
      for (int g = blockIdx.x; g < p.n_groups; g += gridDim.x, ++i) {
        const int b = g / p.groups_per_img;
        const int oh0 = (g - b * p.groups_per_img) * 4;
        const int ab = i & 1;
        if (i >= 2) {
          mbar_wait(smem_u32(&s_accempty[ab]), ((i >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        const uint32_t d_tmem = tmem_base + ab * 256;
        for (int k = 0; k < 13; ++k) {
          const int d = stem2_row(k);
          const int h = 2 * oh0 - 3 + d;
          if (k != 0 && (h < 0 || h >= p.IH)) continue;
          // window of 4 filter rows: even d -> [6,4,2,0] from position (6-d)/2, odd d -> [5,3,1] from (5-d)/2
          const int odd = d & 1;
          const int p0 = odd ? (5 - d) / 2 : (6 - d) / 2;   // exact divisions (numerators are even)
          const int npos = odd ? 3 : 4;
          uint32_t m[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) m[q] = (p0 + q >= 0 && p0 + q < npos) ? 0u : 0xFFFFFFFFu;
          const int a_off = (odd ? 4 * 4096 : 0) + p0 * 4096;  // may be negative: masked lanes, valid addresses
          const uint32_t a_lo = lo0 + (static_cast<uint32_t>(static_cast<int>(sW) + a_off) >> 4);
          for (int pl = 0; pl < p.x_planes; ++pl) {
            const int s = ctr % stages;
            mbar_wait(smem_u32(&s_xfull[s]), (ctr / stages) & 1);
            tc_fence_after();
            const uint32_t b_lo = lo0 + (stage_addr(s) >> 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                tc_mma_f16_masked(d_tmem, a_lo + j * (kS2Plane >> 4) + kk * 2, b_lo + j * 8 + kk * 2, hi, idesc, m[0], m[1],
                                  m[2], m[3], (k | j | kk | pl) != 0 ? 1u : 0u);
            }
            tc_commit(smem_u32(&s_xempty[s]));
            ++ctr;
          }
        }
        tc_commit(smem_u32(&s_accfull[ab]));
      }
    }
    __syncwarp();
    tc_fence_before();
  } else {
    // ================================ epilogue (warps 0-3: warp = output row of the group, lane = channel) ================
    const uint32_t t_lane = static_cast<uint32_t>(warp * 32) << 16;
    int i = 0;
    for (int g = blockIdx.x; g < p.n_groups; g += gridDim.x, ++i) {
      const int b = g / p.groups_per_img;
      const int oh = (g - b * p.groups_per_img) * 4 + warp;
      const int ab = i & 1;
      const bool row_valid = oh < p.OH;
      const int64_t row_base = (static_cast<int64_t>(b) * p.OH + oh) * p.OW * 32;
      float* yrow32 = static_cast<float*>(p.y) + row_base + lane;
      const bool arow = p.add != nullptr;
      // exact-input stem: the normalisation shift folded into a bias that depends on which taps lie inside the image
      float bcol[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      if (p.bias5 && row_valid) {
        const int rc = oh < 2 ? oh : (oh <= p.OH - 3 ? 2 : 3 + oh - (p.OH - 2));
#pragma unroll
        for (int k = 0; k < 5; ++k) bcol[k] = __ldg(p.bias5 + (rc * 5 + k) * 32 + lane);
      }
      float sum = 0.f, ssq = 0.f;
      mbar_wait(smem_u32(&s_accfull[ab]), (i >> 1) & 1);
      tc_fence_after();
      const int n_chunks = (p.n_cols + 31) >> 5;
      // fp16 outputs (and the fp16 tensor added in split mode) pass through a per-warp shared-memory tile [32 pixels][32
      // channels]: the accumulator arrives channel-per-lane (64-byte stores per pixel: 0.85 ms per launch with a residual
      // plane), the tile turns it into 16-byte vectors, 512 contiguous bytes per warp instruction.
      __half* s_hi = reinterpret_cast<__half*>(smem + (sStage - smem_u32(smem))) + warp * 2048;
      __half* s_lo = s_hi + 1024;
      const int px4 = lane >> 2, c4 = lane & 3;
      for (int ch = 0; ch < n_chunks; ++ch) {
        float av[32];
        if (arow && row_valid) {
          // four 16-byte loads per lane (8 pixels x 64 B per instruction), staged, then read back channel-per-lane
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int px = px4 + 8 * k, ow = ch * 32 + px;
            uint4 u = make_uint4(0, 0, 0, 0);
            if (ow < p.OW) u = __ldg(reinterpret_cast<const uint4*>(p.add + row_base + static_cast<int64_t>(ow) * 32) + c4);
            *reinterpret_cast<uint4*>(s_hi + px * 32 + c4 * 8) = u;
          }
          __syncwarp();
#pragma unroll
          for (int e = 0; e < 32; ++e) av[e] = __half2float(s_hi[e * 32 + lane]);
          __syncwarp();
        }
        float v[32];
        tmem_ld32(tmem_base + t_lane + ab * 256 + ch * 32, v);
        tmem_ld_wait();
        if (ch == n_chunks - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&s_accempty[ab]));
        }
        if (row_valid) {
          if (arow) {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] += av[e];
          }
          if (p.bias5) {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const int ow = ch * 32 + e;
              v[e] += ow < 2 ? (ow == 0 ? bcol[0] : bcol[1]) : (ow <= p.OW - 3 ? bcol[2] : (ow == p.OW - 2 ? bcol[3] : bcol[4]));
            }
          }
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            if (ch * 32 + e < p.OW) {
              sum += v[e];
              ssq = fmaf(v[e], v[e], ssq);
            }
          }
        }
        if (p.out_fp32) {
          if (row_valid) {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (ch * 32 + e < p.OW) yrow32[static_cast<int64_t>(ch * 32 + e) * 32] = v[e];
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const __half h = __float2half_rn(v[e]);
            s_hi[e * 32 + lane] = h;
            if (p.y_lo) s_lo[e * 32 + lane] = __float2half_rn(v[e] - __half2float(h));
          }
          __syncwarp();
          if (row_valid) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int px = px4 + 8 * k, ow = ch * 32 + px;
              if (ow < p.OW) {
                const int64_t o = row_base + static_cast<int64_t>(ow) * 32 + c4 * 8;
                *reinterpret_cast<uint4*>(static_cast<__half*>(p.y) + o) = *reinterpret_cast<const uint4*>(s_hi + px * 32 + c4 * 8);
                if (p.y_lo) *reinterpret_cast<uint4*>(p.y_lo + o) = *reinterpret_cast<const uint4*>(s_lo + px * 32 + c4 * 8);
              }
            }
          }
          __syncwarp();
        }
      }
      if (p.stats) {
        for (int o = 1; o < p.cpg; o <<= 1) {
          sum += __shfl_xor_sync(0xffffffffu, sum, o);
          ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
        }
        if (row_valid && (lane % p.cpg) == 0) {
          double* st = p.stats + (static_cast<int64_t>(b) * p.G + lane / p.cpg) * 2;
          atomicAdd(st, static_cast<double>(sum));
          atomicAdd(st + 1, static_cast<double>(ssq));
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

// OIHW fp32 [32][Cin][7][7] -> [pair j][position][cout][64] fp16: positions 0..3 = filter rows 6,4,2,0, positions 4..6 =
// rows 5,3,1; column = (s & 1) * 32 + c with s = 2j + (s & 1); tap s = 7 and channels >= Cin stay zero.
// lo != 0: the residual plane w - fp16(w) of the split-fp16 representation
__global__ void pack_w_stem2_kernel(const float* __restrict__ w, int Cin, __half* __restrict__ wr, int lo) {
  const int total = 32 * Cin * 49;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int s = i % 7, r = (i / 7) % 7, c = (i / 49) % Cin, n = i / (49 * Cin);
    const int pos = (r & 1) ? 4 + (5 - r) / 2 : (6 - r) / 2;
    const float x = w[i];
    const __half h = __float2half_rn(x);
    wr[(((s >> 1) * 7 + pos) * 32 + n) * 64 + (s & 1) * 32 + c] = lo ? __float2half_rn(x - __half2float(h)) : h;
  }
}

int pack_w_stem2_launch(const float* w, int Cin, __half* wr, int lo, cudaStream_t st) {
  PNVO_REQUIRE(w && wr && Cin <= 32, "pack_w_stem2: bad arguments");
  pack_w_stem2_kernel<<<ceil_div(32 * Cin * 49, 256), 256, 0, st>>>(w, Cin, wr, lo);
  count_launch();
  return check_launch("pack_w_stem2");
}

int conv_stem2_supported(int IH, int IW) {
  const int OW = (IW + 6 - 7) / 2 + 1;
  const int n_cols = ceil_div(OW, 16) * 16;
  return (n_cols <= 240 && IH >= 7) ? 1 : 0;
}

int conv_stem2_fwd_launch(const __half* x, const __half* wr, void* y, double* stats, int B, int IH, int IW, int G, int cpg,
                          cudaStream_t st, const __half* x_lo, const __half* add, int out_fp32, const float* bias5,
                          __half* y_lo) {
  PNVO_REQUIRE(x && wr && y, "conv_stem2: null pointer");
  PNVO_REQUIRE(conv_stem2_supported(IH, IW), "conv_stem2: unsupported geometry %dx%d", IH, IW);
  PNVO_REQUIRE(!stats || (cpg >= 1 && cpg <= 32 && (cpg & (cpg - 1)) == 0 && G * cpg == 32), "conv_stem2: bad group config");
  Stem2Args a{};
  a.y = y; a.stats = stats; a.B = B; a.IH = IH;
  a.add = add; a.out_fp32 = out_fp32; a.x_planes = x_lo ? 2 : 1;
  a.bias5 = bias5;
  a.y_lo = y_lo;
  PNVO_REQUIRE(!(y_lo && out_fp32), "conv_stem2: y_lo excludes out_fp32");
  a.OH = (IH + 6 - 7) / 2 + 1;
  a.OW = (IW + 6 - 7) / 2 + 1;
  a.G = G; a.cpg = cpg;
  PNVO_REQUIRE(!bias5 || (a.OH >= 4 && a.OW >= 4), "conv_stem2: the border bias needs >= 4 outputs per axis");
  a.n_cols = ceil_div(a.OW, 16) * 16;
  a.xrow_bytes = ((a.n_cols + 8) * 128 + 1023) & ~1023;
  a.groups_per_img = ceil_div(a.OH, 4);
  a.n_groups = B * a.groups_per_img;
  constexpr int kEpiBytes = 4 * 4096;  // per-warp output staging tiles (value + residual planes)
  a.stages = std::min(4, (231000 - 1024 - kEpiBytes - kS2W) / a.xrow_bytes);
  PNVO_REQUIRE(a.stages >= 2, "conv_stem2: input rows too wide for the shared-memory ring");
  const int smem_bytes = a.stages * a.xrow_bytes + kS2W + 1024 + kEpiBytes;
  const int Wp = stem_padded_width(IW);
  alignas(64) ConvTmaps tm;
  memset(&tm, 0, sizeof(tm));
  // pixel pairs as the innermost 128-byte dimension: [B, IH, Wp/2, 64]; the box may run past Wp/2 (zero fill)
  if (tmap_tiled4d(&tm.a, x, B, IH, Wp / 2, 64, a.n_cols + 8)) return -1;
  if (x_lo && tmap_tiled4d(&tm.a_lo, x_lo, B, IH, Wp / 2, 64, a.n_cols + 8)) return -1;
  if (tmap_tiled2d(&tm.b, wr, 4 * 7 * 32, 64, 64, 224, 64)) return -1;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(conv_stem2_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 231000);
    attr = true;
  }
  if (B <= 0) return 0;
  conv_stem2_fwd_kernel<<<std::min(a.n_groups, 148), 192, smem_bytes, st>>>(a, tm);
  count_launch();
  return check_launch("conv_stem2_fwd");
}

}  // namespace pnvo

// ---------------------------------------------------------------------------------------------------------
// Stem weight gradient, second formulation (same idea as the forward above: make the UMMA N dimension wide).
//
//   dW[n, r, s, c] = sum_{b, oh, ow} dy[b, oh, ow, n] * x[b, 2 oh - 3 + r, 2 ow - 3 + s, c]
//
// GEMM-K = output column ow.  For one input row h the filter rows that use it have the parity of h + 3:
// r = r0 + 2q (r0 = (h+3) & 1, q = 0..3) and belong to the output rows oh = t - q, t = (h + 3 - r0) / 2.  So
//   D_{r0, jh}[(tap pair block, pixel parity, c), (q, n)] += x_row_h[ow + pairs 2jh, 2jh+1]^T  *  [dy_t | dy_{t-1} | dy_{t-2} | dy_{t-3}]
// with A = the staged input row (MN-major, M = 128 = two adjacent tap pairs x 64, blocks one 128-byte pair row apart,
// exactly as conv_stem_wgrad_kernel) and B = a WINDOW OF FOUR dy ROWS (MN-major, N = 128 = 4 blocks x 32 channels,
// blocks one ring slot = LBO apart): N = 128 runs at the full MMA rate (64 clk) instead of 40 clk for N = 32, and
// every staged input row is used for all four filter rows it feeds.  The block -> filter-row map is fixed per
// parity class, so only 2 x 2 accumulators (128 columns each = all 512 TMEM columns) live for the whole kernel.
// dy rows sit in a 6-slot ring stored in DESCENDING row order (slot = -oh mod 6) so a window is 4 ascending slots;
// slots 0..2 are mirrored behind the ring (slots 6..8) so windows never wrap.  Persistent CTA; one flush at the end.
// ---------------------------------------------------------------------------------------------------------
namespace pnvo {

struct StemWg2Args {
  float* dw;
  int w_ld;
  int B, IH, OH, OW;
  int t_per_unit, units_per_img, n_units;
};
static constexpr int kWg2XRow = 23552;          // 184 pixel pairs x 128 B
static constexpr int kWg2DyRow = 11264;         // 176 pixels x 64 B
static constexpr int kWg2XStages = 4;
static constexpr int kWg2DySlots = 9;           // 6-slot ring + 3 mirrors
static constexpr int kWg2KSteps = 11;           // 176 output columns / 16

__global__ void __launch_bounds__(192) conv_stem_wgrad2_kernel(const StemWg2Args p, const __grid_constant__ ConvTmaps tm) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_full[kWg2XStages];
  __shared__ __align__(8) uint64_t s_empty[kWg2XStages];
  __shared__ __align__(8) uint64_t s_accum;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t sDy = smem_base + kWg2XStages * kWg2XRow;
  const int t_last = (p.IH + 2) / 2;  // largest t with an input row inside the image: h = 2t - 3 <= IH - 1

  if (tid == 0) {
    for (int s = 0; s < kWg2XStages; ++s) {
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
      auto load_dy = [&](int oh, int b, uint32_t bar) -> uint32_t {
        const int slot = ((-oh) % 6 + 6) % 6;
        tma_load_4d(sDy + slot * kWg2DyRow, &tm.b, bar, 0, 0, oh, b);   // rows outside [0, OH) / columns >= OW: zeros
        if (slot < 3) {
          tma_load_4d(sDy + (slot + 6) * kWg2DyRow, &tm.b, bar, 0, 0, oh, b);
          return 2u * kWg2DyRow;
        }
        return static_cast<uint32_t>(kWg2DyRow);
      };
      int ctr = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        const int b = u / p.units_per_img;
        const int t0 = 1 + (u - b * p.units_per_img) * p.t_per_unit;
        const int t1 = min(t_last + 1, t0 + p.t_per_unit);
        // the dy slots of the previous unit may still be read: wait until every stage it used has been released
        if (ctr > 0) {
          for (int k = max(0, ctr - kWg2XStages); k < ctr; ++k)
            mbar_wait(smem_u32(&s_empty[k % kWg2XStages]), (k / kWg2XStages) & 1);
        }
        for (int t = t0; t < t1; ++t) {
          for (int r0 = 0; r0 < 2; ++r0) {
            const int h = 2 * t - 3 + r0;
            if (h < 0 || h >= p.IH) continue;
            const int s = ctr % kWg2XStages;
            // (stages re-used inside a unit: wait for the MMAs of the row 4 stages back)
            if (ctr >= kWg2XStages) mbar_wait(smem_u32(&s_empty[s]), ((ctr / kWg2XStages) & 1) ^ 1);
            const uint32_t bar = smem_u32(&s_full[s]);
            uint32_t tx = kWg2XRow;
            // dy rows travel with the first input row that needs them
            const bool first_of_t = (r0 == 0) || (2 * t - 3 < 0);
            uint32_t dy_tx = 0;
            if (first_of_t) {
              if (t == t0) {
                // unit start: the whole window t0-3 .. t0 (issued below, after expect_tx)
                for (int oh = t0 - 3; oh <= t0; ++oh) dy_tx += ((((-oh) % 6 + 6) % 6) < 3) ? 2u * kWg2DyRow : kWg2DyRow;
              } else {
                dy_tx = ((((-t) % 6 + 6) % 6) < 3) ? 2u * kWg2DyRow : kWg2DyRow;
              }
            }
            mbar_arrive_expect_tx(bar, tx + dy_tx);
            tma_load_4d(smem_base + s * kWg2XRow, &tm.a, bar, 0, 0, h, b);
            if (first_of_t) {
              if (t == t0) {
                for (int oh = t0 - 3; oh <= t0; ++oh) load_dy(oh, b, bar);
              } else {
                load_dy(t, b, bar);
              }
            }
            ++ctr;
          }
        }
      }
    }
  } else if (warp == 4) {
    if (elect_one()) {
      // ================================ MMA issuer ================================
      const uint32_t idesc = umma_idesc_f16(128, 128, 1, 1);
      const uint64_t da0 = umma_desc(0, 128, 1024, 128);           // x: 128-byte pair rows, blocks one row apart
      const uint64_t db0 = umma_desc(0, kWg2DyRow, 512, 64);       // dy: 64-byte pixel rows, blocks one ring slot apart
      const uint32_t a_hi = static_cast<uint32_t>(da0 >> 32), a_lo0 = static_cast<uint32_t>(da0);
      const uint32_t b_hi = static_cast<uint32_t>(db0 >> 32), b_lo0 = static_cast<uint32_t>(db0);
      uint32_t started = 0;  // bit (r0 * 2 + jh): accumulator already initialised
      int ctr = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        const int b = u / p.units_per_img;
        const int t0 = 1 + (u - b * p.units_per_img) * p.t_per_unit;
        const int t1 = min(t_last + 1, t0 + p.t_per_unit);
        (void)b;
        for (int t = t0; t < t1; ++t) {
          const int slot = ((-t) % 6 + 6) % 6;  // window = slots slot .. slot+3 (rows t, t-1, t-2, t-3)
          const uint32_t b_lo = b_lo0 + ((sDy + slot * kWg2DyRow) >> 4);
          for (int r0 = 0; r0 < 2; ++r0) {
            const int h = 2 * t - 3 + r0;
            if (h < 0 || h >= p.IH) continue;
            const int s = ctr % kWg2XStages;
            mbar_wait(smem_u32(&s_full[s]), (ctr / kWg2XStages) & 1);
            tc_fence_after();
            const uint32_t x_lo = a_lo0 + ((smem_base + s * kWg2XRow) >> 4);
#pragma unroll
            for (int jh = 0; jh < 2; ++jh) {
              const uint32_t acc_bit = 1u << (r0 * 2 + jh);
              const uint32_t d_tmem = tmem_base + static_cast<uint32_t>((r0 * 2 + jh) * 128);
              for (int q = 0; q < kWg2KSteps; ++q) {
                // K step q = output columns 16q .. 16q+15 -> staged pair rows (16q + 2jh) .., dy pixels 16q ..
                tc_mma_f16_parts(d_tmem, x_lo + static_cast<uint32_t>(16 * q + 2 * jh) * 8, a_hi, b_lo + static_cast<uint32_t>(q) * 64,
                                 b_hi, idesc, ((started & acc_bit) != 0u || q != 0) ? 1u : 0u);
              }
              started |= acc_bit;
            }
            tc_commit(smem_u32(&s_empty[s]));
            ++ctr;
          }
        }
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
    const int blk = tid >> 6, par = (tid >> 5) & 1, c = tid & 31;  // M index = (tap pair block, pixel parity, channel)
    for (int r0 = 0; r0 < 2; ++r0) {
      for (int jh = 0; jh < 2; ++jh) {
        const int s = 2 * (2 * jh + blk) + par;
        for (int q = 0; q < 4; ++q) {
          const int r = r0 + 2 * q;
          float v[32];
          tmem_ld32(t_row + (r0 * 2 + jh) * 128 + q * 32, v);
          tmem_ld_wait();
          if (s < 7 && r < 7) {
            float* dst = p.dw + (r * 7 + s) * 32 + c;
#pragma unroll
            for (int n = 0; n < 32; ++n) atomicAdd(dst + static_cast<int64_t>(n) * p.w_ld, v[n]);
          }
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

int conv_stem_wgrad2_supported(int IH, int IW) {
  const int OW = (IW + 6 - 7) / 2 + 1;
  return (OW <= 176 && OW >= 16 && IH >= 7) ? 1 : 0;
}

int conv_stem_wgrad2_launch(const __half* x, const __half* dy, float* dw, int w_ld, int B, int IH, int IW, cudaStream_t st) {
  PNVO_REQUIRE(x && dy && dw, "conv_stem_wgrad2: null pointer");
  PNVO_REQUIRE(conv_stem_wgrad2_supported(IH, IW), "conv_stem_wgrad2: unsupported geometry %dx%d", IH, IW);
  PNVO_REQUIRE(w_ld >= 49 * 32, "conv_stem_wgrad2: w_ld too small");
  StemWg2Args a{};
  a.dw = dw; a.w_ld = w_ld; a.B = B; a.IH = IH;
  a.OH = (IH + 6 - 7) / 2 + 1;
  a.OW = (IW + 6 - 7) / 2 + 1;
  const int n_t = (IH + 2) / 2;          // t = 1 .. t_last
  a.units_per_img = std::max(1, std::min(4, n_t / 8));
  a.t_per_unit = ceil_div(n_t, a.units_per_img);
  a.units_per_img = ceil_div(n_t, a.t_per_unit);
  a.n_units = B * a.units_per_img;
  if (B <= 0) return 0;
  const int Wp = stem_padded_width(IW);
  alignas(64) ConvTmaps tm;
  memset(&tm, 0, sizeof(tm));
  if (tmap_tiled4d(&tm.a, x, B, IH, Wp / 2, 64, kWg2XRow / 128)) return -1;        // 184 pixel pairs per box
  if (tmap_tiled4d(&tm.b, dy, B, a.OH, a.OW, 32, kWg2DyRow / 64, 64)) return -1;   // 176 pixels x 32 channels, SWIZZLE_64B
  const int smem_bytes = kWg2XStages * kWg2XRow + kWg2DySlots * kWg2DyRow + 1024;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(conv_stem_wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    attr = true;
  }
  conv_stem_wgrad2_kernel<<<std::min(a.n_units, 148), 192, smem_bytes, st>>>(a, tm);
  count_launch();
  return check_launch("conv_stem_wgrad2");
}

}  // namespace pnvo
