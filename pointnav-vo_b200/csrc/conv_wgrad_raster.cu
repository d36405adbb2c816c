// Weight gradient of the 3x3 / stride 1 / pad 1 convolutions with 32 or 64 channels (layer1 / layer2 of the
// GroupNorm ResNet, resnet.py:11-26) on the shared-memory raster of conv_raster.cu -- no im2col expansion:
//
//   dW[n, (r, s, c)] = sum_{b, oh, ow} dy[b, oh, ow, n] * x[b, oh - 1 + r, ow - 1 + s, c]
//
// GEMM-K is the output position.  A unit = (sample, T output rows); its T+2 input rows (W+2 pixels wide, zero halo
// from the TMA out-of-range fill) and its T rows of dy are staged as linear rasters of pitch P = W+2: dy's two
// extra columns per row are out of range and arrive as ZEROS, so junk positions contribute nothing.  Both operands
// are MN-major exactly as NHWC lays them out: a K row is one pixel = 64 / 128 bytes of consecutive channels.  For
// filter row r the A operand of one tcgen05.mma is 128/C adjacent taps (M blocks one pixel = LBO apart) starting at
// raster position (k + r*P + s0): the same staged bytes serve all nine taps.  C = 32: 3 accumulators (taps s = 0..3
// of each r, the 4th is junk); C = 64: 6 accumulators (s = {0,1} and {2, junk}).  The accumulators stay in TMEM for
// every unit the persistent CTA processes and are flushed once with coalesced fp32 atomics.
//   warp 5: TMA producer    warp 4: TMEM owner + MMA issuer    warps 0-3: final epilogue
#include "common.cuh"
#include "ops.cuh"
#include "tmap.cuh"

namespace pnvo {

struct WgRasterArgs {
  float* dw;
  int w_ld;
  int B, H, W;
  int P, T, n_k, rows_in;
  int units_per_img, n_units;
  int x_bytes, dy_bytes;  // shared-memory bytes of one staged x / dy raster (multiples of 1024)
};

template <int C>
__global__ void __launch_bounds__(192) conv_wgrad_raster_kernel(const WgRasterArgs p,
                                                                const __grid_constant__ ConvTmaps tm) {
  constexpr int N = C;                   // Cout == Cin for these layers
  constexpr int kPix = C * 2;            // bytes per pixel = K-row bytes of both operands
  constexpr int kNB = 128 / C;           // taps (M blocks) per MMA
  constexpr int kMPR = (3 + kNB - 1) / kNB;  // MMAs per filter row
  constexpr int kAcc = 3 * kMPR;
  constexpr int kCols = kAcc * N <= 128 ? 128 : 512;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_full[2];
  __shared__ __align__(8) uint64_t s_empty[2];
  __shared__ __align__(8) uint64_t s_accum;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t stage_bytes = static_cast<uint32_t>(p.x_bytes + p.dy_bytes);

  // zero both stages once: the tails behind the TMA boxes are read by the MMAs of the last K step / the junk tap
  // and must be finite (x) resp. zero (dy)
  for (uint32_t off = tid * 16; off < 2 * stage_bytes; off += blockDim.x * 16)
    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(smem_base + off), "r"(0u) : "memory");
  fence_proxy_async_smem();
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&s_full[s]), 1);
      mbar_init(smem_u32(&s_empty[s]), 1);
    }
    mbar_init(smem_u32(&s_accum), 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(&s_tmem), kCols);
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
      const uint32_t tx = static_cast<uint32_t>(p.rows_in + p.T) * p.P * kPix;
      int i = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++i) {
        const int slot = i & 1;
        if (i >= 2) mbar_wait(smem_u32(&s_empty[slot]), ((i >> 1) & 1) ^ 1);
        const int b = u / p.units_per_img;
        const int h0 = (u - b * p.units_per_img) * p.T;
        const uint32_t bar = smem_u32(&s_full[slot]);
        const uint32_t sX = smem_base + slot * stage_bytes;
        mbar_arrive_expect_tx(bar, tx);
        tma_load_4d(sX, &tm.a, bar, 0, -1, h0 - 1, b);            // x rows h0-1 .. h0+T, pixels -1 .. W
        tma_load_4d(sX + p.x_bytes, &tm.b, bar, 0, 0, h0, b);     // dy rows h0 .. h0+T-1, pixels 0 .. W+1 (>= W: zeros)
      }
    }
  } else if (warp == 4) {
    if (elect_one()) {
      // ================================ MMA issuer ================================
      const uint32_t idesc = umma_idesc_f16(128, N, 1, 1);
      // (lo, hi) descriptors: hi (SBO, version, swizzle) and the LBO bits of lo are shared by A and B; per-row deltas
      // are loop invariant -> one add per operand between two MMAs
      const uint64_t d0 = umma_desc(0, kPix, 8 * kPix, kPix);
      const uint32_t hi = static_cast<uint32_t>(d0 >> 32), lo0 = static_cast<uint32_t>(d0);
      uint32_t row_a[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) row_a[r] = static_cast<uint32_t>(r * p.P * kPix) >> 4;
      int i = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++i) {
        const int slot = i & 1;
        mbar_wait(smem_u32(&s_full[slot]), (i >> 1) & 1);
        tc_fence_after();
        const uint32_t sX = smem_base + slot * stage_bytes;
        uint32_t x_lo = lo0 + (sX >> 4), d_lo = lo0 + ((sX + p.x_bytes) >> 4);
        for (int q = 0; q < p.n_k; ++q, x_lo += kPix, d_lo += kPix) {  // 16 positions * kPix bytes >> 4 = kPix
          // K step q = output raster positions 16q .. 16q+15
#pragma unroll
          for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int h = 0; h < kMPR; ++h)
              tc_mma_f16_lohi(tmem_base + static_cast<uint32_t>((r * kMPR + h) * N), x_lo + row_a[r] + ((h * kNB * kPix) >> 4),
                              d_lo, hi, idesc, (i | q) != 0 ? 1u : 0u);
          }
        }
        tc_commit(smem_u32(&s_empty[slot]));
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
    const int blk = tid / C, c = tid % C;  // M index = (tap block, channel)
    for (int a = 0; a < kAcc; ++a) {
      const int r = a / kMPR, s = (a % kMPR) * kNB + blk;
#pragma unroll
      for (int ch = 0; ch < N / 32; ++ch) {
        float v[32];
        tmem_ld32(t_row + a * N + ch * 32, v);
        tmem_ld_wait();
        if (s < 3) {
          float* dst = p.dw + (r * 3 + s) * C + c;
#pragma unroll
          for (int n = 0; n < 32; ++n) atomicAdd(dst + static_cast<int64_t>(ch * 32 + n) * p.w_ld, v[n]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kCols);
  }
}

static bool wg_raster_plan(const WgradArgs& a, WgRasterArgs& r, int& smem_bytes) {
  if (a.R != 3 || a.S != 3 || a.mul != 1 || a.pad != 1 || a.pad_w != 1) return false;
  if (a.IH != a.OH || a.IW != a.OW) return false;
  if (!(a.Cin == 32 || a.Cin == 64) || a.n_total != a.Cin || a.ld_dy != a.n_total) return false;
  if (a.x_row_pitch != 0 && a.x_row_pitch != a.IW) return false;
  if (a.w_ld < 9 * a.Cin) return false;
  const int P = a.IW + 2;
  if (P > 256 || a.IW < 8 || a.IH < 4) return false;
  const int pix = a.Cin * 2;
  double best = -1.0;
  for (int T = 2; T <= std::min(a.IH, 64); ++T) {
    const int rows_in = T + 2;
    const int n_k = ceil_div(T * P, 16);
    const int x_pos = std::max(rows_in * P, n_k * 16 + 2 * P + 4);
    const int x_bytes = (x_pos * pix + 1023) & ~1023;
    const int dy_bytes = (n_k * 16 * pix + 1023) & ~1023;
    const int smem = 2 * (x_bytes + dy_bytes) + 1024;
    if (smem > 200 * 1024) break;
    const int upi = ceil_div(a.IH, T);
    const int n_units = a.B * upi;
    const int waves = ceil_div(n_units, 148);
    const double balance = n_units >= 148 ? static_cast<double>(n_units) / (waves * 148.0) : 1.0;
    const double eff = (static_cast<double>(a.IH) * a.IW) / (static_cast<double>(upi) * n_k * 16) * balance *
                       (1.0 - 0.15 * 2.0 / (T + 2));
    if (eff > best + 1e-9) {
      best = eff;
      r.P = P; r.T = T; r.n_k = n_k; r.rows_in = rows_in; r.units_per_img = upi; r.n_units = n_units;
      r.x_bytes = x_bytes; r.dy_bytes = dy_bytes;
      smem_bytes = smem;
    }
  }
  if (best < 0.5) return false;
  r.dw = a.dw; r.w_ld = a.w_ld; r.B = a.B; r.H = a.IH; r.W = a.IW;
  return true;
}

int wgrad_raster_supported(const WgradArgs& a) {
  WgRasterArgs r{};
  int smem = 0;
  return wg_raster_plan(a, r, smem) ? 1 : 0;
}

template <int C>
static int wg_raster_launch_t(const WgRasterArgs& r, const ConvTmaps& tm, int smem, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(conv_wgrad_raster_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  conv_wgrad_raster_kernel<C><<<std::min(r.n_units, 148), 192, smem, st>>>(r, tm);
  count_launch();
  return check_launch("conv_wgrad_raster");
}

int wgrad_raster_launch(const WgradArgs& a, cudaStream_t st) {
  WgRasterArgs r{};
  int smem = 0;
  PNVO_REQUIRE(wg_raster_plan(a, r, smem), "wgrad_raster: unsupported geometry");
  if (a.B <= 0) return 0;
  alignas(64) ConvTmaps tm;
  memset(&tm, 0, sizeof(tm));
  const int sw = a.Cin * 2 == 128 ? 128 : 64;
  if (tmap_tiled4d(&tm.a, a.x, a.B, a.IH, a.IW, a.Cin, r.P, sw, r.rows_in)) return -1;
  if (tmap_tiled4d(&tm.b, a.dy, a.B, a.OH, a.OW, a.n_total, r.P, sw, r.T)) return -1;
  if (a.Cin == 32) return wg_raster_launch_t<32>(r, tm, smem, st);
  return wg_raster_launch_t<64>(r, tm, smem, st);
}

}  // namespace pnvo
