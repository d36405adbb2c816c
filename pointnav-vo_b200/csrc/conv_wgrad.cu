// Weight gradient of the convolution as a tcgen05 GEMM whose reduction dimension is the output pixel:
//
//   dWp[n, k] = sum_m dy[m, n] * A[m, k]        A = im2col(x) as in conv_igemm.cu, k = (r, s, c)
//
// computed as D[k, n] (UMMA M = 128 rows of k, N = Cout tile) so that both operands are "MN-major": a
// stage holds 64 pixels; every pixel row is 128 bytes of consecutive k (A) or consecutive n (B), which
// is exactly how NHWC activations and gradients lie in memory -- no transposes.  Each CTA owns `mt`
// 128-row tiles of k (separate TMEM accumulators sharing the same dy stage), a Cout tile, and a
// contiguous range of 64-pixel chunks (split over grid.z); partial results are combined with
// coalesced fp32 atomics into the packed dW buffer.
#include "common.cuh"
#include "ops.cuh"

namespace pnvo {

static constexpr int kPix = 64;  // pixels (GEMM-K) per stage
static constexpr int kWProducerThreads = 128;

__global__ void __launch_bounds__(160) conv_wgrad_kernel(const WgradArgs p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_full[4];
  __shared__ __align__(8) uint64_t s_empty[4];
  __shared__ __align__(8) uint64_t s_accum;
  __shared__ uint32_t s_tmem;
  __shared__ short2 s_tap[128];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int stages = p.stages;
  const int tile0 = blockIdx.x * p.mt;                     // first 128-row k tile of this CTA
  const int mt = min(p.mt, p.n_mtiles - tile0);            // tiles actually present
  const int n0 = blockIdx.y * p.N;
  const int total_chunks = ceil_div(p.M, kPix);
  const int c_begin = blockIdx.z * p.chunks_per_split;
  const int c_end = min(total_chunks, c_begin + p.chunks_per_split);
  const int n_chunks = max(0, c_end - c_begin);
  const int nb = (p.N + 63) >> 6;                           // 64-channel blocks of the dy tile
  const uint32_t a_bytes = static_cast<uint32_t>(p.mt) * 2 * 8192;
  const uint32_t b_bytes = static_cast<uint32_t>(nb) * 8192;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;

  for (int i = tid; i < p.R * p.S; i += blockDim.x) s_tap[i] = make_short2(i / p.S, i % p.S);
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&s_full[s]), kWProducerThreads);
      mbar_init(smem_u32(&s_empty[s]), 1);
    }
    mbar_init(smem_u32(&s_accum), 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(&s_tmem), p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;

  if (n_chunks > 0) {
    if (warp < 4) {
      // =============================== producers ===============================
      const int row = tid & 63;   // pixel row inside the stage
      const int half = tid >> 6;  // which half of the column blocks this thread fills
      const int rx = row & 7;
      const uint32_t row_off = static_cast<uint32_t>((row >> 3) * 1024 + (row & 7) * 128);
      const int cmask = p.cmask;
      const int ohw = p.OH * p.OW;
      const int n_ablocks = mt * 2;

      auto issue_stage = [&](int it) {
        const int s = it % stages;
        const uint32_t sA = smem_base + s * stage_bytes;
        const uint32_t sB = sA + a_bytes;
        const int m = (c_begin + it) * kPix + row;
        const bool row_valid = m < p.M;
        int b = 0, oh = 0, ow = 0;
        if (row_valid) {
          b = m / ohw;
          const int rem = m - b * ohw;
          oh = rem / p.OW;
          ow = rem - oh * p.OW;
        }
        const int ohb = oh * p.mul - p.pad, owb = ow * p.mul - p.pad_w;
        const __half* __restrict__ xb = p.x + static_cast<int64_t>(b) * p.IH * p.IW * p.Cin;
        // ---- A: im2col rows, blocks of 64 consecutive k ----
        for (int blk = half; blk < n_ablocks; blk += 2) {
          const int kf0 = (tile0 * 2 + blk) * 64;
          const uint32_t dst = sA + static_cast<uint32_t>(blk) * 8192 + row_off;
          const __half* src = nullptr;
          bool ok = false;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int kf = kf0 + j * 8;
            if (j == 0 || (kf & cmask) == 0) {
              ok = false;
              if (row_valid && kf < p.K) {
                const int tap = kf >> p.cin_log2;
                const short2 rs = s_tap[tap];
                const int ih = ohb + rs.x, iw = owb + rs.y;
                if (ih >= 0 && iw >= 0 && ih < p.IH && iw < p.IW) {
                  ok = true;
                  src = xb + (static_cast<int64_t>(ih) * p.IW + iw) * p.Cin + (kf & cmask);
                }
              }
            } else {
              src += 8;
            }
            cp_async_16(dst + ((j ^ rx) << 4), ok ? static_cast<const void*>(src) : static_cast<const void*>(p.x),
                        ok ? 16u : 0u);
          }
        }
        // ---- B: dy rows ----
        const __half* dyr = p.dy + static_cast<int64_t>(row_valid ? m : 0) * p.ld_dy + n0;
        const int n_bchunks = p.N >> 3;  // 16-byte chunks per pixel row
        for (int q = half; q < n_bchunks; q += 2) {
          const int blk = q >> 3, j = q & 7;
          cp_async_16(sB + static_cast<uint32_t>(blk) * 8192 + row_off + ((j ^ rx) << 4), dyr + q * 8,
                      row_valid ? 16u : 0u);
        }
      };

      for (int it = 0; it < n_chunks; ++it) {
        const int s = it % stages;
        if (it >= stages) mbar_wait(smem_u32(&s_empty[s]), ((it / stages) & 1) ^ 1);
        issue_stage(it);
        cp_async_commit();
        if (it >= 2) {
          cp_async_wait<2>();
          fence_proxy_async_smem();
          mbar_arrive(smem_u32(&s_full[(it - 2) % stages]));
        }
      }
      if (n_chunks >= 2) {
        cp_async_wait<1>();
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&s_full[(n_chunks - 2) % stages]));
      }
      cp_async_wait<0>();
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&s_full[(n_chunks - 1) % stages]));

      // =============================== epilogue ===============================
      mbar_wait(smem_u32(&s_accum), 0);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
      for (int i = 0; i < mt; ++i) {
        const int kf = (tile0 + i) * 128 + tid;
        for (int ch = 0; ch * 32 < p.N; ++ch) {
          float v[32];
          tmem_ld32(t_row + i * p.N + ch * 32, v);
          tmem_ld_wait();
          if (kf < p.K) {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const int n = n0 + ch * 32 + c;
              if (ch * 32 + c < p.N && n < p.n_total) atomicAdd(p.dw + static_cast<int64_t>(n) * p.w_ld + kf, v[c]);
            }
          }
        }
      }
      tc_fence_before();
    } else {
      // =============================== MMA issuer ===============================
      if (lane == 0) {
        const uint32_t idesc = umma_idesc_f16(128, p.N, 1, 1);
        for (int it = 0; it < n_chunks; ++it) {
          const int s = it % stages;
          mbar_wait(smem_u32(&s_full[s]), (it / stages) & 1);
          tc_fence_after();
          const uint32_t sA = smem_base + s * stage_bytes;
          const uint64_t bdesc = umma_desc_sw128(sA + a_bytes, 8192, 1024);
#pragma unroll
          for (int k = 0; k < kPix / 16; ++k) {
            // 16 pixels of K = two 8-row groups = 2048 bytes
            const uint64_t koff = static_cast<uint64_t>((k * 2048) >> 4);
            for (int i = 0; i < mt; ++i) {
              const uint64_t adesc = umma_desc_sw128(sA + static_cast<uint32_t>(i) * 16384, 8192, 1024);
              tc_mma_f16(tmem_base + static_cast<uint32_t>(i * p.N), adesc + koff, bdesc + koff, idesc,
                         (it | k) != 0 ? 1u : 0u);
            }
          }
          tc_commit(smem_u32(&s_empty[s]));
        }
        tc_commit(smem_u32(&s_accum));
      }
      __syncwarp();
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

int wgrad_plan(WgradArgs& a) {
  PNVO_REQUIRE(a.Cin >= 8 && a.Cin % 8 == 0 && (a.R * a.S == 1 || (a.Cin & (a.Cin - 1)) == 0),
               "wgrad: Cin=%d must be a multiple of 8 (power of two unless 1x1)", a.Cin);
  PNVO_REQUIRE(a.R * a.S <= 128, "wgrad: filter too large");
  PNVO_REQUIRE(a.n_total % 16 == 0, "wgrad: padded Cout=%d must be a multiple of 16", a.n_total);
  if (a.R * a.S == 1) {
    a.cin_log2 = 30;
    a.cmask = 0x3fffffff;
  } else {
    a.cin_log2 = 0;
    while ((1 << a.cin_log2) < a.Cin) ++a.cin_log2;
    a.cmask = a.Cin - 1;
  }
  a.M = a.B * a.OH * a.OW;
  a.K = a.R * a.S * a.Cin;
  PNVO_REQUIRE(a.w_ld >= a.K, "wgrad: w_ld %d < K %d", a.w_ld, a.K);
  a.n_mtiles = ceil_div(a.K, 128);
  int N = a.n_total;
  if (N > 256) {
    N = 256;
    while (a.n_total % N) N -= 32;
  }
  a.N = N;
  a.n_ntiles = a.n_total / N;
  const int nb = (N + 63) / 64;
  int mt = std::min(a.n_mtiles, 512 / N);
  while (mt > 1 && (mt * 16384 + nb * 8192) > 64 * 1024) --mt;
  a.mt = mt;
  int cols = 32;
  while (cols < mt * N) cols <<= 1;
  a.tmem_cols = cols;
  a.stages = 3;
  a.smem_bytes = a.stages * (mt * 16384 + nb * 8192) + 1024;
  a.grid_x = ceil_div(a.n_mtiles, mt);
  a.grid_y = a.n_ntiles;
  const int total_chunks = ceil_div(a.M, kPix);
  // split the pixel range so that the grid is about two waves of 148 SMs, at least 4 chunks per CTA
  int splits = std::max(1, (2 * 148) / std::max(1, a.grid_x * a.grid_y));
  splits = std::min(splits, std::max(1, total_chunks / 4));
  a.chunks_per_split = ceil_div(total_chunks, splits);
  a.grid_z = ceil_div(total_chunks, a.chunks_per_split);
  return 0;
}

int wgrad_launch(WgradArgs a, cudaStream_t st) {
  if (wgrad_plan(a)) return -1;
  PNVO_REQUIRE(a.x && a.dy && a.dw, "wgrad: null pointer");
  if (a.M == 0) return 0;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  conv_wgrad_kernel<<<dim3(a.grid_x, a.grid_y, a.grid_z), 160, a.smem_bytes, st>>>(a);
  count_launch();
  return check_launch("conv_wgrad");
}

}  // namespace pnvo
