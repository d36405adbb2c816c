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
#include "tmap.cuh"

namespace pnvo {

static constexpr int kPix = 64;  // pixels (GEMM-K) per stage
static constexpr int kWProducerThreads = 128;

__global__ void __launch_bounds__(160) conv_wgrad_kernel(const WgradArgs p, const __grid_constant__ ConvTmaps tm) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_full[4];
  __shared__ __align__(8) uint64_t s_empty[4];
  __shared__ __align__(8) uint64_t s_accum;
  __shared__ uint32_t s_tmem;
  __shared__ short2 s_tap[64];

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
  // TMA path: A blocks are chunk_k (32 or 64) k-columns wide, dy blocks min(64, N) channels wide; a pixel row of a
  // block is 64 or 128 bytes (SWIZZLE_64B / SWIZZLE_128B).  The cp.async path always uses 128-byte rows.
  const uint32_t a_row = p.tma ? static_cast<uint32_t>(p.chunk_k) * 2 : 128u;
  const int b_cols = p.tma ? min(64, p.N) : 64;
  const uint32_t b_row = static_cast<uint32_t>(b_cols) * 2;
  const uint32_t b_bytes = p.tma ? static_cast<uint32_t>(p.N / b_cols) * kPix * b_row : static_cast<uint32_t>(nb) * 8192;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;

  for (int i = tid; i < p.R * p.S; i += blockDim.x) s_tap[i] = make_short2(i / p.S, i % p.S);
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&s_full[s]), p.tma ? 1 : kWProducerThreads);
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

  if (n_chunks > 0 && p.tma) {
    // k tiles past the end of K (zero-padded rows of dW) are never loaded: clear them once so the MMA reads zeros
    if ((tile0 + p.mt) * 128 > p.K) {
      for (uint32_t off = tid * 16; off < static_cast<uint32_t>(stages) * stage_bytes; off += blockDim.x * 16)
        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(smem_base + off), "r"(0u) : "memory");
      fence_proxy_async_smem();
    }
    __syncthreads();
    if (warp == 0 && elect_one()) {
      // ================================ TMA producer ================================
      tma_prefetch_desc(&tm.a);
      tma_prefetch_desc(&tm.b);
      const int ohw = p.OH * p.OW;
      const int blocks_per_tile = 128 / p.chunk_k;
      const uint32_t a_blk_bytes = kPix * a_row, b_blk_bytes = kPix * b_row;
      const int n_bblk = p.N / b_cols;
      int n_ablk = 0;  // blocks with k < K
      for (int blk = 0; blk < mt * blocks_per_tile; ++blk) n_ablk += ((tile0 * 128 + blk * p.chunk_k) < p.K) ? 1 : 0;
      const uint32_t tx = static_cast<uint32_t>(n_ablk) * a_blk_bytes + static_cast<uint32_t>(n_bblk) * b_blk_bytes;
      for (int it = 0; it < n_chunks; ++it) {
        const int s = it % stages;
        if (it >= stages) mbar_wait(smem_u32(&s_empty[s]), ((it / stages) & 1) ^ 1);
        const uint32_t bar = smem_u32(&s_full[s]);
        const uint32_t sA = smem_base + s * stage_bytes;
        const int mc = (c_begin + it) * kPix;
        const int n_img = mc / ohw;
        const int rem = mc - n_img * ohw;
        const int p0 = rem / p.OW, q0 = rem - p0 * p.OW;
        const int w0 = q0 * p.mul - p.pad_w, h0 = p0 * p.mul - p.pad;
        mbar_arrive_expect_tx(bar, tx);
        for (int blk = 0; blk < mt * blocks_per_tile; ++blk) {
          const int kf = tile0 * 128 + blk * p.chunk_k;
          if (kf < p.K) {
            const short2 rs = s_tap[kf >> p.cin_log2];
            tma_load_im2col_4d(sA + blk * a_blk_bytes, &tm.a, bar, kf & p.cmask, w0, h0, n_img,
                               static_cast<uint16_t>(rs.y), static_cast<uint16_t>(rs.x));
          }
        }
        for (int blk = 0; blk < n_bblk; ++blk)
          tma_load_2d(sA + a_bytes + blk * b_blk_bytes, &tm.b, bar, n0 + blk * b_cols, mc);
      }
    }
  }
  if (n_chunks > 0) {
    if (warp < 4) {
      if (!p.tma) {
      // =============================== producers ===============================
      // Thread t owns 16-byte chunk j = t % 8 of pixel rows (t / 8) + 16 i, i = 0..3, in every 64-wide
      // column block: the 8 lanes of a row fetch one contiguous 128-byte line (coalesced L2 requests).
      const int j = tid & 7, rsub = tid >> 3;
      const uint32_t t_off = static_cast<uint32_t>((rsub >> 3) * 1024 + (rsub & 7) * 128 + ((j ^ (rsub & 7)) << 4));
      const int ohw = p.OH * p.OW;
      const int n_ablocks = mt * 2;
      const int n_bblocks = (p.N + 63) >> 6;
      // running (sample, oh, ow) of this thread's 4 pixel rows; advanced by 64 pixels per stage
      int pb[4], poh[4], pow_[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = c_begin * kPix + rsub + 16 * i;
        pb[i] = m / ohw;
        const int rem = m - pb[i] * ohw;
        poh[i] = rem / p.OW;
        pow_[i] = rem - poh[i] * p.OW;
      }

      auto issue_stage = [&](int it) {
        const int s = it % stages;
        const uint32_t sA = smem_base + s * stage_bytes + t_off;
        const uint32_t sB = sA + a_bytes;
        const __half* rowptr[4];
        const __half* dyptr[4];
        int rlo[4], rspan[4], slo[4], sspan[4];
        uint32_t dy_ok[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int m = (c_begin + it) * kPix + rsub + 16 * i;
          const bool valid = m < p.M;
          const int ohb = poh[i] * p.mul - p.pad, owb = pow_[i] * p.mul - p.pad_w;
          rowptr[i] = p.x + ((static_cast<int64_t>(pb[i]) * p.IH + ohb) * p.IW + owb) * p.Cin;
          // tap (r, s) is inside the input iff rlo <= r < rlo + rspan and slo <= s < slo + sspan
          rlo[i] = max(0, -ohb);
          rspan[i] = valid ? max(0, min(p.R, p.IH - ohb) - rlo[i]) : 0;
          slo[i] = max(0, -owb);
          sspan[i] = max(0, min(p.S, p.IW - owb) - slo[i]);
          dyptr[i] = p.dy + static_cast<int64_t>(valid ? m : 0) * p.ld_dy + n0 + j * 8;
          dy_ok[i] = valid ? 16u : 0u;
          // advance to the next stage
          pow_[i] += kPix;
          while (pow_[i] >= p.OW) { pow_[i] -= p.OW; ++poh[i]; }
          while (poh[i] >= p.OH) { poh[i] -= p.OH; ++pb[i]; }
        }
        // ---- A: im2col rows, blocks of 64 consecutive k ----
        for (int blk = 0; blk < n_ablocks; ++blk) {
          const int kf = (tile0 * 2 + blk) * 64 + j * 8;
          int r = -(1 << 20), sx = 0, toff = 0;
          if (kf < p.K) {
            const int tap = kf >> p.cin_log2;
            const short2 rs = s_tap[tap];
            r = rs.x;
            sx = rs.y;
            toff = (r * p.IW + sx) * p.Cin + (kf & p.cmask);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const bool ok = (static_cast<unsigned>(r - rlo[i]) < static_cast<unsigned>(rspan[i])) &&
                            (static_cast<unsigned>(sx - slo[i]) < static_cast<unsigned>(sspan[i]));
            cp_async_16(sA + blk * 8192 + i * 2048, rowptr[i] + toff, ok ? 16u : 0u);
          }
        }
        // ---- B: dy rows ----
        for (int blk = 0; blk < n_bblocks; ++blk) {
          if (blk * 64 + j * 8 < p.N) {
#pragma unroll
            for (int i = 0; i < 4; ++i) cp_async_16(sB + blk * 8192 + i * 2048, dyptr[i] + blk * 64, dy_ok[i]);
          }
        }
      };

      const int la = p.lookahead;
      for (int it = 0; it < n_chunks; ++it) {
        const int s = it % stages;
        if (it >= stages) mbar_wait(smem_u32(&s_empty[s]), ((it / stages) & 1) ^ 1);
        issue_stage(it);
        cp_async_commit();
        if (it >= la) {
          cp_async_wait_dyn(la);
          fence_proxy_async_smem();
          mbar_arrive(smem_u32(&s_full[(it - la) % stages]));
        }
      }
      for (int rem = min(la, n_chunks) - 1; rem >= 0; --rem) {
        cp_async_wait_dyn(rem);
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&s_full[(n_chunks - 1 - rem) % stages]));
      }

      }  // !tma
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
          // The flush is bound by the NUMBER of L2 atomic operations (measured ~90 per clock chip-wide: 9.4 M scalar atomics
          // = 50 of the 58 us of a layer3 launch).  dW rows are contiguous along kf, the accumulator has one kf per lane:
          // a 4 x 4 transpose inside every group of four lanes (two shuffle rounds) gives each lane FOUR consecutive kf of one
          // column, flushed with one 16-byte red.global.add.v4.f32 -- a quarter of the operations.
          const int kf4 = (tile0 + i) * 128 + (tid & ~3);
          const int q = lane & 3;
#pragma unroll
          for (int c0 = 0; c0 < 32; c0 += 4) {
            float a0 = v[c0], a1 = v[c0 + 1], a2 = v[c0 + 2], a3 = v[c0 + 3];
            // round 1 (lane ^ 1): 2 x 2 blocks
            {
              const float s0 = (q & 1) ? a0 : a1, s1 = (q & 1) ? a2 : a3;
              const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
              if (q & 1) { a0 = r0; a2 = r1; } else { a1 = r0; a3 = r1; }
            }
            // round 2 (lane ^ 2)
            {
              const float s0 = (q & 2) ? a0 : a2, s1 = (q & 2) ? a1 : a3;
              const float r0 = __shfl_xor_sync(0xffffffffu, s0, 2), r1 = __shfl_xor_sync(0xffffffffu, s1, 2);
              if (q & 2) { a0 = r0; a1 = r1; } else { a2 = r0; a3 = r1; }
            }
            // lane q now holds column c0 + q for kf4 .. kf4 + 3
            const int col = ch * 32 + c0 + q, n = n0 + col;
            if (kf4 < p.K && col < p.N && n < p.n_total) {
              float* dst = p.dw + static_cast<int64_t>(n) * p.w_ld + kf4;
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a0), "f"(a1), "f"(a2), "f"(a3) : "memory");
            }
          }
        }
      }
      tc_fence_before();
    } else {
      // =============================== MMA issuer ===============================
      if (elect_one()) {
        const uint32_t idesc = umma_idesc_f16(128, p.N, 1, 1);
        // MN-major operands: LBO = distance between successive M/N blocks (64 pixel rows), SBO = 8 pixel rows.  The
        // descriptors are split into loop-invariant (lo, hi) words: per MMA one add per operand (the first version rebuilt
        // the 64-bit A descriptor inside the innermost loop and the issuing thread, not the tensor pipe, set the pace).
        const uint64_t da0 = umma_desc(0, kPix * a_row, 8 * a_row, a_row);
        const uint64_t db0 = umma_desc(0, kPix * b_row, 8 * b_row, b_row);
        const uint32_t a_hi = static_cast<uint32_t>(da0 >> 32), b_hi = static_cast<uint32_t>(db0 >> 32);
        const uint32_t a_lo0 = static_cast<uint32_t>(da0) + (smem_base >> 4), b_lo0 = static_cast<uint32_t>(db0) + ((smem_base + a_bytes) >> 4);
        const uint32_t stage16 = stage_bytes >> 4;
        const uint32_t ka = static_cast<uint32_t>(16 * a_row) >> 4, kb16 = static_cast<uint32_t>(16 * b_row) >> 4;
        uint32_t slot = 0, phase = 0;
        for (int it = 0; it < n_chunks; ++it) {
          mbar_wait(smem_u32(&s_full[slot]), phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + slot * stage16, b_lo = b_lo0 + slot * stage16;
#pragma unroll
          for (int k = 0; k < kPix / 16; ++k) {
            // 16 pixels of GEMM-K = 16 rows of each block
            for (int i = 0; i < mt; ++i)
              tc_mma_f16_parts(tmem_base + static_cast<uint32_t>(i * p.N), a_lo + static_cast<uint32_t>(i) * 1024u + k * ka, a_hi,
                               b_lo + k * kb16, b_hi, idesc, (it | k) != 0 ? 1u : 0u);
          }
          tc_commit(smem_u32(&s_empty[slot]));
          if (++slot == static_cast<uint32_t>(stages)) {
            slot = 0;
            phase ^= 1u;
          }
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
  PNVO_REQUIRE(a.R * a.S <= 64, "wgrad: filter too large");
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
  a.tma = (a.Cin % 32 == 0 && a.R == a.S && a.pad == a.pad_w && a.force_generic != 1) ? 1 : 0;
  a.chunk_k = (a.tma && a.Cin % 64 != 0) ? 32 : 64;
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
  while (mt > 1 && (mt * 16384 + nb * 8192) > 48 * 1024) --mt;  // (64 KB / 3 stages was measured slower)
  a.mt = mt;
  int cols = 32;
  while (cols < mt * N) cols <<= 1;
  a.tmem_cols = cols;
  a.stages = std::max(3, std::min(4, (200 * 1024) / (mt * 16384 + nb * 8192)));
  a.lookahead = a.stages - 2;
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
  if (a.force_generic == 0 && a.x && a.dy && a.dw && wgrad_direct1_supported(a)) return wgrad_direct1_launch(a, st);
  if (a.force_generic == 0 && a.x && a.dy && a.dw && wgrad_raster_supported(a)) return wgrad_raster_launch(a, st);
  if (wgrad_plan(a)) return -1;
  PNVO_REQUIRE(a.x && a.dy && a.dw, "wgrad: null pointer");
  PNVO_REQUIRE(a.tma || a.x_row_pitch == 0 || a.x_row_pitch == a.IW, "wgrad: padded input rows need the TMA path");
  if (a.M == 0) return 0;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  alignas(64) ConvTmaps tm;
  memset(&tm, 0, sizeof(tm));
  if (a.tma) {
    if (tmap_im2col(&tm.a, a.x, a.B, a.IH, a.IW, a.Cin, a.R, a.S, a.mul, a.pad, a.chunk_k, kPix, a.x_row_pitch)) return -1;
    if (tmap_tiled2d(&tm.b, a.dy, a.M, a.n_total, a.ld_dy, kPix, std::min(64, a.N))) return -1;
  }
  conv_wgrad_kernel<<<dim3(a.grid_x, a.grid_y, a.grid_z), 160, a.smem_bytes, st>>>(a, tm);
  count_launch();
  return check_launch("conv_wgrad");
}

}  // namespace pnvo
