// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 + TMEM), NHWC fp16 in, fp32
// accumulate.  One kernel covers the forward conv (resnet.py:11-26,156-164; vo_cnn.py:85-91) and the
// data gradient (same GEMM over flipped/transposed weights, stride turned into a divisibility test):
//
//   D[m, n] = sum_k A[m, k] * Wp[n, k]      m = output pixel (b, oh, ow) flattened over the batch
//                                           k = (r, s, c) flattened, c fastest (matches NHWC)
//   A[m, (r,s,c)] = x[b, (oh*mul - pad + r)/div, (ow*mul - pad + s)/div, c]   (0 when out of range or
//                                                                              not divisible)
//
// CTA = 128 output pixels x N (<=256) output channels.  Warps 0-3: im2col producers (16-byte cp.async
// with zero-fill into a 128B-swizzled K-major tile), then the epilogue; warp 4: TMEM allocator and the
// single-thread tcgen05.mma issuer.  smem ring of `stages` x {A 16 KB, B N*128 B}; full/empty mbarriers;
// accumulator (128 lanes x N fp32 columns) lives in TMEM and is read back with tcgen05.ld.
// Epilogue: optional residual/accumulate input, GroupNorm partial sums (sum, sum of squares per
// (sample, group)) reduced with a warp butterfly and one atomicAdd per value, fp16 (or fp32) store.
#include "common.cuh"
#include "ops.cuh"
#include "tmap.cuh"

namespace pnvo {

static constexpr int kTileM = 128;
static constexpr int kTileK = 64;  // fp16 elements per K stage = one 128-byte swizzle row
static constexpr int kProducerThreads = 128;

template <int V>
__device__ __forceinline__ void warp_reduce_scatter(float* a, int lane) {
  int off = 16;
#pragma unroll
  for (int cnt = V / 2; cnt >= 1; cnt >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < cnt; ++i) {
      const float send = up ? a[i] : a[i + cnt];
      const float recv = __shfl_xor_sync(0xffffffffu, send, off);
      a[i] = (up ? a[i + cnt] : a[i]) + recv;
    }
    off >>= 1;
  }
  for (; off >= 1; off >>= 1) a[0] += __shfl_xor_sync(0xffffffffu, a[0], off);
}

// per-thread (sum, sumsq) of CPG-wide channel groups inside a 32-column chunk, reduced over the rows of
// the warp that belong to the same sample, then one atomicAdd per value.
template <int CPG>
__device__ __forceinline__ void chunk_stats(const float* v, int sample, bool row_valid, double* stats, int G,
                                            int group0, int lane) {
  constexpr int NG = 32 / CPG;
  constexpr int V = 2 * NG;
  float s[V];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    float a = 0.f, q = 0.f;
#pragma unroll
    for (int c = 0; c < CPG; ++c) {
      const float x = v[g * CPG + c];
      a += x;
      q = fmaf(x, x, q);
    }
    s[2 * g] = a;
    s[2 * g + 1] = q;
  }
  int pending = row_valid ? sample : 0x7fffffff;
  while (true) {
    const int cur = __reduce_min_sync(0xffffffffu, pending);
    if (cur == 0x7fffffff) break;
    const bool mine = (pending == cur);
    float a[V];
#pragma unroll
    for (int i = 0; i < V; ++i) a[i] = mine ? s[i] : 0.f;
    warp_reduce_scatter<V>(a, lane);
    constexpr int kLanesPerVal = 32 / V;
    if ((lane & (kLanesPerVal - 1)) == 0) {
      const int e = lane / kLanesPerVal;
      atomicAdd(stats + (static_cast<int64_t>(cur) * G + group0) * 2 + e, static_cast<double>(a[0]));
    }
    if (mine) pending = 0x7fffffff;
  }
}

__global__ void __launch_bounds__(160) conv_igemm_kernel(const ConvArgs p, const __grid_constant__ ConvTmaps tm) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_full[8];
  __shared__ __align__(8) uint64_t s_empty[8];
  __shared__ __align__(8) uint64_t s_accum;
  __shared__ uint32_t s_tmem;
  __shared__ short2 s_tap[64];   // tap -> (r, s)
  __shared__ int s_tapoff[64];   // tap -> element offset of the tap relative to the row pointer

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int stages = p.stages;
  const int m0 = blockIdx.x * kTileM;
  const int n0 = blockIdx.y * p.N;
  const uint32_t row_bytes = static_cast<uint32_t>(p.chunk_k) * 2;  // 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B, Cin = 32)
  const uint32_t a_bytes = kTileM * row_bytes;
  const uint32_t b_bytes = static_cast<uint32_t>(p.N) * row_bytes;
  const bool split = p.x_lo != nullptr;  // stage = {A, B} or {A, B, A_lo, B_lo}
  const uint32_t pair_bytes = a_bytes + b_bytes;
  const uint32_t stage_bytes = split ? 2 * pair_bytes : pair_bytes;
  // dynamic smem base rounded up to 1024 B (128B swizzle atoms repeat every 1024 B)
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;

  for (int i = tid; i < p.R * p.S; i += blockDim.x) {
    const int r = i / p.S, sx = i % p.S;
    s_tap[i] = make_short2(r, sx);
    // valid taps sit on the stride lattice, where (ohb + r) / div == floor(ohb / div) + ceil(r / div)
    s_tapoff[i] = (((r + p.div - 1) / p.div) * p.IW + (sx + p.div - 1) / p.div) * p.Cin;
  }
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&s_full[s]), p.tma ? 1 : kProducerThreads);
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

  if (warp < 4 && p.tma) {
    // ================================ TMA producer ================================
    // One elected thread per CTA: per K chunk (one filter tap x chunk_k channels) one im2col TMA brings
    // the 128-pixel x chunk_k activation tile (halo / batch tail zero-filled by the TMA unit) and one
    // tiled TMA the N x chunk_k weight tile, both landing 64/128-byte swizzled exactly as tcgen05.mma
    // reads them.  No per-element address arithmetic on the SM.
    if (warp == 0 && elect_one()) {
      tma_prefetch_desc(&tm.a);
      tma_prefetch_desc(&tm.b);
      if (split) {
        tma_prefetch_desc(&tm.a_lo);
        tma_prefetch_desc(&tm.b_lo);
      }
      const int ohw = p.OH * p.OW;
      const int n_img = m0 / ohw;
      const int rem = m0 - n_img * ohw;
      const int p0 = rem / p.OW, q0 = rem - p0 * p.OW;
      const int w0 = q0 * p.mul - p.pad_w, h0 = p0 * p.mul - p.pad;
      const int nkb = p.nkb;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % stages;
        if (kb >= stages) mbar_wait(smem_u32(&s_empty[s]), ((kb / stages) & 1) ^ 1);
        const uint32_t bar = smem_u32(&s_full[s]);
        const uint32_t sA = smem_base + s * stage_bytes;
        const int kf = kb * p.chunk_k;
        const short2 rs = s_tap[kf >> p.cin_log2];
        mbar_arrive_expect_tx(bar, stage_bytes);
        tma_load_im2col_4d(sA, &tm.a, bar, kf & p.cmask, w0, h0, n_img, static_cast<uint16_t>(rs.y),
                           static_cast<uint16_t>(rs.x));
        tma_load_2d(sA + a_bytes, &tm.b, bar, kf, n0);
        if (split) {  // residual planes of both operands, same boxes (stage = {A, B, A_lo, B_lo})
          tma_load_im2col_4d(sA + pair_bytes, &tm.a_lo, bar, kf & p.cmask, w0, h0, n_img, static_cast<uint16_t>(rs.y),
                             static_cast<uint16_t>(rs.x));
          tma_load_2d(sA + pair_bytes + a_bytes, &tm.b_lo, bar, kf, n0);
        }
      }
    }
  }
  if (warp < 4) {
    const int ohw = p.OH * p.OW;
    if (!p.tma) {
    // ============================ im2col / weight producers ============================
    // Thread t owns 16-byte chunk j = t % 8 of rows (t / 8) + 16 i, i = 0..7: the 8 lanes of a row
    // fetch one contiguous 128-byte line (or two 64-byte runs when Cin = 32), so a warp-level cp.async
    // touches 4 lines instead of 32 (the SM -> L2 request rate, not bandwidth, bounded the first version).
    const int j = tid & 7, rsub = tid >> 3;
    const int dshift = (p.div == 2) ? 1 : 0, dmask = p.div - 1;
    const int n_taps = p.R * p.S;
    // Per row: pointer to x[b, floor(ohb/div), floor(owb/div), 0] and a bit mask of the taps that fall
    // inside the input (and on the stride lattice for dgrad).  Per stage and chunk the address is then
    // row pointer + tap offset (smem LUT) and the predicate one bit test -- the first version spent
    // ~40 instructions per 16-byte chunk on this and was issue-bound.
    const __half* rowptr[8];
    unsigned long long tmask[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int mi = m0 + rsub + 16 * i;
      rowptr[i] = p.x;
      tmask[i] = 0ull;
      if (mi < p.M) {
        const int bi = mi / ohw;
        const int rem = mi - bi * ohw;
        const int oh = rem / p.OW;
        const int ohb = oh * p.mul - p.pad;
        const int owb = (rem - oh * p.OW) * p.mul - p.pad_w;
        rowptr[i] = p.x + (static_cast<int64_t>(bi) * p.IH * p.IW +
                           static_cast<int64_t>(ohb >> dshift) * p.IW + (owb >> dshift)) * p.Cin;
        unsigned long long mk = 0ull;
        for (int t = 0; t < n_taps; ++t) {
          const short2 rs = s_tap[t];
          const int th = ohb + rs.x, tw = owb + rs.y;
          const bool ok = (th >= 0) && (tw >= 0) && (((th | tw) & dmask) == 0) && ((th >> dshift) < p.IH) &&
                          ((tw >> dshift) < p.IW);
          mk |= static_cast<unsigned long long>(ok ? 1u : 0u) << t;
        }
        tmask[i] = mk;
      }
    }
    const uint32_t t_off = static_cast<uint32_t>((rsub >> 3) * 1024 + (rsub & 7) * 128 + ((j ^ (rsub & 7)) << 4));
    const int nb_rows = p.N >> 4;  // B rows handled by this thread: rsub + 16 i
    const __half* wrow = p.w + static_cast<int64_t>(n0 + rsub) * p.w_ld + j * 8;
    const int64_t w_step = static_cast<int64_t>(16) * p.w_ld;
    const int64_t lo_dx = split ? (p.x_lo - p.x) : 0, lo_dw = split ? (p.w_lo - p.w) : 0;

    auto issue_stage = [&](int kb) {
      const int s = kb % stages;
      const uint32_t sA = smem_base + s * stage_bytes + t_off;
      const uint32_t sB = sA + a_bytes;
      const int kf = kb * kTileK + j * 8;
      int tap = 0, toff = 0;
      unsigned long long kvalid = 0ull;
      if (kf < p.K) {
        tap = kf >> p.cin_log2;
        toff = s_tapoff[tap] + (kf & p.cmask);
        kvalid = 1ull;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool ok = ((tmask[i] >> tap) & kvalid) != 0ull;
        cp_async_16(sA + i * 2048, rowptr[i] + toff, ok ? 16u : 0u);
      }
      const __half* wsrc = wrow + kb * kTileK;
      for (int i = 0; i < nb_rows; ++i) cp_async_16(sB + i * 2048, wsrc + i * w_step, 16u);
      if (split) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool ok = ((tmask[i] >> tap) & kvalid) != 0ull;
          cp_async_16(sA + pair_bytes + i * 2048, rowptr[i] + toff + lo_dx, ok ? 16u : 0u);
        }
        for (int i = 0; i < nb_rows; ++i) cp_async_16(sB + pair_bytes + i * 2048, wsrc + i * w_step + lo_dw, 16u);
      }
    };

    const int nkb = p.nkb;
    const int la = p.lookahead;  // committed cp.async groups kept in flight per thread
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % stages;
      if (kb >= stages) mbar_wait(smem_u32(&s_empty[s]), ((kb / stages) & 1) ^ 1);
      issue_stage(kb);
      cp_async_commit();
      if (kb >= la) {
        cp_async_wait_dyn(la);
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&s_full[(kb - la) % stages]));
      }
    }
    for (int rem = min(la, nkb) - 1; rem >= 0; --rem) {  // drain
      cp_async_wait_dyn(rem);
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&s_full[(nkb - 1 - rem) % stages]));
    }

    }  // !tma
    const int row = tid;  // epilogue: thread == accumulator row == output pixel m0 + tid
    const int m = m0 + row;
    const bool row_valid = m < p.M;
    const int b = row_valid ? m / ohw : 0;
    // where the row is stored: densely at pixel m, or scattered onto a stride-o_mul lattice of a larger tensor
    int64_t mo = m;
    bool st_valid = row_valid;
    if (p.o_mul) {
      const int rem = m - b * ohw;
      const int oh = rem / p.OW;
      const int gh = oh * p.o_mul + p.o_off_h, gw = (rem - oh * p.OW) * p.o_mul + p.o_off_w;
      st_valid = row_valid && gh < p.o_H && gw < p.o_W;
      mo = (static_cast<int64_t>(b) * p.o_H + gh) * p.o_W + gw;
    }

    // ==================================== epilogue ====================================
    mbar_wait(smem_u32(&s_accum), 0);
    tc_fence_after();
    const int sample = b;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const int n_chunks = p.N >> 5;
    for (int ch = 0; ch < n_chunks + ((p.N & 31) ? 1 : 0); ++ch) {
      float v[32];
      tmem_ld32(t_row + ch * 32, v);
      tmem_ld_wait();
      const int col0 = n0 + ch * 32;  // first output channel of this chunk
      if ((p.N & 31) && ch == n_chunks) {
        // N = 16 (mod 32): upper 16 columns of the last chunk were never written by the MMA
#pragma unroll
        for (int i = 16; i < 32; ++i) v[i] = 0.f;
      }
      if (p.add && st_valid) {
        const __half* ap = p.add + mo * p.ldo + col0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (col0 + q * 8 < p.n_store) {
            const uint4 u = *reinterpret_cast<const uint4*>(ap + q * 8);
            const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __half22float2(h2[e]);
              v[q * 8 + 2 * e] += f.x;
              v[q * 8 + 2 * e + 1] += f.y;
            }
          }
        }
      }
      if (p.stats) {
        const int g0 = col0 / p.cpg;
        switch (p.cpg) {
          case 2: chunk_stats<2>(v, sample, row_valid, p.stats, p.G, g0, lane); break;
          case 4: chunk_stats<4>(v, sample, row_valid, p.stats, p.G, g0, lane); break;
          case 8: chunk_stats<8>(v, sample, row_valid, p.stats, p.G, g0, lane); break;
          case 16: chunk_stats<16>(v, sample, row_valid, p.stats, p.G, g0, lane); break;
          default: chunk_stats<32>(v, sample, row_valid, p.stats, p.G, g0, lane); break;  // cpg >= 32
        }
      }
      if (st_valid) {
        if (p.y_lo) {  // value + residual fp16 planes
          __half* yh = reinterpret_cast<__half*>(p.y) + mo * p.ldo + col0;
          __half* yl = p.y_lo + mo * p.ldo + col0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (col0 + q * 8 < p.n_store) {
              uint4 hi, lo;
              split8(v + 8 * q, hi, lo);
              *reinterpret_cast<uint4*>(yh + q * 8) = hi;
              *reinterpret_cast<uint4*>(yl + q * 8) = lo;
            }
          }
        } else if (p.out_fp32) {
          float* yp = reinterpret_cast<float*>(p.y) + mo * p.ldo + col0;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (col0 + q * 4 < p.n_store)
              *reinterpret_cast<float4*>(yp + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
        } else {
          __half* yp = reinterpret_cast<__half*>(p.y) + mo * p.ldo + col0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (col0 + q * 8 < p.n_store) {
              uint4 u;
              __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(v[q * 8 + 2 * e], v[q * 8 + 2 * e + 1]);
              *reinterpret_cast<uint4*>(yp + q * 8) = u;
            }
          }
        }
      }
    }
    tc_fence_before();
  } else {
    // ================================ MMA issuer (warp 4) ================================
    // One thread issues every MMA: descriptors are (lo, hi) words with a loop-invariant hi, the stage base is a counter
    // and the accumulate flag is a compile-time constant except for the very first MMA (the first version rebuilt two to
    // four 64-bit descriptors per stage and tested the flag per MMA: at N <= 64 the thread, not the pipe, set the pace).
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16(kTileM, p.N, 0, 0);
      const int nkb = p.nkb;
      const uint64_t d0 = umma_desc(0, 16, 8 * row_bytes, row_bytes);
      const uint32_t hi = static_cast<uint32_t>(d0 >> 32);
      const uint32_t base = static_cast<uint32_t>(d0) + (smem_base >> 4);
      const uint32_t a16 = a_bytes >> 4, pair16 = pair_bytes >> 4, stage16 = stage_bytes >> 4;
      const bool k4 = (p.chunk_k >> 4) == 4;   // 4 (64-element chunks) or 2 (32-element chunks) K steps per stage
      uint32_t slot = 0, phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(smem_u32(&s_full[slot]), phase);
        tc_fence_after();
        const uint32_t lo = base + slot * stage16;
        tc_mma_f16_lohi(tmem_base, lo, lo + a16, hi, idesc, kb != 0 ? 1u : 0u);
        tc_mma_f16_lohi_c<true>(tmem_base, lo + 2, lo + a16 + 2, hi, idesc);
        if (k4) {
          tc_mma_f16_lohi_c<true>(tmem_base, lo + 4, lo + a16 + 4, hi, idesc);
          tc_mma_f16_lohi_c<true>(tmem_base, lo + 6, lo + a16 + 6, hi, idesc);
        }
        if (split) {  // + x_lo * w + x * w_lo (the lo * lo term is below fp32 resolution)
          const uint32_t alo = lo + pair16, blo = lo + pair16 + a16;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            tc_mma_f16_lohi_c<true>(tmem_base, alo + 2 * k, lo + a16 + 2 * k, hi, idesc);
            tc_mma_f16_lohi_c<true>(tmem_base, lo + 2 * k, blo + 2 * k, hi, idesc);
          }
          if (k4) {
#pragma unroll
            for (int k = 2; k < 4; ++k) {
              tc_mma_f16_lohi_c<true>(tmem_base, alo + 2 * k, lo + a16 + 2 * k, hi, idesc);
              tc_mma_f16_lohi_c<true>(tmem_base, lo + 2 * k, blo + 2 * k, hi, idesc);
            }
          }
        }
        tc_commit(smem_u32(&s_empty[slot]));  // frees the smem stage once these MMAs have read it
        if (++slot == static_cast<uint32_t>(stages)) {
          slot = 0;
          phase ^= 1u;
        }
      }
      tc_commit(smem_u32(&s_accum));  // accumulator complete -> epilogue
    }
    __syncwarp();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

static int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

int conv_plan(ConvArgs& a) {
  PNVO_REQUIRE(a.Cin >= 8 && a.Cin % 8 == 0 && (a.R * a.S == 1 || (a.Cin & (a.Cin - 1)) == 0),
               "conv: Cin=%d must be a multiple of 8 (power of two unless 1x1)", a.Cin);
  PNVO_REQUIRE(a.R * a.S <= 64, "conv: filter %dx%d too large", a.R, a.S);
  PNVO_REQUIRE(a.div == 1 || a.div == 2, "conv: div=%d unsupported", a.div);
  PNVO_REQUIRE(a.n_total % 16 == 0, "conv: padded Cout=%d must be a multiple of 16", a.n_total);
  if (a.R * a.S == 1) {  // 1x1: tap is always 0, any channel count
    a.cin_log2 = 30;
    a.cmask = 0x3fffffff;
  } else {
    a.cin_log2 = 0;
    while ((1 << a.cin_log2) < a.Cin) ++a.cin_log2;
    a.cmask = a.Cin - 1;
  }
  a.M = a.B * a.OH * a.OW;
  a.K = a.R * a.S * a.Cin;
  // TMA im2col path: unit "dilation" (div == 1), square filter / symmetric padding, channels in chunks of 32 or 64
  const bool split = a.x_lo != nullptr;
  PNVO_REQUIRE((a.x_lo != nullptr) == (a.w_lo != nullptr), "conv: split mode needs both x_lo and w_lo");
  PNVO_REQUIRE(!a.y_lo || !a.out_fp32, "conv: y_lo (value + residual fp16 output) excludes out_fp32");
  PNVO_REQUIRE(!a.o_mul || (!a.stats && a.o_H > 0 && a.o_W > 0), "conv: strided output excludes GroupNorm statistics");
  // (asymmetric padding: any filter shape, the high-side corner of the TMA bounding box is given separately)
  a.tma = (a.div == 1 && a.Cin % 32 == 0 && (a.asym || (a.R == a.S && a.pad == a.pad_w)) && a.force_generic != 1) ? 1 : 0;
  a.chunk_k = (a.tma && a.Cin % 64 != 0) ? 32 : kTileK;
  a.nkb = ceil_div(a.K, a.chunk_k);
  PNVO_REQUIRE(a.w_ld >= a.nkb * a.chunk_k, "conv: packed weight row stride %d < padded K %d", a.w_ld, a.nkb * a.chunk_k);
  // N tile: whole Cout when <= 256, else the largest divisor <= 256 that is a multiple of 32
  int N = a.n_total;
  const int n_max = split ? 128 : 256;  // split stages hold four tiles: keep >= 2 stages in shared memory
  if (N > n_max) {
    N = n_max;
    while (a.n_total % N) N -= 32;
  }
  a.N = N;
  a.tmem_cols = pow2_cols(N);
  if (a.stats) {
    PNVO_REQUIRE(a.cpg >= 2 && (a.cpg & (a.cpg - 1)) == 0, "conv: channels/group=%d must be a power of two >= 2", a.cpg);
    PNVO_REQUIRE(a.cpg <= 32 || N % 32 == 0, "conv: bad group tiling");
    PNVO_REQUIRE(N % 32 == 0, "conv: GroupNorm statistics need Cout %% 32 == 0 (got %d)", N);
  }
  const int stage_bytes = (kTileM + N) * a.chunk_k * 2 * (split ? 2 : 1);
  // two CTAs per SM when a >= 4-stage ring fits in ~100 KB, else one CTA with a deep ring
  int stages = std::min(8, (100 * 1024) / stage_bytes);
  if (stages < 4) stages = std::min(6, (198 * 1024) / stage_bytes);
  if (a.nkb < stages) stages = std::max(2, a.nkb);
  PNVO_REQUIRE(stages >= 2, "conv: stage of %d bytes does not fit twice in shared memory", stage_bytes);
  a.stages = stages;
  a.lookahead = std::max(1, std::min(4, stages - 2));
  a.smem_bytes = stages * stage_bytes + 1024;
  a.grid_x = ceil_div(a.M, kTileM);
  a.grid_y = a.n_total / N;
  return 0;
}

int conv_launch(ConvArgs a, cudaStream_t st) {
  if (a.force_generic == 0 && a.x && a.w && a.y && conv_direct1_supported(a)) return conv_direct1_launch(a, st);
  if (a.force_generic == 0 && a.x && a.w && a.y && conv_raster_supported(a)) return conv_raster_launch(a, st);
  if (a.force_generic == 0 && a.x && a.w && a.y && conv_raster128_supported(a)) return conv_raster128_launch(a, st);
  if (conv_plan(a)) return -1;
  PNVO_REQUIRE(a.x && a.w && a.y, "conv: null pointer");
  if (a.M == 0) return 0;
  static int max_smem_set = 0;
  if (a.smem_bytes > max_smem_set) {
    cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    max_smem_set = 200 * 1024;
  }
  alignas(64) ConvTmaps tm;
  memset(&tm, 0, sizeof(tm));
  if (a.tma) {
    const int ph_hi = a.asym ? a.pad_hi_h : a.pad, pw_hi = a.asym ? a.pad_hi_w : a.pad_w;
    if (tmap_im2col(&tm.a, a.x, a.B, a.IH, a.IW, a.Cin, a.R, a.S, a.mul, a.pad, a.chunk_k, kTileM, 0, a.pad_w, ph_hi, pw_hi)) return -1;
    if (tmap_tiled2d(&tm.b, a.w, a.n_total, a.w_ld, a.w_ld, a.N, a.chunk_k)) return -1;
    if (a.x_lo) {
      if (tmap_im2col(&tm.a_lo, a.x_lo, a.B, a.IH, a.IW, a.Cin, a.R, a.S, a.mul, a.pad, a.chunk_k, kTileM, 0, a.pad_w, ph_hi, pw_hi)) return -1;
      if (tmap_tiled2d(&tm.b_lo, a.w_lo, a.n_total, a.w_ld, a.w_ld, a.N, a.chunk_k)) return -1;
    }
  }
  conv_igemm_kernel<<<dim3(a.grid_x, a.grid_y), 160, a.smem_bytes, st>>>(a, tm);
  count_launch();
  return check_launch("conv_igemm");
}

}  // namespace pnvo
