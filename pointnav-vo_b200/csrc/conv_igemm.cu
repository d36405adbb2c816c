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

namespace pnvo {

static constexpr int kTileM = 128;
static constexpr int kTileK = 64;  // fp16 elements per K stage = one 128-byte swizzle row
static constexpr int kProducerThreads = 128;
static constexpr int kLookahead = 2;  // cp.async groups in flight per producer thread

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
__device__ __forceinline__ void chunk_stats(const float* v, int sample, bool row_valid, float* stats, int G,
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
      atomicAdd(stats + (static_cast<int64_t>(cur) * G + group0) * 2 + e, a[0]);
    }
    if (mine) pending = 0x7fffffff;
  }
}

__global__ void __launch_bounds__(160) conv_igemm_kernel(const ConvArgs p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_full[8];
  __shared__ __align__(8) uint64_t s_empty[8];
  __shared__ __align__(8) uint64_t s_accum;
  __shared__ uint32_t s_tmem;
  __shared__ short2 s_tap[128];  // tap -> (r, s)

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int stages = p.stages;
  const int m0 = blockIdx.x * kTileM;
  const int n0 = blockIdx.y * p.N;
  const uint32_t a_bytes = kTileM * 128;
  const uint32_t b_bytes = static_cast<uint32_t>(p.N) * 128;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  // dynamic smem base rounded up to 1024 B (128B swizzle atoms repeat every 1024 B)
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;

  for (int i = tid; i < p.R * p.S; i += blockDim.x) s_tap[i] = make_short2(i / p.S, i % p.S);
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&s_full[s]), kProducerThreads);
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

  if (warp < 4) {
    // ============================ im2col / weight producers ============================
    const int row = tid;  // A-tile row == output pixel m0 + row
    const int m = m0 + row;
    const bool row_valid = m < p.M;
    int b = 0, oh = 0, ow = 0;
    if (row_valid) {
      const int ohw = p.OH * p.OW;
      b = m / ohw;
      const int rem = m - b * ohw;
      oh = rem / p.OW;
      ow = rem - oh * p.OW;
    }
    const int ohb = oh * p.mul - p.pad, owb = ow * p.mul - p.pad_w;
    const __half* __restrict__ xb = p.x + static_cast<int64_t>(b) * p.IH * p.IW * p.Cin;
    const int dmask = p.div - 1, dshift = (p.div == 2) ? 1 : 0;
    const int cmask = p.cmask;
    const uint32_t a_row_off = static_cast<uint32_t>((row >> 3) * 1024 + (row & 7) * 128);
    const int rx = row & 7;

    auto issue_stage = [&](int kb) {
      const int s = kb % stages;
      const uint32_t sA = smem_base + s * stage_bytes;
      const uint32_t sB = sA + a_bytes;
      // ---- A: 8 chunks of 8 channels for this row ----
      const __half* src = nullptr;
      bool ok = false;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kf = kb * kTileK + j * 8;
        if (j == 0 || (kf & cmask) == 0) {
          ok = false;
          if (row_valid && kf < p.K) {
            const int tap = kf >> p.cin_log2;
            const short2 rs = s_tap[tap];
            const int th = ohb + rs.x, tw = owb + rs.y;
            if (th >= 0 && tw >= 0 && ((th | tw) & dmask) == 0) {
              const int ih = th >> dshift, iw = tw >> dshift;
              if (ih < p.IH && iw < p.IW) {
                ok = true;
                src = xb + (static_cast<int64_t>(ih) * p.IW + iw) * p.Cin + (kf & cmask);
              }
            }
          }
        } else {
          src += 8;
        }
        cp_async_16(sA + a_row_off + ((j ^ rx) << 4), ok ? static_cast<const void*>(src) : static_cast<const void*>(p.x),
                    ok ? 16u : 0u);
      }
      // ---- B: weight rows (always in range: packed weights are zero-padded to [n_total][w_ld]) ----
      for (int r = row; r < p.N; r += kProducerThreads) {
        const __half* wsrc = p.w + static_cast<int64_t>(n0 + r) * p.w_ld + kb * kTileK;
        const uint32_t dst = sB + static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128);
        const int rr = r & 7;
#pragma unroll
        for (int j = 0; j < 8; ++j) cp_async_16(dst + ((j ^ rr) << 4), wsrc + j * 8, 16u);
      }
    };

    const int nkb = p.nkb;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % stages;
      if (kb >= stages) mbar_wait(smem_u32(&s_empty[s]), ((kb / stages) & 1) ^ 1);
      issue_stage(kb);
      cp_async_commit();
      if (kb >= kLookahead) {
        cp_async_wait<kLookahead>();
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&s_full[(kb - kLookahead) % stages]));
      }
    }
    // drain
    if (nkb >= 2) {
      cp_async_wait<1>();
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&s_full[(nkb - 2) % stages]));
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    mbar_arrive(smem_u32(&s_full[(nkb - 1) % stages]));

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
      if (p.add && row_valid) {
        const __half* ap = p.add + static_cast<int64_t>(m) * p.ldo + col0;
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
      if (row_valid) {
        if (p.out_fp32) {
          float* yp = reinterpret_cast<float*>(p.y) + static_cast<int64_t>(m) * p.ldo + col0;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (col0 + q * 4 < p.n_store)
              *reinterpret_cast<float4*>(yp + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
        } else {
          __half* yp = reinterpret_cast<__half*>(p.y) + static_cast<int64_t>(m) * p.ldo + col0;
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
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(kTileM, p.N, 0, 0);
      const int nkb = p.nkb;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % stages;
        mbar_wait(smem_u32(&s_full[s]), (kb / stages) & 1);
        tc_fence_after();
        const uint32_t sA = smem_base + s * stage_bytes;
        const uint64_t adesc = umma_desc_sw128(sA, 16, 1024);
        const uint64_t bdesc = umma_desc_sw128(sA + a_bytes, 16, 1024);
#pragma unroll
        for (int k = 0; k < kTileK / 16; ++k) {
          // +32 bytes (16 fp16 of K) inside the 128-byte swizzle row -> +2 in the (addr >> 4) field
          tc_mma_f16(tmem_base, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                     (kb | k) != 0 ? 1u : 0u);
        }
        tc_commit(smem_u32(&s_empty[s]));  // frees the smem stage once these MMAs have read it
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
  PNVO_REQUIRE(a.R * a.S <= 128, "conv: filter %dx%d too large", a.R, a.S);
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
  a.nkb = ceil_div(a.K, kTileK);
  PNVO_REQUIRE(a.w_ld >= a.nkb * kTileK, "conv: packed weight row stride %d < padded K %d", a.w_ld, a.nkb * kTileK);
  // N tile: whole Cout when <= 256, else the largest divisor <= 256 that is a multiple of 32
  int N = a.n_total;
  if (N > 256) {
    N = 256;
    while (a.n_total % N) N -= 32;
  }
  a.N = N;
  a.tmem_cols = pow2_cols(N);
  if (a.stats) {
    PNVO_REQUIRE(a.cpg >= 2 && (a.cpg & (a.cpg - 1)) == 0, "conv: channels/group=%d must be a power of two >= 2", a.cpg);
    PNVO_REQUIRE(a.cpg <= 32 || N % 32 == 0, "conv: bad group tiling");
    PNVO_REQUIRE(N % 32 == 0, "conv: GroupNorm statistics need Cout %% 32 == 0 (got %d)", N);
  }
  const int stage_bytes = kTileM * 128 + N * 128;
  int stages = (N <= 64) ? 4 : 3;
  if (a.nkb < stages) stages = a.nkb < 2 ? 2 : a.nkb;
  a.stages = stages;
  a.smem_bytes = stages * stage_bytes + 1024;
  a.grid_x = ceil_div(a.M, kTileM);
  a.grid_y = a.n_total / N;
  return 0;
}

int conv_launch(ConvArgs a, cudaStream_t st) {
  if (conv_plan(a)) return -1;
  PNVO_REQUIRE(a.x && a.w && a.y, "conv: null pointer");
  if (a.M == 0) return 0;
  static int max_smem_set = 0;
  if (a.smem_bytes > max_smem_set) {
    cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    max_smem_set = 200 * 1024;
  }
  conv_igemm_kernel<<<dim3(a.grid_x, a.grid_y), 160, a.smem_bytes, st>>>(a);
  count_launch();
  return check_launch("conv_igemm");
}

}  // namespace pnvo
