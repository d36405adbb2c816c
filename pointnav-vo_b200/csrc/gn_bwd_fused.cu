// GroupNorm(+ReLU) backward in ONE pass over HBM (resnet.py:39-55 backward; replaces the reduce + apply pair
// of norm_pool.cu whenever one sample fits in the shared memory of a thread-block cluster).
//
//   dy   = g * [relu_ref > 0] * g_scale
//   S1_c = sum_hw dy,  S2_c = sum_hw dy * xhat          (per sample, per channel; xhat = (x - mean_g) * rstd_g)
//   dx   = rstd_g * (gamma_c * dy - (sum_{c in g} gamma_c S1_c + xhat * sum_{c in g} gamma_c S2_c) / cnt)
//
// A cluster of `cs` CTAs (1, 2, 4 or 8) owns one sample; CTA r owns a contiguous slice of its [HW][C] fp16 data.
// One thread brings the slices of g and x into shared memory with bulk async copies (cp.async.bulk, mbarrier
// completion); the ReLU reference, needed once for the mask, is prefetched into registers: each byte crosses HBM ONCE.  The per-channel sums are reduced in a fixed
// order inside the CTA (shared memory) and across the cluster (distributed shared memory), so the result is
// bit-reproducible and needs neither atomics nor a pre-zeroed buffer; dx (and the masked gradient of the identity
// branch) are then produced from the staged slices.  HBM traffic: 3 reads + 1-2 writes per element instead of
// 6 + 1-2; registers stay small, so several CTAs per SM overlap their load / reduce / store phases.
#include <cooperative_groups.h>

#include <cstdlib>
#include "common.cuh"
#include "elem.cuh"

namespace cg = cooperative_groups;

namespace pnvo {

static constexpr int kGbfThreads = 256;

__device__ __forceinline__ void unpack8(const uint4& u, float* v) {
  const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = __half22float2(h2[e]);
    v[2 * e] = f.x;
    v[2 * e + 1] = f.y;
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// slice: vectors (16 B) per CTA, multiple of C/8; slice_bytes = slice * 16 rounded up to 128.
// The ReLU reference is NOT staged: every thread prefetches its <= kGbfMaxIt vectors of it into registers while the
// bulk copies of g / x fly (it is needed once, for the mask), which keeps the CTA at two slices of shared memory.
static constexpr int kGbfMaxIt = 9;

// XF32: the raw conv output x is fp32 (split-precision forward): its slice takes two fp16-sized slices of shared memory
template <bool XF32>
__device__ __forceinline__ void gbf_load_x(const uint4* s_x, int i, float* x) {
  if (XF32) {
    const float4* xf = reinterpret_cast<const float4*>(s_x) + 2 * i;
    const float4 a = xf[0], b = xf[1];
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
    x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  } else {
    unpack8(s_x[i], x);
  }
}

template <bool XF32>
__global__ void __launch_bounds__(kGbfThreads) gn_bwd_fused_kernel(const GnBwdArgs a, const int n8, const int cs,
                                                                   const int slice, const int slice_bytes,
                                                                   const int n_slots) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ __align__(8) uint64_t s_bar;
  const int C = a.C, G = a.G;
  uint4* s_g = reinterpret_cast<uint4*>(s_raw);
  uint4* s_x = reinterpret_cast<uint4*>(s_raw + slice_bytes);
  float* s_part = reinterpret_cast<float*>(s_raw + (XF32 ? 3 : 2) * slice_bytes);  // [n_slots][17] reduction scratch
  float* s_loc = s_part + n_slots * 17;                               // [2C] CTA sums (read by peers)
  float* s_tot = s_loc + 2 * C;                                       // [2C] sample totals
  float* s_coef = s_tot + 2 * C;                                      // [3C] A_c, B_c, C_c
  float* s_mr = s_coef + 3 * C;                                       // [2C] mean_c, rstd_c
  float* s_gamma = s_mr + 2 * C;                                      // [C]

  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = blockIdx.x;  // gridDim.x == cluster size
  const int b = blockIdx.y;
  const int c8 = C >> 3;
  const int cc = (tid % c8) * 8;  // fixed 8-channel chunk of this thread (slice and 256 are multiples of c8)
  const int first = rank * slice;
  const int len = max(0, min(slice, n8 - first));
  const int64_t base = static_cast<int64_t>(b) * n8 + first;

  if (tid == 0) {
    mbar_init(smem_u32(&s_bar), 1);
    fence_mbar_init();
    const uint32_t bytes = static_cast<uint32_t>(len) * 16u;
    const uint32_t bar = smem_u32(&s_bar);
    mbar_arrive_expect_tx(bar, bytes * (XF32 ? 3u : 2u));
    if (len > 0) {
      bulk_g2s(smem_u32(s_g), reinterpret_cast<const uint4*>(a.g) + base, bytes, bar);
      if (XF32) bulk_g2s(smem_u32(s_x), reinterpret_cast<const uint4*>(a.x) + 2 * base, bytes * 2u, bar);
      else bulk_g2s(smem_u32(s_x), reinterpret_cast<const uint4*>(a.x) + base, bytes, bar);
    }
  }
  // ReLU reference of this thread's vectors -> registers (in flight together with the bulk copies)
  const bool has_y = a.relu_ref != nullptr;
  uint4 yq[kGbfMaxIt];
  if (has_y) {
    const uint4* __restrict__ yp = reinterpret_cast<const uint4*>(a.relu_ref) + base;
#pragma unroll
    for (int k = 0; k < kGbfMaxIt; ++k) {
      const int i = tid + k * kGbfThreads;
      yq[k] = (i < len) ? __ldg(yp + i) : make_uint4(0, 0, 0, 0);
    }
  }
  // per-channel mean / rstd of this sample (forward statistics) and gamma while the copies fly
  for (int c = tid; c < C; c += kGbfThreads) {
    const int g = c / a.cpg;
    const double s = a.stats[(static_cast<int64_t>(b) * G + g) * 2], q = a.stats[(static_cast<int64_t>(b) * G + g) * 2 + 1];
    const double m = s / static_cast<double>(a.cnt);
    const double var = fmax(q / static_cast<double>(a.cnt) - m * m, 0.0);
    s_mr[2 * c] = static_cast<float>(m);
    s_mr[2 * c + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(a.eps)));
    s_gamma[c] = (c < a.C_real) ? a.gamma[c] : 0.f;
  }
  __syncthreads();  // barrier initialised before anybody waits on it
  mbar_wait(smem_u32(&s_bar), 0);

  // thread partials: sum g, sum g*x over this thread's vectors; masked g goes back to shared memory
  float sd[8], sgx[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) sd[e] = sgx[e] = 0.f;
#pragma unroll
  for (int k = 0; k < kGbfMaxIt; ++k) {
    const int i = tid + k * kGbfThreads;
    if (i < len) {
      uint4 gq = s_g[i];
      if (has_y) {
        // mask: keep g where the saved post-ReLU output is > 0 (sign / zero test on the fp16 bits)
        uint32_t* gw = reinterpret_cast<uint32_t*>(&gq);
        const uint32_t* yw = reinterpret_cast<const uint32_t*>(&yq[k]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t y = yw[e];
          const uint32_t lo = ((y & 0x7fffu) != 0u && (y & 0x8000u) == 0u) ? 0x0000ffffu : 0u;
          const uint32_t hi = ((y & 0x7fff0000u) != 0u && (y & 0x80000000u) == 0u) ? 0xffff0000u : 0u;
          gw[e] &= (lo | hi);
        }
        s_g[i] = gq;
      }
      float g[8], x[8];
      unpack8(gq, g);
      gbf_load_x<XF32>(s_x, i, x);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        sd[e] += g[e];
        sgx[e] = fmaf(g[e], x[e], sgx[e]);
      }
    }
  }
  // lanes l, l + c8, l + 2 c8 ... of a warp own the same channels: fold them with shuffles first (fixed order)
  if (c8 < 32) {
    for (int off = 16; off >= c8; off >>= 1) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        sd[e] += __shfl_xor_sync(0xffffffffu, sd[e], off);
        sgx[e] += __shfl_xor_sync(0xffffffffu, sgx[e], off);
      }
    }
  }
  const int slot = (c8 < 32) ? ((lane < c8) ? warp * c8 + lane : -1) : tid;
  if (slot >= 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s_part[slot * 17 + 2 * e] = sd[e];
      s_part[slot * 17 + 2 * e + 1] = sgx[e];
    }
  }
  __syncthreads();
  // CTA totals in a fixed order: output o = chunk * 16 + v  <->  (channel = chunk*8 + v/2, which = v & 1);
  // the slots holding chunk q are q, q + c8, q + 2 c8, ...
  const int per_chunk = n_slots / c8;
  for (int o = tid; o < 2 * C; o += kGbfThreads) {
    const int q = o >> 4, v = o & 15;
    float t = 0.f;
    for (int j = 0; j < per_chunk; ++j) t += s_part[(q + j * c8) * 17 + v];
    s_loc[o] = t;
  }
  cluster.sync();
  for (int o = tid; o < 2 * C; o += kGbfThreads) {
    float t = 0.f;
    for (int r = 0; r < cs; ++r) t += *(cluster.map_shared_rank(s_loc + o, r));
    s_tot[o] = t;
  }
  cluster.sync();  // every CTA has finished reading its peers' s_loc
  for (int c = tid; c < C; c += kGbfThreads) {
    // (sum g, sum g*x) -> (S1, S2 = rstd * (sum g*x - mean * sum g)), scaled by g_scale
    const float mean = s_mr[2 * c], rstd = s_mr[2 * c + 1];
    const float S1 = s_tot[2 * c] * a.g_scale;
    const float S2 = rstd * (s_tot[2 * c + 1] - mean * s_tot[2 * c]) * a.g_scale;
    s_loc[2 * c] = S1;  // s_loc is free again: per-channel (S1, S2)
    s_loc[2 * c + 1] = S2;
    if (rank == 0) {
      a.sums[(static_cast<int64_t>(b) * C + c) * 2] = S1;
      a.sums[(static_cast<int64_t>(b) * C + c) * 2 + 1] = S2;
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += kGbfThreads) {
    const int g = c / a.cpg;
    float T1 = 0.f, T2 = 0.f;
    for (int k = g * a.cpg; k < (g + 1) * a.cpg && k < a.C_real; ++k) {
      const float gm = s_gamma[k];
      T1 = fmaf(gm, s_loc[2 * k], T1);
      T2 = fmaf(gm, s_loc[2 * k + 1], T2);
    }
    const float mean = s_mr[2 * c], rstd = s_mr[2 * c + 1];
    const float k1 = rstd * T1 / a.cnt, k2 = rstd * T2 / a.cnt;
    // dx = rstd*gamma*dy - k1 - xhat*k2 = A*g + Bc + Cc*x
    s_coef[3 * c] = rstd * s_gamma[c] * a.g_scale;
    s_coef[3 * c + 1] = -k1 + mean * rstd * k2;
    s_coef[3 * c + 2] = -rstd * k2;
  }
  __syncthreads();
  float cA[8], cB[8], cC[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    cA[e] = s_coef[3 * (cc + e)];
    cB[e] = s_coef[3 * (cc + e) + 1];
    cC[e] = s_coef[3 * (cc + e) + 2];
  }
  uint4* __restrict__ dxp = reinterpret_cast<uint4*>(a.dx) + base;
  uint4* __restrict__ dyp = a.dy_out ? reinterpret_cast<uint4*>(a.dy_out) + base : nullptr;
  const bool scale_dy = a.g_scale != 1.f;
  for (int i = tid; i < len; i += kGbfThreads) {
    const uint4 gq = s_g[i];
    float g[8], x[8];
    unpack8(gq, g);
    gbf_load_x<XF32>(s_x, i, x);
    uint4 u;
    __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float d0 = fmaf(cA[2 * e], g[2 * e], fmaf(cC[2 * e], x[2 * e], cB[2 * e]));
      const float d1 = fmaf(cA[2 * e + 1], g[2 * e + 1], fmaf(cC[2 * e + 1], x[2 * e + 1], cB[2 * e + 1]));
      h2[e] = __floats2half2_rn(d0, d1);
    }
    dxp[i] = u;
    if (dyp) {
      if (scale_dy) {
#pragma unroll
        for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(g[2 * e] * a.g_scale, g[2 * e + 1] * a.g_scale);
        dyp[i] = u;
      } else {
        dyp[i] = gq;
      }
    }
  }
}

struct GbfPlan {
  int cs, slice, slice_bytes, n_slots, smem;
};
static bool gbf_plan(const GnBwdArgs& a, GbfPlan& p) {
  if (a.C % 8 != 0) return false;
  const int c8 = a.C / 8;
  if (c8 > kGbfThreads || (kGbfThreads % c8) != 0) return false;
  const int64_t n8 = static_cast<int64_t>(a.HW) * c8;
  if (n8 <= 0 || n8 > (1 << 24)) return false;
  const int n_slots = c8 < 32 ? 8 * c8 : kGbfThreads;
  // smallest cluster whose CTAs stage <= 100 KB (two CTAs per SM): fewer, fatter CTAs and smaller clusters won over four
  // 48 KB CTAs per SM (cluster launch / cluster.sync latency, not bandwidth, bounded the small layers); a thread holds at
  // most kGbfMaxIt vectors of the ReLU reference in registers
  for (int cs = 1; cs <= 8; cs <<= 1) {
    int slice = static_cast<int>(ceil_div64(n8, cs));
    slice = ceil_div(slice, c8) * c8;
    if (slice > kGbfMaxIt * kGbfThreads) continue;
    const int slice_bytes = (slice * 16 + 127) & ~127;
    const int smem = (a.x_fp32 ? 3 : 2) * slice_bytes + (n_slots * 17 + 10 * a.C) * 4;
    // measured (B = 256, sum over the 20 launches of a ResNet-18 step): limit 48 KB 0.79 ms, 75 KB 0.70 ms, 100 KB 0.66 ms
    static const int lim_kb = getenv("PNVO_GBF_SMEM_KB") ? atoi(getenv("PNVO_GBF_SMEM_KB")) : 100;
    // (8 CTAs: up to 112 KB, still two CTAs per SM -- the fp32 raw outputs of layer1 need 100.1 KB)
    if (smem <= lim_kb * 1024 || (cs == 8 && smem <= 112 * 1024)) {
      p = GbfPlan{cs, slice, slice_bytes, n_slots, smem};
      return true;
    }
  }
  return false;
}

// 1 when the fused kernel can take this shape (the caller falls back to reduce + apply otherwise)
int gn_bwd_fused_supported(const GnBwdArgs& a) {
  GbfPlan p;
  return gbf_plan(a, p) ? 1 : 0;
}

int gn_bwd_fused_launch(const GnBwdArgs& a, int B, cudaStream_t st) {
  GbfPlan p;
  PNVO_REQUIRE(gbf_plan(a, p), "gn_bwd_fused: unsupported shape C=%d HW=%d", a.C, a.HW);
  if (B <= 0 || a.HW <= 0) return 0;
  const int n8 = a.HW * (a.C / 8);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(gn_bwd_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    cudaFuncSetAttribute(gn_bwd_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    attr = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.cs, B, 1);
  cfg.blockDim = dim3(kGbfThreads, 1, 1);
  cfg.dynamicSmemBytes = p.smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = p.cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  const cudaError_t e = a.x_fp32 ? cudaLaunchKernelEx(&cfg, gn_bwd_fused_kernel<true>, a, n8, p.cs, p.slice, p.slice_bytes, p.n_slots)
                                 : cudaLaunchKernelEx(&cfg, gn_bwd_fused_kernel<false>, a, n8, p.cs, p.slice, p.slice_bytes, p.n_slots);
  if (e != cudaSuccess) {
    set_error("gn_bwd_fused: %s", cudaGetErrorString(e));
    return -2;
  }
  count_launch();
  return check_launch("gn_bwd_fused");
}

}  // namespace pnvo
