// HBM-bound elementwise / reduction kernels around the tensor-core convolutions (sm_100a):
// input assembly + RunningMeanAndVar (vo_cnn.py:110-176, running_mean_and_var.py:22-63), GroupNorm apply
// with fused ReLU / residual / 3x3-s2 max-pool (resnet.py:39-55,165-168), GroupNorm backward, weight
// (un)packing, the tiny regression head (vo_cnn.py:216-227) and a flat-bucket Adam.
// All activation tensors are NHWC fp16 with the channel count padded to a multiple of 8, so every
// thread moves 16-byte vectors.
#include <cstdlib>
#include "common.cuh"
#include "elem.cuh"

namespace pnvo {

// ------------------------------------------------------------------------------------------------
// zero fill
// ------------------------------------------------------------------------------------------------
__global__ void zero_kernel(uint4* __restrict__ p, int64_t n16, unsigned char* tail, int ntail) {
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = i0; i < n16; i += stride) p[i] = make_uint4(0, 0, 0, 0);
  if (i0 < ntail) tail[i0] = 0;
}
int zero_launch(void* p, int64_t bytes, cudaStream_t st) {
  if (bytes <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(p) & 15) != 0) {  // small unaligned slices of the flat gradient bucket
    const cudaError_t e = cudaMemsetAsync(p, 0, static_cast<size_t>(bytes), st);
    if (e != cudaSuccess) {
      set_error("zero: %s", cudaGetErrorString(e));
      return -2;
    }
    return 0;
  }
  const int64_t n16 = bytes / 16;
  const int ntail = static_cast<int>(bytes - n16 * 16);
  const int blocks = static_cast<int>(std::min<int64_t>(std::max<int64_t>(1, ceil_div64(n16, 256)), 148 * 8));
  zero_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<uint4*>(p), n16, reinterpret_cast<unsigned char*>(p) + n16 * 16,
                                      ntail);
  count_launch();
  return check_launch("zero");
}

// ------------------------------------------------------------------------------------------------
// input assembly: up to 4 NHWC fp32 sources whose channels are [prev | cur] -> one NHWC fp16 tensor
// with channels [prev of every source ..., cur of every source ..., zero pad], normalised per channel.
// A block stages 256 pixels of every source in shared memory with coalesced float4 loads (the
// per-pixel channel runs of the sources are 8..80 bytes, too short for direct vector access), then
// one thread per pixel picks its channels through a constant LUT and stores 16-byte vectors.
// ------------------------------------------------------------------------------------------------
static constexpr int kAsmPix = 256;

__device__ __forceinline__ void stage_sources(const AssembleArgs& a, int64_t pix0, int npix, float* s_src,
                                              int* s_base) {
  int off = 0;
  for (int t = 0; t < a.n_src; ++t) {
    const int nch = a.nch[t];
    const int64_t f0 = pix0 * nch;           // first float of this block in source t
    const int nfl = npix * nch;              // floats to stage
    const float* __restrict__ g = a.src[t] + f0;
    float* d = s_src + off;
    if (threadIdx.x == 0) s_base[t] = off;
    // pix0 is a multiple of 256 so f0*4 is 16-byte aligned for every nch
    const int n4 = nfl >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x)
      reinterpret_cast<float4*>(d)[i] = __ldg(reinterpret_cast<const float4*>(g) + i);
    for (int i = (n4 << 2) + threadIdx.x; i < nfl; i += blockDim.x) d[i] = __ldg(g + i);
    off += kAsmPix * nch;
  }
}

__device__ __forceinline__ void pick_pixel(const AssembleArgs& a, const float* s_src, const int* s_base, int lp,
                                           float* v) {
#pragma unroll
  for (int c = 0; c < kMaxInC; ++c) {
    v[c] = 0.f;
    if (c < a.C) {
      const int t = a.src_idx[c];
      v[c] = s_src[s_base[t] + lp * a.nch[t] + a.src_ch[c]] * a.pre_scale[t];
    }
  }
}

__global__ void __launch_bounds__(kAsmPix) assemble_kernel(const AssembleArgs a) {
  extern __shared__ __align__(16) float s_src[];
  __shared__ float s_scale[kMaxInC], s_shift[kMaxInC];
  __shared__ int s_base[4];
  if (threadIdx.x < kMaxInC) {
    const int c = threadIdx.x;
    s_scale[c] = (c < a.C) ? (a.scale ? a.scale[c] : 1.f) : 0.f;
    s_shift[c] = (c < a.C && a.shift) ? a.shift[c] : 0.f;
  }
  const int64_t pix0 = static_cast<int64_t>(blockIdx.x) * kAsmPix;
  const int npix = static_cast<int>(min(static_cast<int64_t>(kAsmPix), a.n_pix - pix0));
  stage_sources(a, pix0, npix, s_src, s_base);
  __syncthreads();
  const int lp = threadIdx.x;
  if (lp >= npix) return;
  float v[kMaxInC];
  pick_pixel(a, s_src, s_base, lp, v);
  int64_t opix = pix0 + lp;
  if (a.out_pitch > 0) opix = (opix / a.row_w) * a.out_pitch + (opix % a.row_w) + 3;  // zero halo left of the image
  __half* __restrict__ out = a.out + opix * a.Cpad;
#pragma unroll
  for (int q = 0; q < kMaxInC / 8; ++q) {
    if (q * 8 < a.Cpad) {
      uint4 u;
      __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = q * 8 + 2 * e;
        h2[e] = __floats2half2_rn(fmaf(v[c], s_scale[c], s_shift[c]), fmaf(v[c + 1], s_scale[c + 1], s_shift[c + 1]));
      }
      *reinterpret_cast<uint4*>(out + q * 8) = u;
      if (a.out_lo) {  // split-fp16 residual plane
        uint4 ul;
        __half2* l2 = reinterpret_cast<__half2*>(&ul);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = q * 8 + 2 * e;
          const float x0 = fmaf(v[c], s_scale[c], s_shift[c]), x1 = fmaf(v[c + 1], s_scale[c + 1], s_shift[c + 1]);
          l2[e] = __floats2half2_rn(x0 - __low2float(h2[e]), x1 - __high2float(h2[e]));
        }
        *reinterpret_cast<uint4*>(a.out_lo + opix * a.Cpad + q * 8) = ul;
      }
    }
  }
}

// per-channel sum / sum of squares of the assembled (un-normalised) input.  Each source tensor is streamed as a
// flat float4 array: with nch channels per pixel the channel of a float4 lane repeats every lcm(nch,4)/4
// vectors, so a thread whose vector index advances in multiples of that period keeps a FIXED channel per lane
// and accumulates in 8 registers (no gathers, fully coalesced 16-byte loads).  Block partials are combined in
// shared memory in a fixed order and added to the fp64 global accumulator with one atomic per channel.
static constexpr int kStatsThreads = 240;   // multiple of every period (1, 3, 5 for nch = 2, 6, 20)
static constexpr int kStatsVecPerThread = 64;
__global__ void __launch_bounds__(kStatsThreads) input_stats_kernel(const AssembleArgs a, double* __restrict__ stats,
                                                                    int t, int period, int64_t n_vec) {
  __shared__ float s_part[kStatsThreads][8];
  const int nch = a.nch[t];
  const float pre = a.pre_scale[t];
  const float4* __restrict__ src = reinterpret_cast<const float4*>(a.src[t]);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kStatsThreads * kStatsVecPerThread + threadIdx.x;
#pragma unroll 8
  for (int k = 0; k < kStatsVecPerThread; ++k) {
    const int64_t i = base + static_cast<int64_t>(k) * kStatsThreads;
    if (i < n_vec) {
      float4 v = __ldg(src + i);
      v.x *= pre; v.y *= pre; v.z *= pre; v.w *= pre;
      acc[0] += v.x; acc[1] = fmaf(v.x, v.x, acc[1]);
      acc[2] += v.y; acc[3] = fmaf(v.y, v.y, acc[3]);
      acc[4] += v.z; acc[5] = fmaf(v.z, v.z, acc[5]);
      acc[6] += v.w; acc[7] = fmaf(v.w, v.w, acc[7]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s_part[threadIdx.x][i] = acc[i];
  __syncthreads();
  // thread c < nch sums lane e of every thread whose (vector index * 4 + e) % nch == c
  const int c = threadIdx.x;
  if (c < nch) {
    double ds = 0.0, dq = 0.0;
    for (int r = 0; r < period; ++r) {        // residue class of the vector index modulo the period
      for (int e = 0; e < 4; ++e) {
        if ((r * 4 + e) % nch != c) continue;
        float fs = 0.f, fq = 0.f;
        for (int th = r; th < kStatsThreads; th += period) {
          fs += s_part[th][2 * e];
          fq += s_part[th][2 * e + 1];
        }
        ds += fs;
        dq += fq;
      }
    }
    // output channel of (source t, channel c): prev half -> first block, cur half -> second block
    int oc = -1;
    for (int o = 0; o < a.C; ++o)
      if (a.src_idx[o] == t && a.src_ch[o] == c) oc = o;
    if (oc >= 0) {
      atomicAdd(stats + 2 * oc, ds);
      atomicAdd(stats + 2 * oc + 1, dq);
    }
  }
}

// RunningMeanAndVar (running_mean_and_var.py:22-63): optional Chan merge of the batch statistics into
// the running buffers (training), then scale = 1/sqrt(max(var, 1e-2)), shift = -mean*scale.
// stats: [2C] fp64 (sum, sumsq interleaved) + stats[2*kMaxInC] = sample count (all ranks), n_per_sample pixels.
__global__ void rmv_update_kernel(const double* __restrict__ stats, double n_batch, double pix_per_sample,
                                  float* __restrict__ mean, float* __restrict__ var, float* __restrict__ count, int C,
                                  int update, int have_rmv, float* __restrict__ scale, float* __restrict__ shift) {
  const int c = threadIdx.x;
  const float cnt = have_rmv ? *count : 0.f;
  __syncthreads();
  if (c < C) {
    float m = have_rmv ? mean[c] : 0.f, v = have_rmv ? var[c] : 1.f;
    if (update && have_rmv) {
      const double n = n_batch * pix_per_sample;
      const double mu = stats[2 * c] / n;
      double va = stats[2 * c + 1] / n - mu * mu;
      if (va < 0) va = 0;
      const float new_mean = static_cast<float>(mu), new_var = static_cast<float>(va);
      const float new_count = static_cast<float>(n_batch);
      const float m_a = v * cnt, m_b = new_var * new_count;
      const float d = new_mean - m;
      const float M2 = m_a + m_b + d * d * cnt * new_count / (cnt + new_count);
      v = M2 / (cnt + new_count);
      m = (cnt * m + new_count * new_mean) / (cnt + new_count);
      mean[c] = m;
      var[c] = v;
    }
    if (have_rmv) {
      const float sd = sqrtf(fmaxf(v, 1e-2f));
      scale[c] = 1.0f / sd;
      shift[c] = -m / sd;
    } else {
      scale[c] = 1.f;
      shift[c] = 0.f;
    }
  }
  if (c == 0 && update && have_rmv) *count = cnt + static_cast<float>(n_batch);
}

int assemble_launch(const AssembleArgs& a, cudaStream_t st) {
  PNVO_REQUIRE(a.C <= kMaxInC && a.Cpad <= kMaxInC && a.Cpad % 8 == 0 && a.C <= a.Cpad, "assemble: bad channel counts");
  PNVO_REQUIRE(a.n_src >= 1 && a.n_src <= 4, "assemble: n_src");
  int tot_ch = 0;
  for (int t = 0; t < a.n_src; ++t) tot_ch += a.nch[t];
  PNVO_REQUIRE(tot_ch * kAsmPix * 4 <= 48 * 1024, "assemble: %d source channels exceed the staging tile", tot_ch);
  if (a.n_pix <= 0) return 0;
  int tot = 0;
  for (int t = 0; t < a.n_src; ++t) tot += a.nch[t];
  assemble_kernel<<<static_cast<int>(ceil_div64(a.n_pix, kAsmPix)), kAsmPix, tot * kAsmPix * sizeof(float), st>>>(a);
  count_launch();
  return check_launch("assemble");
}
int input_stats_launch(const AssembleArgs& a, double* stats, cudaStream_t st) {
  PNVO_REQUIRE(a.C <= kMaxInC, "input_stats: too many channels");
  if (a.n_pix <= 0) return 0;
  for (int t = 0; t < a.n_src; ++t) {
    const int nch = a.nch[t];
    int period = 1;
    while ((period * 4) % nch != 0) ++period;  // lcm(nch, 4) / 4
    PNVO_REQUIRE(kStatsThreads % period == 0, "input_stats: %d channels per pixel not supported", nch);
    const int64_t n_float = a.n_pix * nch;
    PNVO_REQUIRE(n_float % 4 == 0, "input_stats: source %d size not a multiple of 4 floats", t);
    const int64_t n_vec = n_float / 4;
    // block size 240 and 64 vectors per thread keep (block start) a multiple of every period
    const int blocks = static_cast<int>(ceil_div64(n_vec, static_cast<int64_t>(kStatsThreads) * kStatsVecPerThread));
    input_stats_kernel<<<blocks, kStatsThreads, 0, st>>>(a, stats, t, period, n_vec);
    count_launch();
  }
  return check_launch("input_stats");
}
int rmv_update_launch(const double* stats, double n_batch, double pix_per_sample, float* mean, float* var,
                      float* count, int C, int update, int have_rmv, float* scale, float* shift, cudaStream_t st) {
  PNVO_REQUIRE(C <= kMaxInC, "rmv_update: too many channels");
  rmv_update_kernel<<<1, 64, 0, st>>>(stats, n_batch, pix_per_sample, mean, var, count, C, update, have_rmv, scale,
                                      shift);
  count_launch();
  return check_launch("rmv_update");
}

// ------------------------------------------------------------------------------------------------
// Zero-insertion upsampling by 2: dst[b, 2i, 2j, :] = src[b, i, j, :], every other element of dst = 0.
// The data gradient of a stride-2 convolution is the stride-1 convolution of this tensor with the flipped
// weights, which runs on the TMA / raster kernels instead of the divisibility-testing generic producer; for the
// 1x1 / stride-2 downsample branch it scatters the compact 1x1 result back to the input resolution.
// ------------------------------------------------------------------------------------------------
// One thread per PAIR of destination pixels (w = 2j, 2j + 1) x 8 channels of one sample (blockIdx.y): one load (even rows
// only), two 16-byte stores, 32-bit index arithmetic (the per-destination-vector version with 64-bit div / mod ran at
// 2 TB/s of write traffic).
__global__ void __launch_bounds__(256) upsample2_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int OH,
                                                        int OW, int IH, int IW, int c8) {
  const int b = blockIdx.y;
  const int BW = (IW + 1) >> 1;
  const int items = IH * BW * c8;
  const uint4* __restrict__ s = src + static_cast<int64_t>(b) * OH * OW * c8;
  uint4* __restrict__ d = dst + static_cast<int64_t>(b) * IH * IW * c8;
  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < items; it += gridDim.x * blockDim.x) {
    const int q = it % c8;
    const int t = it / c8;
    const int j = t % BW;
    const int h = t / BW;
    uint4 v = make_uint4(0, 0, 0, 0);
    if ((h & 1) == 0 && (h >> 1) < OH && j < OW) v = __ldg(s + ((h >> 1) * OW + j) * c8 + q);
    const int o = (h * IW + 2 * j) * c8 + q;
    d[o] = v;
    if (2 * j + 1 < IW) d[o + c8] = make_uint4(0, 0, 0, 0);
  }
}
int upsample2_launch(const __half* src, __half* dst, int B, int OH, int OW, int IH, int IW, int C, cudaStream_t st) {
  PNVO_REQUIRE(src && dst && C % 8 == 0, "upsample2: bad arguments");
  const int64_t items = static_cast<int64_t>(IH) * ((IW + 1) / 2) * (C / 8);
  PNVO_REQUIRE(static_cast<int64_t>(IH) * IW * (C / 8) < (1ll << 30), "upsample2: sample too large");
  if (items <= 0 || B <= 0) return 0;
  const int gx = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div64(items, 256 * 2), (148 * 16) / std::max(1, std::min(B, 148 * 16)) + 1)));
  upsample2_kernel<<<dim3(gx, B), 256, 0, st>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), OH, OW,
                                                IH, IW, C / 8);
  count_launch();
  return check_launch("upsample2");
}

// ------------------------------------------------------------------------------------------------
// 2x2 average pool of fp32 NHWC sources -> fp16 NHWC (policy encoder, resnet_policy.py:146-168)
// ------------------------------------------------------------------------------------------------
// out32 (nullable): the pooled values in fp32, [B, OH, OW, ld32] at channel offset coff -- the input of the RunningMeanAndVar
// statistics / normalisation ops when the encoder normalises its visual inputs (resnet_policy.py:170)
__global__ void __launch_bounds__(256) avgpool2_kernel(const float* __restrict__ src, int B, int H, int W, int C,
                                                       float pre_scale, __half* __restrict__ out, int Cpad, int coff,
                                                       __half* __restrict__ out_lo, float* __restrict__ out32, int ld32) {
  const int OH = H / 2, OW = W / 2;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(B) * OH * OW;
  if (idx >= total) return;
  const int ow = static_cast<int>(idx % OW);
  const int oh = static_cast<int>((idx / OW) % OH);
  const int b = static_cast<int>(idx / (static_cast<int64_t>(OW) * OH));
  const float* p = src + ((static_cast<int64_t>(b) * H + 2 * oh) * W + 2 * ow) * C;
  const bool div255 = pre_scale == 1.f / 255.f;
  for (int c = 0; c < C; ++c) {
    // the reference scales (rgb / 255) before pooling; F.avg_pool2d sums the window then multiplies by 1/4
    float q[4] = {p[c], p[C + c], p[static_cast<int64_t>(W) * C + c], p[static_cast<int64_t>(W) * C + C + c]};
    if (div255) {  // rgb / 255.0: a true division, as the reference does (resnet_policy.py:155)
#pragma unroll
      for (int k = 0; k < 4; ++k) q[k] = __fdiv_rn(q[k], 255.f);
    } else if (pre_scale != 1.f) {
#pragma unroll
      for (int k = 0; k < 4; ++k) q[k] *= pre_scale;
    }
    const float s = ((q[0] + q[1]) + q[2]) + q[3];
    const float v = s * 0.25f;
    if (out32) out32[idx * ld32 + coff + c] = v;
    if (out) {
      const __half h = __float2half_rn(v);
      out[idx * Cpad + coff + c] = h;
      if (out_lo) out_lo[idx * Cpad + coff + c] = __float2half_rn(v - __half2float(h));  // split-fp16 residual plane
    }
  }
}
int avgpool2_launch(const float* src, int B, int H, int W, int C, float pre_scale, __half* out, int Cpad, int coff,
                    cudaStream_t st, __half* out_lo, float* out32, int ld32) {
  const int64_t total = static_cast<int64_t>(B) * (H / 2) * (W / 2);
  if (total <= 0) return 0;
  PNVO_REQUIRE(out || out32, "avgpool2: no output");
  avgpool2_kernel<<<static_cast<int>(ceil_div64(total, 256)), 256, 0, st>>>(src, B, H, W, C, pre_scale, out, Cpad, coff,
                                                                            out_lo, out32, ld32);
  count_launch();
  return check_launch("avgpool2");
}

// ------------------------------------------------------------------------------------------------
// GroupNorm helpers.  stats[b][g] = (sum, sumsq) over the group's cpg_real*HW elements (fp32 partials
// accumulated by the conv epilogue).  Per block (one sample) the affine per channel is put in smem:
//   y = x * a_c + b_c,  a_c = gamma_c * rstd_g,  b_c = beta_c - mean_g * a_c
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void group_mean_rstd(const double* stats, int b, int G, int g, float cnt, float eps,
                                                float& mean, float& rstd) {
  // fp64 sums: mean / variance are formed in fp64 (no cancellation in E[x^2] - E[x]^2) and rounded once
  const double s = stats[(static_cast<int64_t>(b) * G + g) * 2], q = stats[(static_cast<int64_t>(b) * G + g) * 2 + 1];
  const double m = s / static_cast<double>(cnt);
  const double var = fmax(q / static_cast<double>(cnt) - m * m, 0.0);
  mean = static_cast<float>(m);
  rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

__device__ __forceinline__ void load8(const void* base, int64_t idx8, int is_fp32, float* v) {
  if (is_fp32) {
    const float4* p = reinterpret_cast<const float4*>(base) + idx8 * 2;
    const float4 a = __ldg(p), b = __ldg(p + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(base) + idx8);
    const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __half22float2(h2[e]);
      v[2 * e] = f.x;
      v[2 * e + 1] = f.y;
    }
  }
}
__device__ __forceinline__ void store8h(__half* base, int64_t idx8, const float* v) {
  uint4 u;
  __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
  reinterpret_cast<uint4*>(base)[idx8] = u;
}
// residual plane of the split-fp16 representation: v - fp16(v), itself rounded to fp16
__device__ __forceinline__ void store8h_lo(__half* base, int64_t idx8, const float* v) {
  float r[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) r[e] = v[e] - __half2float(__float2half_rn(v[e]));
  store8h(base, idx8, r);
}

// y = [relu]( GN(x) [+ res] ).  SPLIT: residual fp16 planes of res / y (split-precision forward) -- a template parameter,
// not a run-time branch.  This kernel lives on occupancy (40 registers, 6 CTAs per SM): every variant that added
// registers was measured slower over the 32 launches of a training step (570 us): run-time split branches in the loop
// body 658 us, first-iteration loads hoisted above the statistics prologue 678 us, 2 / 4 explicitly batched vectors per
// thread 797 / 1023 us, __launch_bounds__(256, 8) (32 registers, spills) 712 us.
template <bool SPLIT>
__global__ void __launch_bounds__(256) gn_apply_kernel(const GnArgs a) {
  extern __shared__ float s_ab[];  // a_c [C], b_c [C]
  const int b = blockIdx.y;
  const int C = a.C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, rstd;
    group_mean_rstd(a.stats, b, a.G, c / a.cpg, a.cnt, a.eps, mean, rstd);
    const float ga = (c < a.C_real) ? a.gamma[c] * rstd : 0.f;
    s_ab[c] = ga;
    s_ab[C + c] = (c < a.C_real) ? a.beta[c] - mean * ga : 0.f;
  }
  __syncthreads();
  const int c8 = C >> 3;
  const int64_t per_sample = static_cast<int64_t>(a.HW) * c8;
  const int64_t base = static_cast<int64_t>(b) * per_sample;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < per_sample;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cc = static_cast<int>(i % c8) * 8;
    float v[8];
    load8(a.x, base + i, a.x_fp32, v);
    if (SPLIT && a.x_lo) {
      float r[8];
      load8(a.x_lo, base + i, 0, r);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += r[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e], s_ab[cc + e], s_ab[C + cc + e]);
    if (a.res) {
      float r[8];
      load8(a.res, base + i, 0, r);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += r[e];
      if (SPLIT && a.res_lo) {
        load8(a.res_lo, base + i, 0, r);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] += r[e];
      }
    }
    if (a.relu) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
    }
    store8h(a.y, base + i, v);
    if (SPLIT && a.y_lo) store8h_lo(a.y_lo, base + i, v);
  }
}

// GN + ReLU + MaxPool 3x3 / stride 2 / pad 1; also records the arg-max tap (0..8) for the backward pass
__global__ void __launch_bounds__(256) gn_pool_kernel(const GnArgs a, int H, int W, int PH, int PW,
                                                      uint8_t* __restrict__ argmax) {
  extern __shared__ float s_ab[];
  const int b = blockIdx.y;
  const int C = a.C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, rstd;
    group_mean_rstd(a.stats, b, a.G, c / a.cpg, a.cnt, a.eps, mean, rstd);
    const float ga = (c < a.C_real) ? a.gamma[c] * rstd : 0.f;
    s_ab[c] = ga;
    s_ab[C + c] = (c < a.C_real) ? a.beta[c] - mean * ga : 0.f;
  }
  __syncthreads();
  const int c8 = C >> 3;
  const int per_sample = PH * PW * c8;
  const int64_t in_base = static_cast<int64_t>(b) * H * W * c8;
  const int64_t out_base = static_cast<int64_t>(b) * per_sample;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_sample; i += gridDim.x * blockDim.x) {
    const int q = i % c8;
    const int cc = q * 8;
    const int pp = i / c8;
    const int pw = pp % PW;
    const int ph = pp / PW;
    // all nine taps are fetched before any of them is used (fp16 inputs; the fp32 path keeps the simple loop)
    float best[8];
    int arg[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; arg[e] = 0; }
    float ga[8], gb[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { ga[e] = s_ab[cc + e]; gb[e] = s_ab[C + cc + e]; }
    if (!a.x_fp32 && !a.x_lo) {
      uint4 tapv[9];
      bool ok[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int h = 2 * ph - 1 + t / 3, w = 2 * pw - 1 + t % 3;
        ok[t] = h >= 0 && h < H && w >= 0 && w < W;
        tapv[t] = make_uint4(0, 0, 0, 0);
        if (ok[t]) tapv[t] = __ldg(reinterpret_cast<const uint4*>(a.x) + in_base + (static_cast<int64_t>(h) * W + w) * c8 + q);
      }
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        if (!ok[t]) continue;
        const __half2* h2 = reinterpret_cast<const __half2*>(&tapv[t]);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float v = (e & 1) ? __high2float(h2[e >> 1]) : __low2float(h2[e >> 1]);
          const float y = fmaxf(fmaf(v, ga[e], gb[e]), 0.f);
          if (y > best[e]) { best[e] = y; arg[e] = t; }
        }
      }
    } else if (a.x_lo && !a.x_fp32) {
      // value + residual planes (split-precision plans): the six loads of a window row are issued together
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int h = 2 * ph - 1 + r;
        if (h < 0 || h >= H) continue;
        uint4 hv[3], lv[3];
        bool ok[3];
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int w = 2 * pw - 1 + s;
          ok[s] = w >= 0 && w < W;
          hv[s] = lv[s] = make_uint4(0, 0, 0, 0);
          if (ok[s]) {
            const int64_t o = in_base + (static_cast<int64_t>(h) * W + w) * c8 + q;
            hv[s] = __ldg(reinterpret_cast<const uint4*>(a.x) + o);
            lv[s] = __ldg(reinterpret_cast<const uint4*>(a.x_lo) + o);
          }
        }
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          if (!ok[s]) continue;
          const __half2* h2 = reinterpret_cast<const __half2*>(&hv[s]);
          const __half2* l2 = reinterpret_cast<const __half2*>(&lv[s]);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float v = ((e & 1) ? __high2float(h2[e >> 1]) : __low2float(h2[e >> 1])) +
                            ((e & 1) ? __high2float(l2[e >> 1]) : __low2float(l2[e >> 1]));
            const float y = fmaxf(fmaf(v, ga[e], gb[e]), 0.f);
            if (y > best[e]) { best[e] = y; arg[e] = r * 3 + s; }
          }
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int h = 2 * ph - 1 + r;
        if (h < 0 || h >= H) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int w = 2 * pw - 1 + s;
          if (w < 0 || w >= W) continue;
          float v[8];
          load8(a.x, in_base + (static_cast<int64_t>(h) * W + w) * c8 + q, a.x_fp32, v);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float y = fmaxf(fmaf(v[e], ga[e], gb[e]), 0.f);
            if (y > best[e]) { best[e] = y; arg[e] = r * 3 + s; }
          }
        }
      }
    }
    store8h(a.y, out_base + i, best);
    if (a.y_lo) store8h_lo(a.y_lo, out_base + i, best);
    if (argmax) {
      uint2 u;
      u.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
      u.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
      reinterpret_cast<uint2*>(argmax)[out_base + i] = u;
    }
  }
}

// gradient of MaxPool(ReLU(.)) routed back to the GN output: dy[h,w,c] = sum of g[ph,pw,c] over the
// windows whose arg-max is (h,w) and whose pooled value is > 0
__global__ void __launch_bounds__(256) pool_bwd_kernel(const __half* __restrict__ g, const __half* __restrict__ pooled,
                                                       const uint8_t* __restrict__ argmax, __half* __restrict__ dy,
                                                       int H, int W, int PH, int PW, int C) {
  const int b = blockIdx.y;
  const int c8 = C >> 3;
  const int per_sample = H * W * c8;
  const int64_t out_base = static_cast<int64_t>(b) * per_sample;
  const int64_t p_base = static_cast<int64_t>(b) * PH * PW * c8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_sample; i += gridDim.x * blockDim.x) {
    const int q = i % c8;
    const int pix = i / c8;
    const int w = pix % W;
    const int h = pix / W;
    // the (<= 2 x 2) pooling windows that contain (h, w): ph in {h/2, (h+1)/2}, pw likewise; all loads issued up front
    const int ph0 = h >> 1, ph1 = (h + 1) >> 1, pw0 = w >> 1, pw1 = (w + 1) >> 1;
    uint2 am[4];
    uint4 gv[4], pv[4];
    bool ok[4];
    int tap[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int ph = (k & 2) ? ph1 : ph0, pw = (k & 1) ? pw1 : pw0;
      ok[k] = ph < PH && pw < PW && !((k & 2) && ph1 == ph0) && !((k & 1) && pw1 == pw0);
      tap[k] = (h - (2 * ph - 1)) * 3 + (w - (2 * pw - 1));
      am[k] = make_uint2(0xffffffffu, 0xffffffffu);
      gv[k] = pv[k] = make_uint4(0, 0, 0, 0);
      if (ok[k]) {
        const int64_t pi = p_base + (static_cast<int64_t>(ph) * PW + pw) * c8 + q;
        am[k] = __ldg(reinterpret_cast<const uint2*>(argmax) + pi);
        gv[k] = __ldg(reinterpret_cast<const uint4*>(g) + pi);
        pv[k] = __ldg(reinterpret_cast<const uint4*>(pooled) + pi);
      }
    }
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __half2* g2 = reinterpret_cast<const __half2*>(&gv[k]);
      const __half2* p2 = reinterpret_cast<const __half2*>(&pv[k]);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int t = ((e < 4 ? am[k].x : am[k].y) >> (8 * (e & 3))) & 0xff;
        const float gval = (e & 1) ? __high2float(g2[e >> 1]) : __low2float(g2[e >> 1]);
        const float pval = (e & 1) ? __high2float(p2[e >> 1]) : __low2float(p2[e >> 1]);
        if (ok[k] && t == tap[k] && pval > 0.f) acc[e] += gval;
      }
    }
    store8h(dy, out_base + i, acc);
  }
}

// Same gradient, one thread per 2 x 2 block of outputs x 8 channels.  Rows 2i / 2i+1 only ever belong to the window rows
// i (taps r = 1 / 2) and i+1 (tap r = 0, odd row only), so the block needs exactly the four windows (i..i+1, j..j+1) and
// every window is fetched by four blocks instead of by its nine outputs: 12 loads + 4 stores per 4 outputs instead of
// 27 + 4 (the per-output gather was bound by load/store issue at 1.6 TB/s of DRAM traffic, not by HBM).
__global__ void __launch_bounds__(256) pool_bwd2x2_kernel(const __half* __restrict__ g, const __half* __restrict__ pooled,
                                                          const uint8_t* __restrict__ argmax, __half* __restrict__ dy,
                                                          int H, int W, int PH, int PW, int C) {
  const int b = blockIdx.y;
  const int c8 = C >> 3;
  const int BH = (H + 1) >> 1, BW = (W + 1) >> 1;
  const int per_sample = BH * BW * c8;
  const int64_t out_base = static_cast<int64_t>(b) * H * W * c8;
  const int64_t p_base = static_cast<int64_t>(b) * PH * PW * c8;
  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < per_sample; it += gridDim.x * blockDim.x) {
    const int q = it % c8;
    const int blk = it / c8;
    const int j = blk % BW;
    const int i = blk / BW;
    uint2 am[4];
    uint4 gv[4], pv[4];
    bool ok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int ph = i + (k >> 1), pw = j + (k & 1);
      ok[k] = ph < PH && pw < PW;
      am[k] = make_uint2(0xffffffffu, 0xffffffffu);
      gv[k] = pv[k] = make_uint4(0, 0, 0, 0);
      if (ok[k]) {
        const int64_t pi = p_base + (static_cast<int64_t>(ph) * PW + pw) * c8 + q;
        am[k] = __ldg(reinterpret_cast<const uint2*>(argmax) + pi);
        gv[k] = __ldg(reinterpret_cast<const uint4*>(g) + pi);
        pv[k] = __ldg(reinterpret_cast<const uint4*>(pooled) + pi);
      }
    }
    float acc[4][8];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[o][e] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __half2* g2 = reinterpret_cast<const __half2*>(&gv[k]);
      const __half2* p2 = reinterpret_cast<const __half2*>(&pv[k]);
      const int a = k >> 1, bb = k & 1;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int t = ((e < 4 ? am[k].x : am[k].y) >> (8 * (e & 3))) & 0xff;  // 0xff (never a tap) when !ok
        const float gval = (e & 1) ? __high2float(g2[e >> 1]) : __low2float(g2[e >> 1]);
        const float pval = (e & 1) ? __high2float(p2[e >> 1]) : __low2float(p2[e >> 1]);
        const float gm = pval > 0.f ? gval : 0.f;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const int dh = o >> 1, dw = o & 1;
          // tap of window (a, bb) that lands on output (dh, dw): row 2(i+a)-1+r = 2i+dh  ->  r = dh + 1 - 2a
          const int r = dh + 1 - 2 * a, sx = dw + 1 - 2 * bb;
          if (r >= 0 && sx >= 0 && t == r * 3 + sx) acc[o][e] += gm;  // r, sx fold at compile time
        }
      }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int h = 2 * i + (o >> 1), w = 2 * j + (o & 1);
      if (h < H && w < W) store8h(dy, out_base + (static_cast<int64_t>(h) * W + w) * c8 + q, acc[o]);
    }
  }
}

// GN backward, pass 1: per (sample, channel) sums of dy and dy*xhat, dy = g * [relu_ref > 0]
template <int MINB>
__global__ void __launch_bounds__(256, MINB) gn_bwd_reduce_kernel(const GnBwdArgs a) {
  extern __shared__ float s_mem[];  // mean_g [G], rstd_g [G], acc [2C]
  const int b = blockIdx.y;
  const int C = a.C, G = a.G;
  float* s_mean = s_mem;
  float* s_rstd = s_mem + G;
  float* s_acc = s_mem + 2 * G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) group_mean_rstd(a.stats, b, G, g, a.cnt, a.eps, s_mean[g], s_rstd[g]);
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) s_acc[c] = 0.f;
  __syncthreads();
  const int c8 = C >> 3;
  // thread -> fixed channel chunk, strided over pixels (requires blockDim % c8 == 0 or c8 % blockDim == 0)
  const int64_t per_sample = static_cast<int64_t>(a.HW) * c8;
  const int64_t base = static_cast<int64_t>(b) * per_sample;
  const int64_t start = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  if (stride % c8 == 0) {
    const int cc = static_cast<int>(start % c8) * 8;
    float sd[8], sx[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) sd[e] = sx[e] = 0.f;
#pragma unroll 4
    for (int64_t i = start; i < per_sample; i += stride) {
      float g[8], x[8];
      load8(a.g, base + i, 0, g);
      load8(a.x, base + i, a.x_fp32, x);
      if (a.relu_ref) {
        float y[8];
        load8(a.relu_ref, base + i, 0, y);
#pragma unroll
        for (int e = 0; e < 8; ++e) g[e] = (y[e] > 0.f) ? g[e] * a.g_scale : 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int grp = (cc + e) / a.cpg;
        const float xh = (x[e] - s_mean[grp]) * s_rstd[grp];
        sd[e] += g[e];
        sx[e] = fmaf(g[e], xh, sx[e]);
      }
    }
    // lanes l and l + c8 (+ 2 c8 ...) of a warp own the same 8 channels: combine them with shuffles first
    // (for C = 32 this turns 512 shared-memory atomics per warp on 64 addresses into 64)
    for (int off = 16; off >= c8 && off >= 1; off >>= 1) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        sd[e] += __shfl_xor_sync(0xffffffffu, sd[e], off);
        sx[e] += __shfl_xor_sync(0xffffffffu, sx[e], off);
      }
    }
    if ((threadIdx.x & 31) < c8 || c8 >= 32) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        atomicAdd(&s_acc[2 * (cc + e)], sd[e]);
        atomicAdd(&s_acc[2 * (cc + e) + 1], sx[e]);
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x)
    atomicAdd(a.sums + static_cast<int64_t>(b) * 2 * C + c, s_acc[c]);
}

// GN backward, pass 2: dx = rstd*(gamma*dy - (S1 + xhat*S2)/cnt); optionally also writes dy (masked g)
template <bool CLS, int MINB>
__global__ void __launch_bounds__(256, MINB) gn_bwd_apply_kernel(const GnBwdArgs a) {
  extern __shared__ float s_mem[];  // mean [G], rstd [G], k1 [G], k2 [G], ga [C]
  const int b = blockIdx.y;
  const int C = a.C, G = a.G;
  float* s_mean = s_mem;
  float* s_rstd = s_mem + G;
  float* s_k1 = s_mem + 2 * G;
  float* s_k2 = s_mem + 3 * G;
  float* s_ga = s_mem + 4 * G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float mean, rstd;
    group_mean_rstd(a.stats, b, G, g, a.cnt, a.eps, mean, rstd);
    float S1 = 0.f, S2 = 0.f;
    for (int c = g * a.cpg; c < (g + 1) * a.cpg && c < a.C_real; ++c) {
      const float gm = a.gamma[c];
      S1 = fmaf(gm, a.sums[(static_cast<int64_t>(b) * C + c) * 2], S1);
      S2 = fmaf(gm, a.sums[(static_cast<int64_t>(b) * C + c) * 2 + 1], S2);
    }
    s_mean[g] = mean;
    s_rstd[g] = rstd;
    s_k1[g] = rstd * S1 / a.cnt;
    s_k2[g] = rstd * S2 / a.cnt;
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_ga[c] = (c < a.C_real) ? a.gamma[c] : 0.f;
  __syncthreads();
  const int c8 = C >> 3;
  const int64_t per_sample = static_cast<int64_t>(a.HW) * c8;
  const int64_t base = static_cast<int64_t>(b) * per_sample;
  const int64_t start = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const bool fixed = (stride % c8) == 0;  // then a thread keeps ONE 8-channel chunk: coefficients live in registers
  float kA[8], kB[8], kC[8];             // dx = kA * g + kB + kC * x
  auto coeffs = [&](int cc) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int grp = (cc + e) / a.cpg;
      kA[e] = s_rstd[grp] * s_ga[cc + e];
      kC[e] = -s_rstd[grp] * s_k2[grp];
      kB[e] = -s_k1[grp] - s_mean[grp] * kC[e];
    }
  };
  if (fixed) coeffs(static_cast<int>(start % c8) * 8);
  // border-class sums of dx (exact-input stem): interior pixels (class (2, 2): ~93 % of them) accumulate in registers, the
  // border classes go straight to a shared-memory table
  constexpr bool cls = CLS;                    // launcher: C == 32 and a fixed chunk per thread
  float* s_cls = s_mem + 4 * G + C;            // [25][32]
  float acc_in[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc_in[e] = 0.f;
  if (cls) {
    for (int k = threadIdx.x; k < 25 * 32; k += blockDim.x) s_cls[k] = 0.f;
    __syncthreads();
  }
#pragma unroll 2
  for (int64_t i = start; i < per_sample; i += stride) {
    if (!fixed) coeffs(static_cast<int>(i % c8) * 8);
    float g[8], x[8], dx[8];
    load8(a.g, base + i, 0, g);
    load8(a.x, base + i, a.x_fp32, x);
    if (a.relu_ref) {
      float y[8];
      load8(a.relu_ref, base + i, 0, y);
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] = (y[e] > 0.f) ? g[e] * a.g_scale : 0.f;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) dx[e] = fmaf(kA[e], g[e], fmaf(kC[e], x[e], kB[e]));
    store8h(a.dx, base + i, dx);
    if (a.dy_out) store8h(a.dy_out, base + i, g);
    if (cls) {
      // the sums are those of the STORED (fp16) gradient, as the separate pass over dx computed them
      const int pix = static_cast<int>(i >> 2), chunk = static_cast<int>(i & 3);
      const int oh = pix / a.OW, ow = pix - oh * a.OW;
      const int rc = oh < 2 ? oh : (oh <= a.OH - 3 ? 2 : 3 + oh - (a.OH - 2));
      const int sc = ow < 2 ? ow : (ow <= a.OW - 3 ? 2 : 3 + ow - (a.OW - 2));
      if (rc == 2 && sc == 2) {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc_in[e] += __half2float(__float2half_rn(dx[e]));
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(s_cls + (rc * 5 + sc) * 32 + chunk * 8 + e, __half2float(__float2half_rn(dx[e])));
      }
    }
  }
  if (cls) {
    // a thread keeps one 8-channel chunk (start % 4); lanes 4 apart share it: fold them, then one shared-memory add per value
    const int chunk = static_cast<int>(start & 3);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = acc_in[e];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if ((threadIdx.x & 31) < 4) atomicAdd(s_cls + 12 * 32 + chunk * 8 + e, v);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 25 * 32; k += blockDim.x) {
      const float v = s_cls[k];
      if (v != 0.f) atomicAdd(a.class_sums + k, v);
    }
  }
}

// dgamma[c] = sum_b sums[b][c][1], dbeta[c] = sum_b sums[b][c][0]; one warp per channel, lanes over the batch
__global__ void __launch_bounds__(128) gn_param_grad_kernel(const float* __restrict__ sums, int B, int C, int C_real,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            int accumulate) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C_real) return;
  float dg = 0.f, db = 0.f;
  for (int b = lane; b < B; b += 32) {
    const float2 v = *reinterpret_cast<const float2*>(sums + (static_cast<int64_t>(b) * C + c) * 2);
    db += v.x;
    dg += v.y;
  }
  dg = warp_sum(dg);
  db = warp_sum(db);
  if (lane == 0) {
    if (accumulate) { dgamma[c] += dg; dbeta[c] += db; }
    else { dgamma[c] = dg; dbeta[c] = db; }
  }
}

static int gn_grid_x(int64_t per_sample_items, int B, int c8) {
  // enough CTAs for ~4 waves of 148 SMs x 8 CTAs, multiple of c8-friendly stride handled by the kernel
  int64_t want = std::max<int64_t>(1, (148 * 8 * 2) / std::max(1, B));
  int64_t maxb = std::max<int64_t>(1, ceil_div64(per_sample_items, 256));
  (void)c8;
  return static_cast<int>(std::min(want, maxb));
}

int gn_apply_launch(const GnArgs& a, int B, cudaStream_t st) {
  PNVO_REQUIRE(a.C % 8 == 0 && a.C <= 4096, "gn_apply: C=%d", a.C);
  if (B <= 0 || a.HW <= 0) return 0;
  // every block pays the statistics -> coefficient prologue: at least `min_items` vectors per thread
  // (4 vectors per thread: 579 -> 570 us over the 32 launches of a ResNet-18 step at B = 256, all of it on layer4)
  static const int min_items = getenv("PNVO_GN_MIN_ITEMS") ? atoi(getenv("PNVO_GN_MIN_ITEMS")) : 4;
  const int64_t items = static_cast<int64_t>(a.HW) * (a.C / 8);
  const int gx = static_cast<int>(std::min<int64_t>(gn_grid_x(items, B, a.C / 8),
                                                    std::max<int64_t>(1, ceil_div64(items, 256 * min_items))));
  const size_t sm = 2 * a.C * sizeof(float);
  if (a.y_lo || a.res_lo || a.x_lo) gn_apply_kernel<true><<<dim3(gx, B), 256, sm, st>>>(a);
  else gn_apply_kernel<false><<<dim3(gx, B), 256, sm, st>>>(a);
  count_launch();
  return check_launch("gn_apply");
}
int gn_pool_launch(const GnArgs& a, int B, int H, int W, int PH, int PW, uint8_t* argmax, cudaStream_t st) {
  PNVO_REQUIRE(a.C % 8 == 0 && a.C <= 4096, "gn_pool: C=%d", a.C);
  if (B <= 0) return 0;
  const int gx = gn_grid_x(static_cast<int64_t>(PH) * PW * (a.C / 8), B, a.C / 8);
  gn_pool_kernel<<<dim3(gx, B), 256, 2 * a.C * sizeof(float), st>>>(a, H, W, PH, PW, argmax);
  count_launch();
  return check_launch("gn_pool");
}
int pool_bwd_launch(const __half* g, const __half* pooled, const uint8_t* argmax, __half* dy, int B, int H, int W,
                    int PH, int PW, int C, cudaStream_t st) {
  if (B <= 0) return 0;
  static const int blocked = getenv("PNVO_POOL_BWD_2X2") ? atoi(getenv("PNVO_POOL_BWD_2X2")) : 1;
  if (blocked && PH == (H + 1) / 2 && PW == (W + 1) / 2) {  // 3x3 / stride 2 / pad 1 geometry
    const int gx2 = gn_grid_x(static_cast<int64_t>((H + 1) / 2) * ((W + 1) / 2) * (C / 8), B, C / 8);
    pool_bwd2x2_kernel<<<dim3(gx2, B), 256, 0, st>>>(g, pooled, argmax, dy, H, W, PH, PW, C);
    count_launch();
    return check_launch("pool_bwd2x2");
  }
  const int gx = gn_grid_x(static_cast<int64_t>(H) * W * (C / 8), B, C / 8);
  pool_bwd_kernel<<<dim3(gx, B), 256, 0, st>>>(g, pooled, argmax, dy, H, W, PH, PW, C);
  count_launch();
  return check_launch("pool_bwd");
}
int gn_bwd_reduce_launch(const GnBwdArgs& a, int B, cudaStream_t st) {
  PNVO_REQUIRE(a.C % 8 == 0 && a.C <= 2048, "gn_bwd_reduce: C=%d", a.C);
  if (B <= 0 || a.HW <= 0) return 0;
  const int c8 = a.C / 8;
  PNVO_REQUIRE(256 % c8 == 0 || c8 % 256 == 0, "gn_bwd_reduce: C/8=%d must divide or be a multiple of 256", c8);
  int gx = gn_grid_x(static_cast<int64_t>(a.HW) * c8, B, c8);
  if (c8 > 256) gx = ((gx + c8 / 256 - 1) / (c8 / 256)) * (c8 / 256);  // keep gridDim.x*256 a multiple of c8
  // capped at 80 registers (3 CTAs per SM, ~25 spilled words; uncapped: 128-168 registers): 6.75 -> 6.71 ms per ResNet-18
  // step; PNVO_GN_REDUCE_MINB=1 selects the uncapped kernel
  static const bool cap = !(getenv("PNVO_GN_REDUCE_MINB") && atoi(getenv("PNVO_GN_REDUCE_MINB")) == 1);
  if (cap) gn_bwd_reduce_kernel<3><<<dim3(gx, B), 256, (2 * a.G + 2 * a.C) * sizeof(float), st>>>(a);
  else gn_bwd_reduce_kernel<1><<<dim3(gx, B), 256, (2 * a.G + 2 * a.C) * sizeof(float), st>>>(a);
  count_launch();
  return check_launch("gn_bwd_reduce");
}
int gn_bwd_apply_launch(const GnBwdArgs& a, int B, cudaStream_t st) {
  PNVO_REQUIRE(a.C % 8 == 0 && a.C <= 2048, "gn_bwd_apply: C=%d", a.C);
  if (B <= 0 || a.HW <= 0) return 0;
  const int gx = gn_grid_x(static_cast<int64_t>(a.HW) * (a.C / 8), B, a.C / 8);
  PNVO_REQUIRE(!a.class_sums || (a.C == 32 && a.OH * a.OW == a.HW && a.OH >= 4 && a.OW >= 4 && (static_cast<int64_t>(gx) * 256) % 4 == 0),
               "gn_bwd_apply: border-class sums need C == 32 and the output geometry");
  // capped at 80 registers (3 CTAs per SM, 10-20 spilled words): measured faster than the 95 / 118-register builds -- 6.73 ->
  // 6.70 ms per ResNet-18 step, 24.39 -> 24.00 ms per ResNet-50 step; PNVO_GN_APPLY_MINB=1 selects the uncapped kernels
  static const bool cap = !(getenv("PNVO_GN_APPLY_MINB") && atoi(getenv("PNVO_GN_APPLY_MINB")) == 1);
  const size_t sm = (4 * a.G + a.C + (a.class_sums ? 25 * 32 : 0)) * sizeof(float);
  if (a.class_sums) {
    if (cap) gn_bwd_apply_kernel<true, 3><<<dim3(gx, B), 256, sm, st>>>(a);
    else gn_bwd_apply_kernel<true, 1><<<dim3(gx, B), 256, sm, st>>>(a);
  } else {
    if (cap) gn_bwd_apply_kernel<false, 3><<<dim3(gx, B), 256, sm, st>>>(a);
    else gn_bwd_apply_kernel<false, 1><<<dim3(gx, B), 256, sm, st>>>(a);
  }
  count_launch();
  return check_launch("gn_bwd_apply");
}
int gn_param_grad_launch(const float* sums, int B, int C, int C_real, float* dgamma, float* dbeta, int accumulate,
                         cudaStream_t st) {
  gn_param_grad_kernel<<<ceil_div(C_real * 32, 128), 128, 0, st>>>(sums, B, C, C_real, dgamma, dbeta, accumulate);
  count_launch();
  return check_launch("gn_param_grad");
}

// ------------------------------------------------------------------------------------------------
// weight packing: OIHW fp32 -> [n][ (r*S+s)*Cin_pad + c ] fp16 (fprop / wgrad layout), and the dgrad
// layout [c][ ((R-1-r)*S + (S-1-s))*Cout_pad + n ].  Destination buffers are zero-initialised once; pad
// entries are never written.
// ------------------------------------------------------------------------------------------------
// t_mode & 4: data-gradient weights of a 3x3 / stride 2 / pad 1 convolution split by the PARITY CLASS of the input pixel.
// gx[h, w] = sum over (r, s) with (h + 1 - r), (w + 1 - s) even of dy[(h + 1 - r) / 2, (w + 1 - s) / 2] W[r, s]: for even h
// only r = 1 contributes (dy row h / 2), for odd h the rows r = 2 (dy row (h - 1) / 2) and r = 0 (the next dy row).  Class
// (ph, pw) is therefore a dense stride-1 convolution of dy with (1 + ph) x (1 + pw) taps -- 1 + 2 + 2 + 4 = 9 taps for four
// input pixels instead of 9 taps per pixel over a zero-upsampled dy.  Layout: four matrices [class][ru16(cin_pad)][ld_t],
// element (c, (r' (1 + pw) + s') cout_pad + n) with r' = 0 <-> r = 2 - ph ... (r = 1 when ph = 0; r = 2, 0 when ph = 1).
__device__ __forceinline__ int64_t s2_class_offset(int r, int s, int c, int cin_pad, int ld_t, int cout_pad) {
  const int ph = (r != 1) ? 1 : 0, pw = (s != 1) ? 1 : 0;
  const int rp = (r == 0) ? 1 : 0, sp = (s == 0) ? 1 : 0;
  const int64_t rows = (cin_pad + 15) / 16 * 16;
  return (static_cast<int64_t>(ph * 2 + pw) * rows + c) * ld_t + (rp * (1 + pw) + sp) * cout_pad;
}
__global__ void pack_w_kernel(const float* __restrict__ w, int Cout, int Cin, int R, int S, __half* __restrict__ wp,
                              int cin_pad, int ld_p, __half* __restrict__ wt, int cout_pad, int ld_t, int t_mode,
                              int src_ld) {
  const int64_t total = static_cast<int64_t>(Cout) * Cin * R * S;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int s = static_cast<int>(i % S);
    const int r = static_cast<int>((i / S) % R);
    const int c = static_cast<int>((i / (static_cast<int64_t>(S) * R)) % Cin);
    const int n = static_cast<int>(i / (static_cast<int64_t>(S) * R * Cin));
    // src_ld: elements between source rows (a Linear whose rows carry extra, non-visual columns)
    const float wv = w[static_cast<int64_t>(n) * src_ld + (i - static_cast<int64_t>(n) * Cin * R * S)];
    __half h = __float2half_rn(wv);
    if (t_mode & 2) h = __float2half_rn(wv - __half2float(h));  // residual plane of the split-fp16 representation
    if (wp) wp[static_cast<int64_t>(n) * ld_p + (r * S + s) * cin_pad + c] = h;
    if (wt) {
      if (t_mode & 4) wt[s2_class_offset(r, s, c, cin_pad, ld_t, cout_pad) + n] = h;
      else if ((t_mode & 1) == 0) wt[static_cast<int64_t>(c) * ld_t + ((R - 1 - r) * S + (S - 1 - s)) * cout_pad + n] = h;
      else wt[static_cast<int64_t>((r * S + s) * cin_pad + c) * ld_t + n] = h;
    }
  }
}
__global__ void unpack_dw_kernel(const float* __restrict__ dwp, int Cout, int Cin, int R, int S, int cin_pad, int ld_p,
                                 float* __restrict__ grad, int accumulate, int dst_ld) {
  const int64_t total = static_cast<int64_t>(Cout) * Cin * R * S;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int s = static_cast<int>(i % S);
    const int r = static_cast<int>((i / S) % R);
    const int c = static_cast<int>((i / (static_cast<int64_t>(S) * R)) % Cin);
    const int n = static_cast<int>(i / (static_cast<int64_t>(S) * R * Cin));
    const float v = dwp[static_cast<int64_t>(n) * ld_p + (r * S + s) * cin_pad + c];
    const int64_t o = static_cast<int64_t>(n) * dst_ld + (i - static_cast<int64_t>(n) * Cin * R * S);
    grad[o] = accumulate ? grad[o] + v : v;
  }
}
int pack_w_launch(const float* w, int Cout, int Cin, int R, int S, __half* wp, int cin_pad, int ld_p, __half* wt,
                  int cout_pad, int ld_t, int t_mode, cudaStream_t st, int src_ld) {
  if (src_ld <= 0) src_ld = Cin * R * S;
  const int64_t total = static_cast<int64_t>(Cout) * Cin * R * S;
  if (total <= 0) return 0;
  pack_w_kernel<<<static_cast<int>(std::min<int64_t>(ceil_div64(total, 256), 148 * 4)), 256, 0, st>>>(
      w, Cout, Cin, R, S, wp, cin_pad, ld_p, wt, cout_pad, ld_t, t_mode, src_ld);
  count_launch();
  return check_launch("pack_w");
}
int unpack_dw_launch(const float* dwp, int Cout, int Cin, int R, int S, int cin_pad, int ld_p, float* grad,
                     int accumulate, cudaStream_t st, int dst_ld) {
  if (dst_ld <= 0) dst_ld = Cin * R * S;
  const int64_t total = static_cast<int64_t>(Cout) * Cin * R * S;
  if (total <= 0) return 0;
  unpack_dw_kernel<<<static_cast<int>(std::min<int64_t>(ceil_div64(total, 256), 148 * 4)), 256, 0, st>>>(
      dwp, Cout, Cin, R, S, cin_pad, ld_p, grad, accumulate, dst_ld);
  count_launch();
  return check_launch("unpack_dw");
}

// ------------------------------------------------------------------------------------------------
// Batched variants: one launch walks a device-resident table of descriptors (blockIdx.y = entry) instead of one
// 3-5 us launch per layer -- 22 weight packs, 22 gradient unpacks and 21 GroupNorm parameter-gradient reductions per
// training step become three launches.
// ------------------------------------------------------------------------------------------------
// A block takes kPackRows consecutive output channels (x a chunk of input channels when a row exceeds the staging buffer):
// the OIHW rows are read contiguously, converted once and staged in shared memory; the fprop layout is then written as
// contiguous runs along c and the dgrad / transposed layout as 16-byte vectors along n (the first version walked the
// elements in OIHW order and scattered 2-byte stores into both layouts: 109 us per step for 11 M weights).
static constexpr int kPackRows = 8;
static constexpr int kPackElems = 2304;  // staged elements per row (fp16): a 3x3 x 256-channel filter row
__global__ void __launch_bounds__(256) pack_w_multi_kernel(const PackDesc* __restrict__ tab) {
  __shared__ __align__(16) __half sm[kPackRows][kPackElems];
  const PackDesc d = tab[blockIdx.y];
  const int RS = d.R * d.S;
  const int per = d.Cin * RS;
  const int CC = min(d.Cin, kPackElems / RS);          // channels per chunk (>= 1: RS <= 64)
  const int n_chunks = (d.Cin + CC - 1) / CC;
  const int n_groups = (d.Cout + kPackRows - 1) / kPackRows;
  for (int u = blockIdx.x; u < n_groups * n_chunks; u += gridDim.x) {
    const int n0 = (u / n_chunks) * kPackRows;
    const int cb = (u % n_chunks) * CC;
    const int cc = min(CC, d.Cin - cb);
    const int len = cc * RS;
    __syncthreads();  // the previous unit's readers are done with sm
    for (int e = 0; e < kPackRows; ++e) {
      const bool row = n0 + e < d.Cout;
      const float* src = d.w + static_cast<int64_t>(n0 + e) * d.src_ld + cb * RS;
      for (int j = threadIdx.x; j < len; j += blockDim.x) {
        const float wv = row ? src[j] : 0.f;
        __half h = __float2half_rn(wv);
        if (d.t_mode & 2) h = __float2half_rn(wv - __half2float(h));  // residual plane (split-fp16 mode)
        sm[e][j] = h;
      }
    }
    __syncthreads();
    if (d.wp) {
      for (int e = 0; e < kPackRows && n0 + e < d.Cout; ++e) {
        __half* dst = d.wp + static_cast<int64_t>(n0 + e) * d.ld_p + cb;
        for (int k = threadIdx.x; k < len; k += blockDim.x) {
          const int tap = k / cc, c = k - tap * cc;
          dst[tap * d.cin_pad + c] = sm[e][c * RS + tap];
        }
      }
    }
    if (d.wt) {
      // n0 is a multiple of 8 and cout_pad / ld_t are multiples of 8: the 8 output channels of (c, tap) are one aligned
      // 16-byte vector (rows beyond Cout hold zeros, like the zero-initialised padding they overwrite)
      for (int k = threadIdx.x; k < len; k += blockDim.x) {
        const int c = k / RS, tap = k - c * RS;
        uint4 v;
        __half* hv = reinterpret_cast<__half*>(&v);
#pragma unroll
        for (int e = 0; e < kPackRows; ++e) hv[e] = sm[e][k];
        int64_t o;
        if (d.t_mode & 4) {
          const int r = tap / d.S, sx = tap - r * d.S;
          o = s2_class_offset(r, sx, cb + c, d.cin_pad, d.ld_t, d.cout_pad) + n0;
        } else if ((d.t_mode & 1) == 0) {
          const int r = tap / d.S, sx = tap - r * d.S;
          o = static_cast<int64_t>(cb + c) * d.ld_t + ((d.R - 1 - r) * d.S + (d.S - 1 - sx)) * d.cout_pad + n0;
        } else {
          o = static_cast<int64_t>(tap * d.cin_pad + cb + c) * d.ld_t + n0;
        }
        *reinterpret_cast<uint4*>(d.wt + o) = v;
      }
    }
  }
}
// Gradient unpack, same staging: packed rows are read as contiguous runs along c, OIHW rows written contiguously.
static constexpr int kUnpackRows = 4;
__global__ void __launch_bounds__(256) unpack_dw_multi_kernel(const UnpackDesc* __restrict__ tab) {
  __shared__ float sm[kUnpackRows][kPackElems];
  const UnpackDesc d = tab[blockIdx.y];
  const int RS = d.R * d.S;
  const int CC = min(d.Cin, kPackElems / RS);
  const int n_chunks = (d.Cin + CC - 1) / CC;
  const int n_groups = (d.Cout + kUnpackRows - 1) / kUnpackRows;
  for (int u = blockIdx.x; u < n_groups * n_chunks; u += gridDim.x) {
    const int n0 = (u / n_chunks) * kUnpackRows;
    const int cb = (u % n_chunks) * CC;
    const int cc = min(CC, d.Cin - cb);
    const int len = cc * RS;
    __syncthreads();
    for (int e = 0; e < kUnpackRows && n0 + e < d.Cout; ++e) {
      const float* src = d.dwp + static_cast<int64_t>(n0 + e) * d.ld_p + cb;
      for (int k = threadIdx.x; k < len; k += blockDim.x) {
        const int tap = k / cc, c = k - tap * cc;
        sm[e][c * RS + tap] = src[tap * d.cin_pad + c];
      }
    }
    __syncthreads();
    for (int e = 0; e < kUnpackRows && n0 + e < d.Cout; ++e) {
      float* dst = d.grad + static_cast<int64_t>(n0 + e) * d.dst_ld + cb * RS;
      for (int j = threadIdx.x; j < len; j += blockDim.x) dst[j] = d.accumulate ? dst[j] + sm[e][j] : sm[e][j];
    }
  }
}
__global__ void __launch_bounds__(128) gn_param_grad_multi_kernel(const GnParamDesc* __restrict__ tab, int B) {
  const GnParamDesc d = tab[blockIdx.y];
  const int lane = threadIdx.x & 31;
  for (int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < d.C_real; c += (gridDim.x * blockDim.x) >> 5) {
    float dg = 0.f, db = 0.f;
    for (int b = lane; b < B; b += 32) {
      const float2 v = *reinterpret_cast<const float2*>(d.sums + (static_cast<int64_t>(b) * d.C + c) * 2);
      db += v.x;
      dg += v.y;
    }
    dg = warp_sum(dg);
    db = warp_sum(db);
    if (lane == 0) {
      d.dgamma[c] = dg;
      d.dbeta[c] = db;
    }
  }
}
int multi_launch(int code, const void* table, int n, int B, cudaStream_t st) {
  PNVO_REQUIRE(table && n > 0, "multi op: empty table");
  if (code == PNVO_OP_PACK_W_MULTI) pack_w_multi_kernel<<<dim3(64, n), 256, 0, st>>>(static_cast<const PackDesc*>(table));
  else if (code == PNVO_OP_UNPACK_DW_MULTI) unpack_dw_multi_kernel<<<dim3(64, n), 256, 0, st>>>(static_cast<const UnpackDesc*>(table));
  else gn_param_grad_multi_kernel<<<dim3(8, n), 128, 0, st>>>(static_cast<const GnParamDesc*>(table), B);
  count_launch();
  return check_launch("multi op");
}

// ------------------------------------------------------------------------------------------------
// regression head (vo_cnn.py:216-227).  fc1 runs on the tensor cores as a 6x11 "convolution"; here:
//   bias_relu   : h = relu(z + b1) (fp32 z from the GEMM) -> fp32 h and fp16 h
//   head_fwd    : out[b][o] = <h[b], W2[o]> + b2[o], one warp per (b, o), shuffle reduction
//   head_bwd    : dW2, db2 (reduction over the batch), dz = (W2^T dout) * [h > 0] -> fp16, db1
// ------------------------------------------------------------------------------------------------
__global__ void bias_relu_kernel(const float* __restrict__ z, const float* __restrict__ bias, int B, int N, int relu,
                                 float* __restrict__ h32, __half* __restrict__ h16) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(B) * N) return;
  float v = z[i] + bias[i % N];
  if (relu) v = fmaxf(v, 0.f);
  if (h32) h32[i] = v;
  if (h16) h16[i] = __float2half_rn(v);
}
__global__ void head_fwd_kernel(const float* __restrict__ h, const float* __restrict__ W, const float* __restrict__ bias,
                                int B, int K, int O, float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * O) return;
  const int b = warp / O, o = warp % O;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc = fmaf(h[static_cast<int64_t>(b) * K + k], W[static_cast<int64_t>(o) * K + k], acc);
  acc = warp_sum(acc);
  if (lane == 0) out[b * O + o] = acc + bias[o];
}
// one block per hidden unit k: dW2[o][k], db1[k]; dz[b][k]
__global__ void __launch_bounds__(128) head_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ h,
                                                       const float* __restrict__ W, int B, int K, int O,
                                                       float* __restrict__ dW, float* __restrict__ db2,
                                                       __half* __restrict__ dz16, float* __restrict__ db1,
                                                       int accumulate, float dh_scale) {
  const int k = blockIdx.x;
  __shared__ float s_red[4][9];
  float accW[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) accW[o] = 0.f;
  float accb1 = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float hv = h[static_cast<int64_t>(b) * K + k];
    float dh = 0.f;
    for (int o = 0; o < O; ++o) {
      const float d = dout[b * O + o];
      accW[o] = fmaf(d, hv, accW[o]);
      dh = fmaf(d, W[static_cast<int64_t>(o) * K + k], dh);
    }
    const float dzv = hv > 0.f ? dh * dh_scale : 0.f;  // h is post-dropout: zeros carry the mask
    dz16[static_cast<int64_t>(b) * K + k] = __float2half_rn(dzv);
    accb1 += dzv;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = 0; o < O; ++o) accW[o] = warp_sum(accW[o]);
  accb1 = warp_sum(accb1);
  if (lane == 0) {
    for (int o = 0; o < O; ++o) s_red[warp][o] = accW[o];
    s_red[warp][8] = accb1;
  }
  __syncthreads();
  if (threadIdx.x <= O || threadIdx.x == 8) {
    const int o = threadIdx.x;
    if (o < O) {
      const float v = s_red[0][o] + s_red[1][o] + s_red[2][o] + s_red[3][o];
      float* dst = dW + static_cast<int64_t>(o) * K + k;
      *dst = accumulate ? *dst + v : v;
    } else if (o == 8) {
      const float v = s_red[0][8] + s_red[1][8] + s_red[2][8] + s_red[3][8];
      db1[k] = accumulate ? db1[k] + v : v;
    }
  }
  if (k == 0 && threadIdx.x < O) {
    float v = 0.f;
    for (int b = 0; b < B; ++b) v += dout[b * O + threadIdx.x];
    db2[threadIdx.x] = accumulate ? db2[threadIdx.x] + v : v;
  }
}
// backward of h = relu(z + b): dz = dh * [h > 0] (fp16 for the tensor-core wgrad/dgrad), db = sum_b dz
__global__ void __launch_bounds__(128) bias_relu_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ h,
                                                            int B, int N, __half* __restrict__ dz16,
                                                            float* __restrict__ db, int accumulate) {
  const int k = blockIdx.x;
  __shared__ float s_red[4];
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float v = h[static_cast<int64_t>(b) * N + k] > 0.f ? dh[static_cast<int64_t>(b) * N + k] : 0.f;
    dz16[static_cast<int64_t>(b) * N + k] = __float2half_rn(v);
    acc += v;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float v = s_red[0] + s_red[1] + s_red[2] + s_red[3];
    db[k] = accumulate ? db[k] + v : v;
  }
}
int bias_relu_bwd_launch(const float* dh, const float* h, int B, int N, __half* dz16, float* db, int accumulate,
                         cudaStream_t st) {
  if (B <= 0 || N <= 0) return 0;
  bias_relu_bwd_kernel<<<N, 128, 0, st>>>(dh, h, B, N, dz16, db, accumulate);
  count_launch();
  return check_launch("bias_relu_bwd");
}
int bias_relu_launch(const float* z, const float* bias, int B, int N, int relu, float* h32, __half* h16,
                     cudaStream_t st) {
  if (B * N <= 0) return 0;
  bias_relu_kernel<<<ceil_div(B * N, 256), 256, 0, st>>>(z, bias, B, N, relu, h32, h16);
  count_launch();
  return check_launch("bias_relu");
}
int head_fwd_launch(const float* h, const float* W, const float* bias, int B, int K, int O, float* out,
                    cudaStream_t st) {
  if (B * O <= 0) return 0;
  head_fwd_kernel<<<ceil_div(B * O * 32, 128), 128, 0, st>>>(h, W, bias, B, K, O, out);
  count_launch();
  return check_launch("head_fwd");
}
int head_bwd_launch(const float* dout, const float* h, const float* W, int B, int K, int O, float* dW, float* db2,
                    __half* dz16, float* db1, int accumulate, float dh_scale, cudaStream_t st) {
  PNVO_REQUIRE(O <= 8, "head_bwd: output_dim %d > 8", O);
  if (B <= 0) return 0;
  head_bwd_kernel<<<K, 128, 0, st>>>(dout, h, W, B, K, O, dW, db2, dz16, db1, accumulate, dh_scale);
  count_launch();
  return check_launch("head_bwd");
}

// ------------------------------------------------------------------------------------------------
// Inverted dropout, in place (vo_cnn.py:218,224: nn.Dropout(p) before both Linear layers).  The keep mask
// is a counter-based hash of (seed, site, element index); the seed lives in device memory and is advanced
// by a one-thread kernel after the last dropout site of a forward pass, so replayed op programs draw fresh
// masks.  The mask is not stored: the backward pass recovers it from the zeros of the post-ReLU tensor.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash_u32(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return static_cast<uint32_t>(x >> 16);
}
__global__ void dropout_kernel(void* __restrict__ buf, int64_t n, int is_fp16, const uint64_t* __restrict__ seed,
                               uint64_t site, uint32_t thresh, float scale) {
  const uint64_t s0 = *seed + site * 0x9E3779B97F4A7C15ULL;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const bool keep = hash_u32(s0 + static_cast<uint64_t>(i) * 0xD6E8FEB86659FD93ULL) >= thresh;
    if (is_fp16) {
      __half* p = static_cast<__half*>(buf);
      p[i] = keep ? __float2half_rn(__half2float(p[i]) * scale) : __float2half_rn(0.f);
    } else {
      float* p = static_cast<float*>(buf);
      p[i] = keep ? p[i] * scale : 0.f;
    }
  }
}
__global__ void seed_advance_kernel(uint64_t* seed) { *seed = *seed * 6364136223846793005ULL + 1442695040888963407ULL; }
int dropout_launch(void* buf, int64_t n, int is_fp16, uint64_t* seed, int site, float p, int advance, cudaStream_t st) {
  PNVO_REQUIRE(buf && seed && p >= 0.f && p < 1.f, "dropout: bad arguments");
  if (n > 0 && p > 0.f) {
    const uint32_t thresh = static_cast<uint32_t>(static_cast<double>(p) * 4294967296.0);
    dropout_kernel<<<static_cast<int>(std::min<int64_t>(ceil_div64(n, 256), 148 * 8)), 256, 0, st>>>(
        buf, n, is_fp16, seed, static_cast<uint64_t>(site), thresh, 1.f / (1.f - p));
    count_launch();
  }
  if (advance) {
    seed_advance_kernel<<<1, 1, 0, st>>>(seed);
    count_launch();
  }
  return check_launch("dropout");
}

// ------------------------------------------------------------------------------------------------
// VO regression loss (vo_cnn_engine.py:135-198): loss = sum_i w_i * mean_b (t_bi - p_bi)^2 (dz optionally
// masked), and its gradient dout_bi = 2 w_i mask_bi (p_bi - t_bi) / B (scaled by grad_scale, e.g. 1/world).
// With per-row data types (geometric-invariance batches, vo_cnn_regression_geo_invariance_engine.py:676-750) the
// reference sums a SEPARATE mean per data type (cur-rel-to-prev rows, prev-rel-to-cur rows): every row is then
// normalised by the number of rows of its own type.  One block; B is a few hundred.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mse_loss_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                                       const float* __restrict__ dz_mask,
                                                       const int64_t* __restrict__ data_types, int B, int O, float w0,
                                                       float w1, float w2, float grad_scale, float* __restrict__ dout,
                                                       float* __restrict__ loss) {
  __shared__ float s_part[8];
  __shared__ int s_cnt[2];
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  if (data_types) {
    int c1 = 0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) c1 += (data_types[b] != 0) ? 1 : 0;
    c1 = __reduce_add_sync(0xffffffffu, c1);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_cnt[1], c1);
    __syncthreads();
    if (threadIdx.x == 0) s_cnt[0] = B - s_cnt[1];
  } else if (threadIdx.x == 0) {
    s_cnt[0] = B;
  }
  __syncthreads();
  const float inv0 = s_cnt[0] > 0 ? 1.f / static_cast<float>(s_cnt[0]) : 0.f;
  const float inv1 = s_cnt[1] > 0 ? 1.f / static_cast<float>(s_cnt[1]) : 0.f;
  float acc = 0.f;
  for (int i = threadIdx.x; i < B * O; i += blockDim.x) {
    const int b = i / O, o = i - b * O;
    float w = (o == 0) ? w0 : ((o == 1) ? w1 : w2);
    if (o == 1 && dz_mask) w *= dz_mask[b];
    w *= (data_types && data_types[b] != 0) ? inv1 : inv0;
    const float d = pred[i] - tgt[i];
    acc = fmaf(w * d, d, acc);
    if (dout) dout[i] = 2.f * w * d * grad_scale;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += s_part[k];
    *loss = t;
  }
}
int mse_loss_launch(const float* pred, const float* tgt, const float* dz_mask, const int64_t* data_types, int B, int O,
                    float w0, float w1, float w2, float grad_scale, float* dout, float* loss, cudaStream_t st) {
  PNVO_REQUIRE(pred && tgt && loss, "mse_loss: null argument");
  mse_loss_kernel<<<1, 256, 0, st>>>(pred, tgt, dz_mask, data_types, B, O, w0, w1, w2, grad_scale, dout, loss);
  count_launch();
  return check_launch("mse_loss");
}

// ------------------------------------------------------------------------------------------------
// Geometric-inversion loss (vo_cnn_regression_geo_invariance_engine.py:367-449,781-792) over n pairs (a = cur relative
// to prev, b = prev relative to cur):
//   rot = mean_i (a.yaw + b.yaw)^2
//   pos = mean_{i,k} m_ik (b.xz + R(b.yaw) a.xz)_k^2,  R = [[c, s], [-s, c]] (left-handed), m_i1 = 0 for MOVE_FORWARD
// loss[0] += weight * (rot + pos) (added to the regression loss already there), loss[1] = rot, loss[2] = pos;
// dout += weight * grad_scale * d(rot + pos)/d(pred).
// Which rows pair up: with data_types == null the batch is taken as interleaved [a0, b0, a1, b1, ...] (every row);
// with data_types the reference's selection is reproduced (:781-792): only rows whose action is TURN_LEFT / TURN_RIGHT
// take part, in batch order, and that sub-sequence must alternate [cur-rel-to-prev, prev-rel-to-cur, ...] (:373-374
// asserts it) -- a violation (or an odd count) sets *err and contributes nothing.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) geo_inv_loss_kernel(const float* __restrict__ pred,
                                                           const int64_t* __restrict__ actions,
                                                           const int64_t* __restrict__ data_types, int B, int O,
                                                           int move_forward, int turn_left, int turn_right, float weight,
                                                           float grad_scale, float* __restrict__ dout,
                                                           float* __restrict__ loss, int* __restrict__ err) {
  extern __shared__ int s_rows[];  // valid rows in batch order
  __shared__ float s_rot[8], s_pos[8];
  __shared__ int s_nvalid, s_bad;
  if (threadIdx.x == 0) {
    int nv = 0, bad = 0;
    if (data_types) {
      for (int b = 0; b < B; ++b) {
        const int64_t a = actions[b];
        if (a == turn_left || a == turn_right) {
          if (data_types[b] != (nv & 1)) bad = 1;
          s_rows[nv++] = b;
        }
      }
      if (nv & 1) bad = 1;
    } else {
      for (int b = 0; b < B; ++b) s_rows[b] = b;
      nv = B;
    }
    s_nvalid = bad ? 0 : nv;
    s_bad = bad;
    if (bad && err) *err = 1;
  }
  __syncthreads();
  const int n = s_nvalid / 2;
  float acc_rot = 0.f, acc_pos = 0.f;
  const float inv_n = n > 0 ? 1.f / static_cast<float>(n) : 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int ra = s_rows[2 * i], rb = s_rows[2 * i + 1];
    const float* a = pred + static_cast<int64_t>(ra) * O;
    const float* b = pred + static_cast<int64_t>(rb) * O;
    const float m = (actions[ra] == move_forward) ? 0.f : 1.f;
    const float yaw = a[2] + b[2];
    const float c = cosf(b[2]), s = sinf(b[2]);
    const float p0 = c * a[0] + s * a[1], p1 = -s * a[0] + c * a[1];
    const float e0 = b[0] + p0, e1 = b[1] + p1;
    acc_rot += yaw * yaw;
    acc_pos += e0 * e0 + m * e1 * e1;
    if (dout) {
      const float k = weight * grad_scale * inv_n;
      float* da = dout + static_cast<int64_t>(ra) * O;
      float* db = dout + static_cast<int64_t>(rb) * O;
      da[0] += k * (e0 * c - m * e1 * s);
      da[1] += k * (e0 * s + m * e1 * c);
      da[2] += k * 2.f * yaw;
      db[0] += k * e0;
      db[1] += k * m * e1;
      db[2] += k * (2.f * yaw + e0 * p1 - m * e1 * p0);
    }
  }
  acc_rot = warp_sum(acc_rot);
  acc_pos = warp_sum(acc_pos);
  if ((threadIdx.x & 31) == 0) {
    s_rot[threadIdx.x >> 5] = acc_rot;
    s_pos[threadIdx.x >> 5] = acc_pos;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = 0.f, q = 0.f;
    for (int k = 0; k < 8; ++k) { r += s_rot[k]; q += s_pos[k]; }
    r *= inv_n;
    q *= inv_n * 0.5f;
    loss[0] += weight * (r + q);
    loss[1] = r;
    loss[2] = q;
  }
}
int geo_inv_loss_launch(const float* pred, const int64_t* actions, const int64_t* data_types, int B, int O,
                        int move_forward, int turn_left, int turn_right, float weight, float grad_scale, float* dout,
                        float* loss, int* err, cudaStream_t st) {
  PNVO_REQUIRE(pred && actions && loss && O >= 3, "geo_inv_loss: bad arguments");
  PNVO_REQUIRE(data_types || B % 2 == 0, "geo_inv_loss: B must be even without data types (interleaved pairs)");
  PNVO_REQUIRE(B <= 8192, "geo_inv_loss: at most 8192 rows");
  if (B == 0) return 0;
  geo_inv_loss_kernel<<<1, 256, B * sizeof(int), st>>>(pred, actions, data_types, B, O, move_forward, turn_left,
                                                       turn_right, weight, grad_scale, dout, loss, err);
  count_launch();
  return check_launch("geo_inv_loss");
}

// ------------------------------------------------------------------------------------------------
// Adam over a flat fp32 bucket (torch.optim.Adam semantics, weight_decay = 0, amsgrad = False)
// ------------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float bc1,
                            float bc2_sqrt, float grad_scale) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * grad_scale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}
int adam_launch(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps,
                int step, float grad_scale, cudaStream_t st) {
  if (n <= 0) return 0;
  const float bc1 = 1.f - powf(b1, static_cast<float>(step));
  const float bc2 = 1.f - powf(b2, static_cast<float>(step));
  adam_kernel<<<static_cast<int>(std::min<int64_t>(ceil_div64(n, 256), 148 * 8)), 256, 0, st>>>(
      p, g, m, v, n, lr, b1, b2, eps, bc1, sqrtf(bc2), grad_scale);
  count_launch();
  return check_launch("adam");
}

}  // namespace pnvo
