// VO input pipeline on the device: the reference derives the discretised-depth and
// top-down channels on the host (vo/dataset/regression_geo_invariance_iter_dataset.py:237-267), ships 30 fp32
// channels per pixel (120 B) and re-reads them for the running statistics and the concat / normalise
// (vo_cnn.py:110-176, running_mean_and_var.py:22-63).  Here the step's inputs stay in their raw form -- uint8 rgb
// pairs [B,H,W,6], fp32 depth pairs [B,H,W,2] and the top-down maps [B,H,W,2] produced by pnvo_topdown_project --
// (22 B per pixel) and two streaming kernels read exactly those bytes:
//   raw_stats    per-channel sum / sum of squares of the assembled (un-normalised) input: rgb exactly in integers,
//                the one-hot depth bins as counts, depth / top-down in fp32 -> fp64 atomics per block
//   raw_assemble [prev: rgb/255, depth, one-hot bins, top-down | cur: ...] -> (v*scale + shift) -> NHWC fp16,
//                channels padded to Cpad, optional W-padded rows for the stem kernel
// The one-hot bin follows base_trainer_with_vo.py:135-167 exactly (fp32 edges, >= / <, closed last bin).
#include "common.cuh"
#include "elem.cuh"

namespace pnvo {

static constexpr int kRawMaxBins = 16;

struct RawArgs {
  const uint8_t* rgb;   // [n_pix][6] (prev rgb, cur rgb) or null
  const void* depth;    // [n_pix][2] fp32 (or fp16 when depth_fp16: the dataset's storage type, widened exactly) or null
  int depth_fp16;
  const float* td;      // [n_pix][2] or null
  const float* edges;   // n_dd + 1 fp32 bin edges (device)
  int use_rgb, use_depth, n_dd, use_td;
  int C, Cpad;          // 2 * per-frame channels, padded channel count (multiple of 8, <= 32)
  int64_t n_pix;
  const float* scale;
  const float* shift;
  __half* out;
  __half* out_lo;       // split-fp16 mode (nullable): value - fp16(value); handled by the generic kernel variant
  int row_w, out_pitch;
  int n_lo;             // exact-input stem: 2 extra channels C, C+1 = residuals t - fp16(t) of the two top-down values
  double* stats;        // [2C] (sum, sumsq) interleaved, accumulated
  int exact;            // exact-input stem: the stored values are CONSTANT affine maps of the raw ones ((byte - 128) / 256
                        // for rgb, the raw value otherwise) -- scale / shift are ignored, so the assembled tensor does not
                        // depend on the batch statistics and the two passes can be one (raw_assemble with stats != null)
  // Geometric-invariance augmentation on the device (regression_geo_invariance_iter_dataset.py:342-420): output sample b
  // is source pair pair_map[b] >> 1, with prev / cur swapped when pair_map[b] & 1.  n_pix then counts OUTPUT pixels and
  // hw = pixels per sample; the raw tensors hold only the source pairs (no second copy over PCIe, no second top-down).
  const int32_t* pair_map;
  int hw;
};

// source pixel + swap flag of output pixel p under the pair map
__device__ __forceinline__ float2 ld_depth_pair(const RawArgs& a, int64_t p) {
  if (a.depth_fp16) return __half22float2(__ldg(reinterpret_cast<const __half2*>(a.depth) + p));
  return __ldg(reinterpret_cast<const float2*>(a.depth) + p);
}

__device__ __forceinline__ int64_t mapped_pixel(const RawArgs& a, int64_t p, bool& flip) {
  const uint32_t p32 = static_cast<uint32_t>(p), hw = static_cast<uint32_t>(a.hw);  // launchers require n_pix < 2^31
  const uint32_t b = p32 / hw;
  const int32_t m = __ldg(a.pair_map + b);
  flip = (m & 1) != 0;
  return static_cast<int64_t>(m >> 1) * a.hw + (p32 - b * hw);
}

// bin i <=> e_i <= d < e_{i+1}, last bin closed at e_n (base_trainer_with_vo.py:143-154).  floor(d * n) is at most
// one bin off (edges are fl(i/n)); the comparisons against the actual fp32 edges make the result exact.
__device__ __forceinline__ int depth_bin(float d, const float* s_edges, int n) {
  if (!(d >= s_edges[0] && d <= s_edges[n])) return -1;  // the reference asserts against this
  int k = min(n - 1, max(0, static_cast<int>(d * static_cast<float>(n))));
  if (d < s_edges[k]) --k;
  else if (k + 1 < n && d >= s_edges[k + 1]) ++k;
  return k;
}

// values per block: rgb 6 x (sum, sumsq) as uint32, depth / td 2 x (sum, sumsq) fp32, bins 2 x kRawMaxBins counts
static constexpr int kRawVals = 12 + 4 + 4 + 2 * kRawMaxBins;

__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
__device__ __forceinline__ double warp_sum_f64(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// Per-thread accumulators of the batch statistics of the assembled (un-normalised, reference-unit) input.
// A thread sees <= 255 pixels (launchers), so its per-bin counts fit in 8 bits: bins 0-7 / 8-15 of a frame are
// packed into two 64-bit words and bumped with one shift-add instead of a compare chain per bin.
struct RawAcc {
  uint32_t ri[12];
  float fs[8];
  unsigned long long cb[4];  // [frame][bins 0-7 | bins 8-15]
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < 12; ++i) ri[i] = 0u;
#pragma unroll
    for (int i = 0; i < 8; ++i) fs[i] = 0.f;
    cb[0] = cb[1] = cb[2] = cb[3] = 0ull;
  }
  // w: the three 16-bit words of the (possibly swapped) rgb pair; d / t: depth and top-down pairs; b0 / b1: depth bins
  __device__ __forceinline__ void add(const uint32_t* w, float2 d, float2 t, int b0, int b1) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const uint32_t lo = w[k] & 0xff, hi = w[k] >> 8;
      ri[4 * k] += lo; ri[4 * k + 1] += lo * lo;
      ri[4 * k + 2] += hi; ri[4 * k + 3] += hi * hi;
    }
    fs[0] += d.x; fs[1] = fmaf(d.x, d.x, fs[1]);
    fs[2] += d.y; fs[3] = fmaf(d.y, d.y, fs[3]);
    fs[4] += t.x; fs[5] = fmaf(t.x, t.x, fs[5]);
    fs[6] += t.y; fs[7] = fmaf(t.y, t.y, fs[7]);
    const unsigned long long i0 = (b0 >= 0) ? (1ull << ((b0 & 7) * 8)) : 0ull;
    const unsigned long long i1 = (b1 >= 0) ? (1ull << ((b1 & 7) * 8)) : 0ull;
    cb[0] += (b0 < 8) ? i0 : 0ull;
    cb[1] += (b0 < 8) ? 0ull : i0;
    cb[2] += (b1 < 8) ? i1 : 0ull;
    cb[3] += (b1 < 8) ? 0ull : i1;
  }
  // warp reduction (integers exactly, floats in fp64), across the 8 warps through shared memory, then thread c < C gathers
  // the (sum, sumsq) of output channel c and adds it to the global accumulator.  All 256 threads of the block must call.
  __device__ __forceinline__ void flush(const RawArgs& a, double (*s_red)[kRawVals]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      const uint32_t x = warp_sum_u32(ri[i]);
      if (lane == 0) s_red[warp][i] = static_cast<double>(x);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double x = warp_sum_f64(static_cast<double>(fs[i]));
      if (lane == 0) s_red[warp][12 + i] = x;
    }
#pragma unroll
    for (int i = 0; i < 2 * kRawMaxBins; ++i) {
      const int f = i / kRawMaxBins, bin = i % kRawMaxBins;
      const uint32_t x = warp_sum_u32(static_cast<uint32_t>((cb[2 * f + (bin >> 3)] >> ((bin & 7) * 8)) & 0xffull));
      if (lane == 0) s_red[warp][20 + i] = static_cast<double>(x);
    }
    __syncthreads();
    const int c = threadIdx.x;
    if (c < a.C) {
      const int cf = a.C >> 1;
      const int f = c / cf;
      int k = c - f * cf;
      auto tot = [&](int i) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_red[w][i];
        return t;
      };
      double s = 0.0, q = 0.0;
      bool done = false;
      if (a.use_rgb) {
        if (k < 3) {
          // rgb channel j = 3f + k lives in word j / 2, half j & 1: ri index 4*(j/2) + 2*(j&1)
          const int j = 3 * f + k;
          const int base = 4 * (j >> 1) + 2 * (j & 1);
          s = tot(base) / 255.0;
          q = tot(base + 1) / (255.0 * 255.0);
          done = true;
        }
        k -= 3;
      }
      if (!done && a.use_depth) {
        if (k == 0) { s = tot(12 + 2 * f); q = tot(12 + 2 * f + 1); done = true; }
        k -= 1;
      }
      if (!done && a.n_dd > 0) {
        if (k < a.n_dd) { s = q = tot(20 + f * kRawMaxBins + k); done = true; }
        k -= a.n_dd;
      }
      if (!done && a.use_td && k == 0) { s = tot(16 + 2 * f); q = tot(16 + 2 * f + 1); }
      atomicAdd(a.stats + 2 * c, s);
      atomicAdd(a.stats + 2 * c + 1, q);
    }
  }
};

// RGB / DEP / TD: 0 or 1; NDD: number of one-hot bins (compile-time layout), or -1 = every flag read at run time
// LO: also write the residual plane value - fp16(value) (split-fp16 forward); the generic variant tests a.out_lo at run time
// STATS: also accumulate the batch statistics of the reference-unit input (exact-input stem only: the stored values
// do not depend on them), saving raw_stats' second pass over the raw tensors.
template <int RGB, int DEP, int NDD, int TD, bool MAP = false, bool LO = false, bool STATS = false>
__global__ void __launch_bounds__(256, STATS ? 3 : 1) raw_assemble_kernel(const RawArgs a) {
  __shared__ __align__(16) float s_scale[kMaxInC];
  __shared__ __align__(16) float s_shift[kMaxInC];
  __shared__ float s_edges[kRawMaxBins + 1];
  __shared__ double s_red[STATS ? 8 : 1][kRawVals];
  RawAcc acc;
  if (STATS) acc.clear();
  if (threadIdx.x < kMaxInC) {
    const int c = threadIdx.x;
    if (a.exact) {
      // exact-input stem (stem_exact.cu): rgb is stored as (byte - 128) / 256 = (byte / 255) * (255 / 256) - 1 / 2, every
      // other channel as its raw value; the fp16 rounding of the result is exact
      const int cf = a.C >> 1;
      const bool rgb = a.use_rgb && c < a.C && (c % cf) < 3;
      s_scale[c] = (c < a.C + a.n_lo) ? (rgb ? 255.f / 256.f : 1.f) : 0.f;
      s_shift[c] = rgb ? -0.5f : 0.f;
    } else {
      s_scale[c] = (c < a.C + a.n_lo) ? (a.scale ? a.scale[c] : 1.f) : 0.f;
      s_shift[c] = (c < a.C + a.n_lo && a.shift) ? a.shift[c] : 0.f;
    }
  }
  if (threadIdx.x <= a.n_dd && a.n_dd > 0) s_edges[threadIdx.x] = a.edges[threadIdx.x];
  __syncthreads();
  constexpr bool kGeneric = NDD < 0;
  const int use_rgb = kGeneric ? a.use_rgb : RGB, use_depth = kGeneric ? a.use_depth : DEP;
  const int n_dd = kGeneric ? a.n_dd : NDD, use_td = kGeneric ? a.use_td : TD;
  const int cf = 3 * use_rgb + use_depth + n_dd + use_td;  // channels per frame
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < a.n_pix;
       p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float v[kMaxInC];
#pragma unroll
    for (int c = 0; c < kMaxInC; ++c) v[c] = 0.f;
    float rgbv[6] = {0, 0, 0, 0, 0, 0};
    float2 d = make_float2(0.f, 0.f), t = make_float2(0.f, 0.f);
    const int64_t p_out = p;
    bool flip = false;
    if (MAP) p = mapped_pixel(a, p_out, flip);
    uint32_t sw[3] = {0u, 0u, 0u};
    int sbin[2] = {-1, -1};
    if (use_rgb) {
      const ushort* r16 = reinterpret_cast<const ushort*>(a.rgb + p * 6);
      const uint32_t w0 = r16[0], w1 = r16[1], w2 = r16[2];
      if (STATS) { sw[0] = w0; sw[1] = w1; sw[2] = w2; }
      // rgb / 255 as the reference divides (vo_cnn.py:117-118): fp32 division
      rgbv[0] = __fdiv_rn(static_cast<float>(w0 & 0xff), 255.f);
      rgbv[1] = __fdiv_rn(static_cast<float>(w0 >> 8), 255.f);
      rgbv[2] = __fdiv_rn(static_cast<float>(w1 & 0xff), 255.f);
      rgbv[3] = __fdiv_rn(static_cast<float>(w1 >> 8), 255.f);
      rgbv[4] = __fdiv_rn(static_cast<float>(w2 & 0xff), 255.f);
      rgbv[5] = __fdiv_rn(static_cast<float>(w2 >> 8), 255.f);
    }
    if (use_depth || n_dd > 0) d = ld_depth_pair(a, p);
    if (use_td) t = __ldg(reinterpret_cast<const float2*>(a.td) + p);
    if (MAP) {
      if (flip) {  // prev <-> cur
#pragma unroll
        for (int c = 0; c < 3; ++c) { const float x = rgbv[c]; rgbv[c] = rgbv[c + 3]; rgbv[c + 3] = x; }
        d = make_float2(d.y, d.x);
        t = make_float2(t.y, t.x);
      }
      p = p_out;
    }
    // channel placement, fully unrolled so that v[] stays in registers; with a compile-time layout every
    // comparison below folds away
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      int cur = f * cf;
      const float dd = f ? d.y : d.x;
      if (use_rgb) {
#pragma unroll
        for (int c = 0; c < kMaxInC; ++c) {
          if (c == cur) v[c] = rgbv[3 * f];
          if (c == cur + 1) v[c] = rgbv[3 * f + 1];
          if (c == cur + 2) v[c] = rgbv[3 * f + 2];
        }
        cur += 3;
      }
      if (use_depth) {
#pragma unroll
        for (int c = 0; c < kMaxInC; ++c)
          if (c == cur) v[c] = dd;
        cur += 1;
      }
      if (n_dd > 0) {
        const int bin = depth_bin(dd, s_edges, n_dd);
        if (STATS) sbin[f] = bin;
#pragma unroll
        for (int c = 0; c < kMaxInC; ++c)
          if (c >= cur && c < cur + n_dd && c - cur == bin) v[c] = 1.f;
        cur += n_dd;
      }
      if (use_td) {
        const float tv = f ? t.y : t.x;
#pragma unroll
        for (int c = 0; c < kMaxInC; ++c)
          if (c == cur) v[c] = tv;
        if (a.n_lo) {  // exact-input stem: the fp16 rounding residual of the top-down value rides in channel C + f
          const float lo = tv - __half2float(__float2half_rn(tv));
#pragma unroll
          for (int c = 0; c < kMaxInC; ++c)
            if (c == 2 * cf + f) v[c] = lo;
        }
      }
    }
    if (STATS) acc.add(sw, d, t, sbin[0], sbin[1]);  // (the STATS variant is never MAP: no swap to undo)
    int64_t opix = p;
    if (a.out_pitch > 0) opix = (p / a.row_w) * a.out_pitch + (p % a.row_w) + 3;  // zero halo left of the image
    __half* __restrict__ out = a.out + opix * a.Cpad;
#pragma unroll
    for (int q = 0; q < kMaxInC / 8; ++q) {
      if (q * 8 < a.Cpad) {
        uint4 u;
        __half2* h2 = reinterpret_cast<__half2*>(&u);
        // 128-bit shared-memory reads of the per-channel constants (ncu: 64 scalar LDS per pixel throttled the MIO)
        const float4 sa = reinterpret_cast<const float4*>(s_scale)[2 * q], sb = reinterpret_cast<const float4*>(s_scale)[2 * q + 1];
        const float4 ha = reinterpret_cast<const float4*>(s_shift)[2 * q], hb = reinterpret_cast<const float4*>(s_shift)[2 * q + 1];
        const int c = q * 8;
        h2[0] = __floats2half2_rn(fmaf(v[c], sa.x, ha.x), fmaf(v[c + 1], sa.y, ha.y));
        h2[1] = __floats2half2_rn(fmaf(v[c + 2], sa.z, ha.z), fmaf(v[c + 3], sa.w, ha.w));
        h2[2] = __floats2half2_rn(fmaf(v[c + 4], sb.x, hb.x), fmaf(v[c + 5], sb.y, hb.y));
        h2[3] = __floats2half2_rn(fmaf(v[c + 6], sb.z, hb.z), fmaf(v[c + 7], sb.w, hb.w));
        *reinterpret_cast<uint4*>(out + q * 8) = u;
        if ((kGeneric || LO) && a.out_lo) {
          const float x[8] = {fmaf(v[c], sa.x, ha.x),     fmaf(v[c + 1], sa.y, ha.y), fmaf(v[c + 2], sa.z, ha.z),
                              fmaf(v[c + 3], sa.w, ha.w), fmaf(v[c + 4], sb.x, hb.x), fmaf(v[c + 5], sb.y, hb.y),
                              fmaf(v[c + 6], sb.z, hb.z), fmaf(v[c + 7], sb.w, hb.w)};
          uint4 ul;
          __half2* l2 = reinterpret_cast<__half2*>(&ul);
#pragma unroll
          for (int e = 0; e < 4; ++e)
            l2[e] = __floats2half2_rn(x[2 * e] - __low2float(h2[e]), x[2 * e + 1] - __high2float(h2[e]));
          *reinterpret_cast<uint4*>(a.out_lo + opix * a.Cpad + q * 8) = ul;
        }
      }
    }
  }
  if (STATS) acc.flush(a, s_red);
}

__global__ void __launch_bounds__(256) raw_stats_kernel(const RawArgs a) {
  __shared__ float s_edges[kRawMaxBins + 1];
  __shared__ double s_red[8][kRawVals];
  if (threadIdx.x <= a.n_dd && a.n_dd > 0) s_edges[threadIdx.x] = a.edges[threadIdx.x];
  __syncthreads();
  RawAcc acc;
  acc.clear();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t po = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; po < a.n_pix; po += stride) {
    uint32_t w[3] = {0u, 0u, 0u};
    float2 d = make_float2(0.f, 0.f), t = make_float2(0.f, 0.f);
    bool flip = false;
    const int64_t p = a.pair_map ? mapped_pixel(a, po, flip) : po;
    if (a.use_rgb) {
      const ushort* r16 = reinterpret_cast<const ushort*>(a.rgb + p * 6);
      w[0] = r16[0]; w[1] = r16[1]; w[2] = r16[2];
    }
    if (a.depth) d = ld_depth_pair(a, p);
    if (a.use_td) t = __ldg(reinterpret_cast<const float2*>(a.td) + p);
    if (flip) {  // prev <-> cur: bytes [0 1 2 | 3 4 5] -> [3 4 5 | 0 1 2]
      const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
      w[0] = (w1 >> 8) | ((w2 & 0xff) << 8);
      w[1] = (w2 >> 8) | ((w0 & 0xff) << 8);
      w[2] = (w0 >> 8) | ((w1 & 0xff) << 8);
      d = make_float2(d.y, d.x);
      t = make_float2(t.y, t.x);
    }
    int b0 = -1, b1 = -1;
    if (a.n_dd > 0) {
      b0 = depth_bin(d.x, s_edges, a.n_dd);
      b1 = depth_bin(d.y, s_edges, a.n_dd);
    }
    acc.add(w, d, t, b0, b1);
  }
  acc.flush(a, s_red);
}

static int raw_check(const RawArgs& a) {
  const int cf = 3 * a.use_rgb + a.use_depth + a.n_dd + a.use_td;
  PNVO_REQUIRE(cf >= 1 && 2 * cf == a.C && a.C <= kMaxInC, "raw input: %d channels per frame do not match C=%d", cf, a.C);
  PNVO_REQUIRE(a.n_dd >= 0 && a.n_dd <= kRawMaxBins, "raw input: %d depth bins (max %d)", a.n_dd, kRawMaxBins);
  PNVO_REQUIRE(!a.use_rgb || a.rgb, "raw input: rgb missing");
  PNVO_REQUIRE(!(a.use_depth || a.n_dd) || a.depth, "raw input: depth missing");
  PNVO_REQUIRE(!a.n_dd || a.edges, "raw input: bin edges missing");
  PNVO_REQUIRE(!a.use_td || a.td, "raw input: top-down maps missing");
  return 0;
}

int raw_assemble_launch(const RawArgs& a, cudaStream_t st) {
  if (raw_check(a)) return -1;
  PNVO_REQUIRE(a.out && a.Cpad % 8 == 0 && a.Cpad <= kMaxInC && a.C + a.n_lo <= a.Cpad, "raw_assemble: bad output layout");
  PNVO_REQUIRE(a.n_lo == 0 || (a.n_lo == 2 && a.use_td && !a.out_lo), "raw_assemble: n_lo is 0, or 2 with top-down channels and no residual plane");
  if (a.n_pix <= 0) return 0;
  int blocks = static_cast<int>(std::min<int64_t>(ceil_div64(a.n_pix, 256), 148 * 16));
  if (a.stats) {
    // one pass: assembled tensor + batch statistics (exact-input stem, full 30-channel layout, no pair map)
    PNVO_REQUIRE(a.exact && !a.pair_map && !a.out_lo && a.use_rgb && a.use_depth && a.n_dd == 10 && a.use_td,
                 "raw_assemble: the fused statistics need the exact-input stem with the rgb + depth + 10 bins + top-down layout");
    // a thread must see <= 255 pixels (8-bit packed bin counts)
    blocks = static_cast<int>(std::min<int64_t>(std::max<int64_t>(blocks, ceil_div64(a.n_pix, 256 * 128)), ceil_div64(a.n_pix, 256)));
    raw_assemble_kernel<1, 1, 10, 1, false, false, true><<<blocks, 256, 0, st>>>(a);
    count_launch();
    return check_launch("raw_assemble+stats");
  }
  if (a.pair_map) {
    PNVO_REQUIRE(a.hw > 0 && a.n_pix % a.hw == 0 && a.n_pix < (1ll << 31), "raw_assemble: pair map needs pixels per sample");
    if (a.use_rgb && a.use_depth && a.n_dd == 10 && a.use_td) {
      if (a.out_lo) raw_assemble_kernel<1, 1, 10, 1, true, true><<<blocks, 256, 0, st>>>(a);
      else raw_assemble_kernel<1, 1, 10, 1, true><<<blocks, 256, 0, st>>>(a);
    } else {
      raw_assemble_kernel<0, 0, -1, 0, true><<<blocks, 256, 0, st>>>(a);
    }
  } else if (a.use_rgb && a.use_depth && a.n_dd == 10 && a.use_td) {
    if (a.out_lo) raw_assemble_kernel<1, 1, 10, 1, false, true><<<blocks, 256, 0, st>>>(a);
    else raw_assemble_kernel<1, 1, 10, 1><<<blocks, 256, 0, st>>>(a);
  } else if (a.use_rgb && a.use_depth && a.n_dd == 0 && !a.use_td) {
    if (a.out_lo) raw_assemble_kernel<1, 1, 0, 0, false, true><<<blocks, 256, 0, st>>>(a);
    else raw_assemble_kernel<1, 1, 0, 0><<<blocks, 256, 0, st>>>(a);
  } else {
    raw_assemble_kernel<0, 0, -1, 0><<<blocks, 256, 0, st>>>(a);
  }
  count_launch();
  return check_launch("raw_assemble");
}

int raw_stats_launch(const RawArgs& a, cudaStream_t st) {
  if (raw_check(a)) return -1;
  PNVO_REQUIRE(a.stats, "raw_stats: null accumulator");
  PNVO_REQUIRE(!a.pair_map || (a.hw > 0 && a.n_pix % a.hw == 0 && a.n_pix < (1ll << 31)), "raw_stats: pair map needs pixels per sample");
  if (a.n_pix <= 0) return 0;
  // a thread must see <= 255 pixels (8-bit packed bin counts; 32-bit integer rgb sums are exact far beyond): <= 128
  int64_t blocks = std::max<int64_t>(148 * 8, ceil_div64(a.n_pix, 256 * 128));
  blocks = std::min<int64_t>(blocks, ceil_div64(a.n_pix, 256));
  raw_stats_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(a);
  count_launch();
  return check_launch("raw_stats");
}

int raw_op(int code, const int32_t* i, const float* f, void* const* p, cudaStream_t st) {
  // p0 = rgb u8, p1 = depth, p2 = top-down, p3 = edges, p4 = scale, p5 = shift, p6 = out fp16 / fp64 stats, p7 = out_lo,
  // p8 = pair_map (int32 per output sample, nullable); i10 = pixels per sample (with pair_map); i11 = depth is fp16
  // i0 = use_rgb, i1 = use_depth, i2 = n_dd, i3 = use_td, i4 = C, i5 = Cpad, i6|i7 = n_pix, i8 = row_w, i9 = out_pitch,
  // i12 = n_lo (exact-input stem: top-down residual channels), i13 = exact (constant storage maps), p9 = fused statistics
  (void)f;
  RawArgs a{};
  a.rgb = static_cast<const uint8_t*>(p[0]);
  a.depth = p[1];
  a.depth_fp16 = i[11];
  a.td = static_cast<const float*>(p[2]);
  a.edges = static_cast<const float*>(p[3]);
  a.scale = static_cast<const float*>(p[4]);
  a.shift = static_cast<const float*>(p[5]);
  a.use_rgb = i[0]; a.use_depth = i[1]; a.n_dd = i[2]; a.use_td = i[3]; a.C = i[4]; a.Cpad = i[5];
  a.n_pix = (static_cast<int64_t>(static_cast<uint32_t>(i[7])) << 32) | static_cast<uint32_t>(i[6]);
  a.row_w = i[8]; a.out_pitch = i[9];
  a.pair_map = static_cast<const int32_t*>(p[8]); a.hw = i[10];
  a.n_lo = i[12];
  a.exact = i[13];
  if (code == PNVO_OP_RAW_ASSEMBLE) {
    a.out = static_cast<__half*>(p[6]);
    a.out_lo = static_cast<__half*>(p[7]);
    a.stats = static_cast<double*>(p[9]);   // nullable: fused batch statistics (exact-input stem)
    return raw_assemble_launch(a, st);
  }
  a.stats = static_cast<double*>(p[6]);
  return raw_stats_launch(a, st);
}

}  // namespace pnvo
