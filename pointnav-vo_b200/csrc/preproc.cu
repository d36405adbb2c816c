// HBM-bound preprocessing kernels of the PointNav-VO hot path (sm_100a):
//   depth discretisation (a7), egocentric top-down projection (a8), GAE / return scan (a13),
//   batched goal update (a10).  All arithmetic that feeds an index map uses explicit round-to-nearest
//   intrinsics (no FMA contraction) so the maps are bit-identical to the reference's fp32 torch path.
#include "common.cuh"
#include "elem.cuh"

namespace pnvo {

// ------------------------------------------------------------------------------------------------
// a7: depth discretisation (base_trainer_with_vo.py:135-167)
//   bin i  <=>  d >= e_i && d < e_{i+1}; the last bin is closed at e_n = 1.0.  For d in [0,1] this is
//   idx = #{ i in 1..n-1 : d >= e_i }.  One warp handles 32 consecutive pixels; the one-hot rows are
//   written with lanes striding over the 32*n_ch contiguous floats (bin index fetched by shuffle).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) discretize_kernel(const float* __restrict__ depth, int64_t n_pix,
                                                         const float* __restrict__ edges, int n_ch,
                                                         float* __restrict__ onehot, int64_t stride,
                                                         uint8_t* __restrict__ index, int32_t* err_count) {
  __shared__ float s_edges[65];
  for (int i = threadIdx.x; i <= n_ch; i += blockDim.x) s_edges[i] = edges[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const float lo = s_edges[0], hi = s_edges[n_ch];
  for (int64_t p0 = warp_global * 32; p0 < n_pix; p0 += n_warps * 32) {
    const int64_t p = p0 + lane;
    int idx = 255;
    if (p < n_pix) {
      const float d = __ldg(depth + p);
      if (d >= lo && d <= hi) {
        idx = 0;
        for (int i = 1; i < n_ch; ++i) idx += (d >= s_edges[i]) ? 1 : 0;
      } else if (err_count) {
        atomicAdd(err_count, 1);
      }
      if (index) index[p] = static_cast<uint8_t>(idx);
    }
    if (onehot) {
      const int n_valid = static_cast<int>(min(static_cast<int64_t>(32), n_pix - p0));
      if (stride == n_ch) {
        // fully contiguous rows: coalesced stores over the 32*n_ch floats of this warp
        float* base = onehot + p0 * stride;
        for (int f = lane; f < 32 * n_ch; f += 32) {
          const int pp = f / n_ch;
          const int c = f - pp * n_ch;
          const int bin = __shfl_sync(0xffffffffu, idx, pp);
          if (pp < n_valid) base[f] = (bin == c) ? 1.0f : 0.0f;
        }
      } else {
        for (int f = lane; f < 32 * n_ch; f += 32) {
          const int pp = f / n_ch;
          const int c = f - pp * n_ch;
          const int bin = __shfl_sync(0xffffffffu, idx, pp);
          if (pp < n_valid) onehot[(p0 + pp) * stride + c] = (bin == c) ? 1.0f : 0.0f;
        }
      }
    }
  }
}

// index-only form (no one-hot rows): four pixels per thread, one 16-byte load and one 4-byte store (the scalar kernel's
// byte stores held it at 19 % of the HBM peak on 5 B per pixel)
__global__ void __launch_bounds__(256) discretize_index4_kernel(const float4* __restrict__ depth, int64_t n4,
                                                                const float* __restrict__ edges, int n_ch,
                                                                uchar4* __restrict__ index, int32_t* err_count) {
  __shared__ float s_edges[65];
  for (int i = threadIdx.x; i <= n_ch; i += blockDim.x) s_edges[i] = edges[i];
  __syncthreads();
  const float lo = s_edges[0], hi = s_edges[n_ch];
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < n4;
       p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 d4 = __ldg(depth + p);
    const float d[4] = {d4.x, d4.y, d4.z, d4.w};
    int idx[4];
    int bad = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      idx[k] = 255;
      if (d[k] >= lo && d[k] <= hi) {
        idx[k] = 0;
        for (int i = 1; i < n_ch; ++i) idx[k] += (d[k] >= s_edges[i]) ? 1 : 0;
      } else {
        ++bad;
      }
    }
    if (bad && err_count) atomicAdd(err_count, bad);
    index[p] = make_uchar4(static_cast<unsigned char>(idx[0]), static_cast<unsigned char>(idx[1]),
                           static_cast<unsigned char>(idx[2]), static_cast<unsigned char>(idx[3]));
  }
}

// ------------------------------------------------------------------------------------------------
// a8: top-down projection (geometry_utils.py:516-721, Torch fp32 variant), one CTA per frame.
//   phase 1: bounding box of non-zero depth (rows/cols with any element > 0), float4 loads + smem flags
//   phase 2: 3x3 blur (cv2 GaussianBlur ksize 3, sigma 0, zero border: ((a+c)*.25 + b*.5), rows first),
//            only for the <=100 centre rows of the crop; unproject; histogram in shared memory
//            (uint16 pairs packed in uint32 words, 192*341*2 B = 131 KB)
//   phase 3: block max, out = count / max (fp32 division), coalesced store
//   Values outside the crop are zero by construction of the bbox, so zero padding at the crop edge ==
//   reading the frame itself with zero padding at the frame edge.
// ------------------------------------------------------------------------------------------------
struct TopDownArgs {
  const void* depth;  // fp32, or fp16 (the dataset's storage type) with the templated kernel
  int64_t in_stride;
  int64_t in_pix;  // floats between consecutive pixels of one frame (1, or 2 for a [.., 2] depth-pair tensor)
  int in_group, out_group;  // frames interleaved per group: frame n starts at (n / g) * stride + n % g
  int H, W;
  const float* ray;
  pnvo_topdown_consts k;
  float* out;
  int64_t out_frame_stride, out_pix_stride;
  int32_t* count;
  int band_rows;  // rows of the crop whose horizontal blur is staged in shared memory at a time
};

__device__ __forceinline__ float td_ld(const float* p) { return __ldg(p); }
__device__ __forceinline__ float td_ld(const __half* p) { return __half2float(__ldg(p)); }  // exact widening

template <typename T>
__device__ __forceinline__ float blur_h(const T* row, int c, int W, int64_t ps) {
  const float a = (c > 0) ? td_ld(row + (c - 1) * ps) : 0.0f;
  const float b = td_ld(row + c * ps);
  const float d = (c + 1 < W) ? td_ld(row + (c + 1) * ps) : 0.0f;
  return __fadd_rn(__fmul_rn(__fadd_rn(a, d), 0.25f), __fmul_rn(b, 0.5f));
}

template <typename T>
__global__ void __launch_bounds__(1024, 1) topdown_kernel(TopDownArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int H = a.H, W = a.W;
  const int n_cells = H * W;
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw);       // (n_cells+1)/2 words
  const int hist_words = (n_cells + 1) >> 1;
  int* s_row_any = reinterpret_cast<int*>(hist + hist_words);   // H
  int* s_col_any = s_row_any + H;                               // W
  __shared__ int s_bbox[4];
  __shared__ int s_max;
  const int tid = threadIdx.x, nt = blockDim.x;
  const T* __restrict__ D = static_cast<const T*>(a.depth) + static_cast<int64_t>(blockIdx.x / a.in_group) * a.in_stride + blockIdx.x % a.in_group;
  const int64_t ps = a.in_pix;

  for (int i = tid; i < hist_words; i += nt) hist[i] = 0u;
  for (int i = tid; i < H + W; i += nt) s_row_any[i] = 0;
  if (tid == 0) {
    s_bbox[0] = H; s_bbox[1] = -1; s_bbox[2] = W; s_bbox[3] = -1;
    s_max = 0;
  }
  __syncthreads();
  // phase 1: any(depth > 0) per row / column (depth >= 0, so fp32 sum > 0 <=> any element > 0).  A warp owns whole
  // rows (lanes over columns): the row flag is one ballot per 32 columns, the column flags are OR-ed in registers over
  // the warp's rows and written once -- no per-pixel index arithmetic or shared-memory traffic (ncu: the kernel was
  // instruction-issue bound, 236 instructions per pixel).
  {
    const int lane = tid & 31, wid = tid >> 5, n_warps = nt >> 5;
    constexpr int kColChunks = 16;  // up to 512 columns
    unsigned col_any = 0u;          // bit k: column lane + 32k of some row of this warp is > 0
    for (int r = wid; r < H; r += n_warps) {
      const T* row = D + static_cast<int64_t>(r) * W * ps;
      bool any = false;
#pragma unroll 4
      for (int k = 0; k < kColChunks; ++k) {
        const int c = lane + 32 * k;
        if (c < W && td_ld(row + static_cast<int64_t>(c) * ps) > 0.0f) {
          any = true;
          col_any |= 1u << k;
        }
      }
      if (__any_sync(0xffffffffu, any) && lane == 0) s_row_any[r] = 1;
    }
    for (int k = 0; k < kColChunks; ++k) {
      const int c = lane + 32 * k;
      if (c < W && ((col_any >> k) & 1u)) s_col_any[c] = 1;
    }
  }
  __syncthreads();
  for (int i = tid; i < H; i += nt)
    if (s_row_any[i]) { atomicMin(&s_bbox[0], i); atomicMax(&s_bbox[1], i); }
  for (int i = tid; i < W; i += nt)
    if (s_col_any[i]) { atomicMin(&s_bbox[2], i); atomicMax(&s_bbox[3], i); }
  __syncthreads();
  const int r0 = s_bbox[0], r1 = s_bbox[1], c0 = s_bbox[2], c1 = s_bbox[3];
  float* __restrict__ out = a.out + static_cast<int64_t>(blockIdx.x / a.out_group) * a.out_frame_stride + blockIdx.x % a.out_group;
  int32_t* cnt_out = a.count ? a.count + static_cast<int64_t>(blockIdx.x) * n_cells : nullptr;
  if (r1 < 0) {  // geometry_utils.py:519-525: empty frame -> all-zero map
    for (int i = tid; i < n_cells; i += nt) {
      out[static_cast<int64_t>(i) * a.out_pix_stride] = 0.0f;
      if (cnt_out) cnt_out[i] = 0;
    }
    return;
  }
  // phase 2: rows of the crop that are projected (:608-620)
  const int h = r1 - r0 + 1, w = c1 - c0 + 1;
  int ra, rb;
  if (a.k.center_crop) {
    const int c = (h + 1) >> 1;  // ceil(h / 2)
    ra = max(0, c - a.k.rows_around_center);
    rb = min(h, c + a.k.rows_around_center);
  } else {
    ra = 0;
    rb = min(a.k.rows_around_center * 2, h);
  }
  const float fH = static_cast<float>(H), fW = static_cast<float>(W);
  // The horizontal pass is computed ONCE per (row, column) into shared memory, band by band (band + 2 rows of the crop
  // fit next to the histogram), and the vertical pass reads three shared-memory values: 3.1 global loads per point
  // instead of 9, and no per-point index division (a warp owns whole rows).  Same operations in the same order as the
  // per-point version, so the counts stay bit-exact.
  // (warp-aggregating the histogram atomics with match.any was measured slower: 0.36 vs 0.335 ms for 512 frames)
  float* __restrict__ hb = reinterpret_cast<float*>(s_col_any + W);
  const int lane = tid & 31, wid = tid >> 5, n_warps = nt >> 5;
  const int band = a.band_rows;
  for (int b0 = ra; b0 < rb; b0 += band) {
    const int nb = min(band, rb - b0);
    for (int k = wid; k < nb + 2; k += n_warps) {
      const int r = r0 + b0 - 1 + k;  // frame row; rows outside the crop are zero (== zero border of the blur)
      const bool in = r >= r0 && r <= r1;
      const T* row = D + static_cast<int64_t>(r) * W * ps;
      for (int cc = lane; cc < w; cc += 32) hb[k * w + cc] = in ? blur_h(row, c0 + cc, W, ps) : 0.0f;
    }
    __syncthreads();
    for (int k = wid; k < nb; k += n_warps) {
      const float* h3 = hb + k * w;
      for (int cc = lane; cc < w; cc += 32) {
        const float hm = h3[cc], h0 = h3[w + cc], hp = h3[2 * w + cc];
        const float v = __fadd_rn(__fmul_rn(__fadd_rn(hm, hp), 0.25f), __fmul_rn(h0, 0.5f));
        const float z = __fadd_rn(__fmul_rn(v, a.k.depth_scale), a.k.depth_off);  // :558-560
        const float x = __fmul_rn(__ldg(a.ray + c0 + cc), z);                      // :656
        const float nx = __fdiv_rn(__fsub_rn(x, a.k.min_x), a.k.x_den);            // :679
        const float nz = __fdiv_rn(__fsub_rn(z, a.k.depth_off), a.k.z_den);        // :680
        const float frow = __fsub_rn(fH, ceilf(__fmul_rn(fH, nz)));                // :689-691
        const float fcol = floorf(__fmul_rn(fW, nx));                              // :692
        if (frow >= 0.0f && frow < fH && fcol >= 0.0f && fcol < fW) {              // :706-711
          const int cell = static_cast<int>(frow) * W + static_cast<int>(fcol);
          atomicAdd(&hist[cell >> 1], (cell & 1) ? 0x10000u : 1u);
        }
      }
    }
    __syncthreads();
  }
  __syncthreads();
  // phase 3: max + normalise
  int m = 0;
  for (int i = tid; i < hist_words; i += nt) {
    const uint32_t wv = hist[i];
    m = max(m, static_cast<int>(max(wv & 0xFFFFu, wv >> 16)));
  }
  m = warp_max_i(m);
  if ((tid & 31) == 0) atomicMax(&s_max, m);
  __syncthreads();
  const int mx = s_max;
  const float fm = static_cast<float>(mx);
  for (int i = tid; i < n_cells; i += nt) {
    const uint32_t wv = hist[i >> 1];
    const int c = (i & 1) ? static_cast<int>(wv >> 16) : static_cast<int>(wv & 0xFFFFu);
    float o = 0.0f;
    if (c > 0) o = fminf(__fdiv_rn(static_cast<float>(c), fm), 1.0f);  // c > 0 implies mx > 0; most cells are empty
    out[static_cast<int64_t>(i) * a.out_pix_stride] = o;
    if (cnt_out) cnt_out[i] = c;
  }
}

// ------------------------------------------------------------------------------------------------
// a13: GAE / discounted returns (rollout_storage.py:102-120)
// mode 0: lanes = envs (coalesced along N), time loop sequential with the reference's rounding order:
//   delta = (r + (g*V[t+1])*m[t+1]) - V[t];  gae = delta + ((g*tau)*m[t+1])*gae;  ret = gae + V[t]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) gae_seq_kernel(const float* __restrict__ rewards, float* __restrict__ values,
                                                      const float* __restrict__ masks,
                                                      const float* __restrict__ next_value,
                                                      float* __restrict__ returns, int T, int N, int use_gae, float g,
                                                      float gt) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float nv = next_value[n];
  if (use_gae) {
    values[static_cast<int64_t>(T) * N + n] = nv;  // (returns[T] is left untouched, as the reference does)
    if (T == 0) return;
    float gae = 0.0f;
    float v_next = nv;
    // software-pipelined: loads of step t-1 are independent of the recurrence
    float r = rewards[static_cast<int64_t>(T - 1) * N + n];
    float m = masks[static_cast<int64_t>(T) * N + n];
    float v = values[static_cast<int64_t>(T - 1) * N + n];
    for (int t = T - 1; t >= 0; --t) {
      float r_n = 0.f, m_n = 0.f, v_n = 0.f;
      if (t > 0) {
        r_n = rewards[static_cast<int64_t>(t - 1) * N + n];
        m_n = masks[static_cast<int64_t>(t) * N + n];
        v_n = values[static_cast<int64_t>(t - 1) * N + n];
      }
      float delta = __fadd_rn(r, __fmul_rn(__fmul_rn(g, v_next), m));
      delta = __fsub_rn(delta, v);
      gae = __fadd_rn(delta, __fmul_rn(__fmul_rn(gt, m), gae));
      returns[static_cast<int64_t>(t) * N + n] = __fadd_rn(gae, v);
      v_next = v;
      r = r_n; m = m_n; v = v_n;
    }
  } else {
    float ret = nv;
    returns[static_cast<int64_t>(T) * N + n] = nv;
    for (int t = T - 1; t >= 0; --t) {
      const float m = masks[static_cast<int64_t>(t + 1) * N + n];
      const float r = rewards[static_cast<int64_t>(t) * N + n];
      ret = __fadd_rn(__fmul_rn(__fmul_rn(ret, g), m), r);
      returns[static_cast<int64_t>(t) * N + n] = ret;
    }
  }
}

// mode 1: warp-scan over time.  The recurrence y_t = b_t + a_t * y_{t+1} is the composition of affine
// maps f_t(y) = a_t*y + b_t; one warp owns one env, lane l owns the contiguous time chunk
// [l*C, (l+1)*C), composes its chunk locally, then a reversed inclusive shuffle scan composes across
// lanes.  Same result as mode 0 up to fp32 re-association.
__global__ void __launch_bounds__(128) gae_scan_kernel(const float* __restrict__ rewards, float* __restrict__ values,
                                                       const float* __restrict__ masks,
                                                       const float* __restrict__ next_value,
                                                       float* __restrict__ returns, int T, int N, int use_gae,
                                                       float g, float gt) {
  const int lane = threadIdx.x & 31;
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (n >= N) return;
  const float nv = next_value[n];
  const int C = (T + 31) / 32;
  const int t_lo = lane * C, t_hi = min(T, t_lo + C);
  if (lane == 0) {
    if (use_gae) {
      values[static_cast<int64_t>(T) * N + n] = nv;
    } else {
      returns[static_cast<int64_t>(T) * N + n] = nv;
    }
  }
  // local composition over the chunk, from t_hi-1 down to t_lo: y_{t_lo} = A * y_{t_hi} + Bc
  float A = 1.0f, Bc = 0.0f;
  for (int t = t_hi - 1; t >= t_lo; --t) {
    const float m = masks[static_cast<int64_t>(t + 1) * N + n];
    float a_t, b_t;
    if (use_gae) {
      const float vn = (t + 1 == T) ? nv : values[static_cast<int64_t>(t + 1) * N + n];
      a_t = gt * m;
      b_t = rewards[static_cast<int64_t>(t) * N + n] + g * vn * m - values[static_cast<int64_t>(t) * N + n];
    } else {
      a_t = g * m;
      b_t = rewards[static_cast<int64_t>(t) * N + n];
    }
    Bc = b_t + a_t * Bc;
    A = a_t * A;
  }
  // suffix scan across lanes: after it, (A, Bc) of lane l maps y_T-side boundary of lane 31 to y_{t_lo(l)}
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float A2 = __shfl_down_sync(0xffffffffu, A, o);
    const float B2 = __shfl_down_sync(0xffffffffu, Bc, o);
    if (lane + o < 32) {
      Bc = Bc + A * B2;
      A = A * A2;
    }
  }
  // boundary value entering this lane's chunk = y at t_hi = result of lane+1's suffix map applied to y_T
  const float yT = use_gae ? 0.0f : nv;
  float y_in = __shfl_down_sync(0xffffffffu, Bc + A * yT, 1);
  if (lane == 31) y_in = yT;
  float y = y_in;
  for (int t = t_hi - 1; t >= t_lo; --t) {
    const float m = masks[static_cast<int64_t>(t + 1) * N + n];
    if (use_gae) {
      const float vn = (t + 1 == T) ? nv : values[static_cast<int64_t>(t + 1) * N + n];
      const float v = values[static_cast<int64_t>(t) * N + n];
      const float delta = rewards[static_cast<int64_t>(t) * N + n] + g * vn * m - v;
      y = delta + gt * m * y;
      returns[static_cast<int64_t>(t) * N + n] = y + v;
    } else {
      y = y * g * m + rewards[static_cast<int64_t>(t) * N + n];
      returns[static_cast<int64_t>(t) * N + n] = y;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// a10: goal update (geometry_utils.py:115-144), fp64 like the reference's numpy-quaternion path
// ------------------------------------------------------------------------------------------------
__global__ void goal_update_kernel(double* __restrict__ goal, const float* __restrict__ delta,
                                   float* __restrict__ polar, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double dx = delta[3 * i + 0], dz = delta[3 * i + 1], dyaw = delta[3 * i + 2];
  const double vx = goal[3 * i + 0] - dx, vy = goal[3 * i + 1], vz = goal[3 * i + 2] - dz;
  const double c = cos(dyaw), s = sin(dyaw);
  const double x = c * vx - s * vz, z = s * vx + c * vz;
  goal[3 * i + 0] = x;
  goal[3 * i + 1] = vy;
  goal[3 * i + 2] = z;
  polar[2 * i + 0] = static_cast<float>(hypot(-z, x));
  polar[2 * i + 1] = static_cast<float>(-atan2(x, -z));
}

}  // namespace pnvo

namespace pnvo {
int topdown_launch(const void* depth, int64_t in_stride, int64_t in_pix_stride, int in_group, int out_group,
                   int n_frames, int H, int W, const float* ray, const pnvo_topdown_consts* consts, float* out,
                   int64_t out_frame_stride, int64_t out_pix_stride, int32_t* count, void* stream, int depth_fp16);
}
using namespace pnvo;

extern "C" int pnvo_discretize_depth(const float* depth, int64_t n_pix, const float* edges, int n_channels,
                                     float* onehot, int64_t onehot_stride, uint8_t* index, int32_t* err_count,
                                     void* stream) {
  if (n_pix <= 0) return 0;
  PNVO_REQUIRE(depth && edges, "discretize_depth: null input");
  PNVO_REQUIRE(n_channels >= 1 && n_channels <= 64, "discretize_depth: n_channels %d not in [1,64]", n_channels);
  PNVO_REQUIRE(!onehot || onehot_stride >= n_channels, "discretize_depth: onehot_stride < n_channels");
  if (n_pix <= 0) return 0;
  const int threads = 256;
  if (!onehot && index && n_pix % 4 == 0 && (reinterpret_cast<uintptr_t>(depth) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(index) & 3) == 0) {
    const int64_t n4 = n_pix / 4;
    const int blocks4 = static_cast<int>(std::min<int64_t>(ceil_div64(n4, threads), 148 * 16));
    discretize_index4_kernel<<<blocks4, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(depth), n4, edges, n_channels, reinterpret_cast<uchar4*>(index), err_count);
    count_launch();
    return check_launch("discretize_depth");
  }
  int64_t warps = ceil_div64(n_pix, 32);
  int blocks = static_cast<int>(std::min<int64_t>(ceil_div64(warps, threads / 32), 148 * 8));
  discretize_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(depth, n_pix, edges, n_channels, onehot,
                                                                               onehot_stride, index, err_count);
  count_launch();
  return check_launch("discretize_depth");
}

extern "C" int pnvo_topdown_project(const float* depth, int64_t in_stride, int n_frames, int H, int W,
                                    const float* ray, const pnvo_topdown_consts* consts, float* out,
                                    int64_t out_frame_stride, int64_t out_pix_stride, int32_t* count, void* stream) {
  return pnvo::topdown_launch(depth, in_stride, 1, 1, 1, n_frames, H, W, ray, consts, out, out_frame_stride,
                              out_pix_stride, count, stream, 0);
}

extern "C" int pnvo_topdown_project_strided(const float* depth, int64_t in_stride, int64_t in_pix_stride, int n_frames,
                                            int H, int W, const float* ray, const pnvo_topdown_consts* consts,
                                            float* out, int64_t out_frame_stride, int64_t out_pix_stride,
                                            int32_t* count, void* stream) {
  return pnvo::topdown_launch(depth, in_stride, in_pix_stride, static_cast<int>(in_pix_stride),
                              static_cast<int>(out_pix_stride), n_frames, H, W, ray, consts, out, out_frame_stride,
                              out_pix_stride, count, stream, 0);
}

extern "C" int pnvo_topdown_project_strided_f16(const uint16_t* depth, int64_t in_stride, int64_t in_pix_stride,
                                                int n_frames, int H, int W, const float* ray,
                                                const pnvo_topdown_consts* consts, float* out, int64_t out_frame_stride,
                                                int64_t out_pix_stride, int32_t* count, void* stream) {
  return pnvo::topdown_launch(depth, in_stride, in_pix_stride, static_cast<int>(in_pix_stride),
                              static_cast<int>(out_pix_stride), n_frames, H, W, ray, consts, out, out_frame_stride,
                              out_pix_stride, count, stream, 1);
}

namespace pnvo {
int topdown_launch(const void* depth, int64_t in_stride, int64_t in_pix_stride, int in_group, int out_group,
                   int n_frames, int H, int W, const float* ray, const pnvo_topdown_consts* consts, float* out,
                   int64_t out_frame_stride, int64_t out_pix_stride, int32_t* count, void* stream, int depth_fp16) {
  PNVO_REQUIRE(depth && ray && consts && out, "topdown_project: null argument");
  PNVO_REQUIRE(H > 0 && W > 0 && W <= 512 && static_cast<int64_t>(H) * W <= 110000, "topdown_project: frame %dx%d too large", H, W);
  PNVO_REQUIRE(consts->rows_around_center * 2 * W < 65535, "topdown_project: too many points for uint16 counts");
  if (n_frames <= 0) return 0;
  TopDownArgs a;
  PNVO_REQUIRE(in_pix_stride >= 1, "topdown_project: in_pix_stride");
  a.depth = depth; a.in_stride = in_stride; a.in_pix = in_pix_stride; a.in_group = in_group; a.out_group = out_group; a.H = H; a.W = W; a.ray = ray; a.k = *consts;
  a.out = out; a.out_frame_stride = out_frame_stride; a.out_pix_stride = out_pix_stride; a.count = count;
  const size_t smem_hist = static_cast<size_t>((H * W + 1) / 2) * 4 + static_cast<size_t>(H + W) * 4;
  // horizontal-blur staging: as many crop rows (+2 halo rows) as fit in what the histogram leaves of 224 KB
  const int rows_max = 2 * consts->rows_around_center;
  int band = static_cast<int>((224 * 1024 - smem_hist) / (static_cast<size_t>(W) * 4)) - 2;
  PNVO_REQUIRE(band >= 1, "topdown_project: no shared memory left for the blur rows (%dx%d)", H, W);
  if (band > rows_max) band = rows_max;
  band = (rows_max + ceil_div(rows_max, band) - 1) / ceil_div(rows_max, band);  // equal bands
  a.band_rows = band;
  const size_t smem = smem_hist + static_cast<size_t>(band + 2) * W * 4;
  static bool attr_set = false;
  if (!attr_set) {
    // 227 KB per CTA minus the kernel's few static words
    cudaError_t e = cudaFuncSetAttribute(topdown_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(topdown_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    PNVO_REQUIRE(e == cudaSuccess, "topdown_project: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  if (depth_fp16) topdown_kernel<__half><<<n_frames, 1024, smem, static_cast<cudaStream_t>(stream)>>>(a);
  else topdown_kernel<float><<<n_frames, 1024, smem, static_cast<cudaStream_t>(stream)>>>(a);
  count_launch();
  return check_launch("topdown_project");
}
}  // namespace pnvo

// ------------------------------------------------------------------------------------------------
// a14 / 8f-3: PPO clipped-surrogate + clipped-value loss and its gradient in one pass (rl/ppo/ppo.py:101-126).
//   ratio = exp(lp - old_lp); action_loss = -mean(min(ratio * adv, clamp(ratio, 1-c, 1+c) * adv))
//   value_loss = 0.5 * mean(max((v - R)^2, (vp + clamp(v - vp, -c, c) - R)^2))        (or 0.5 * mean((R - v)^2))
// losses[0] = value_loss, losses[1] = action_loss (atomics into a pre-zeroed pair);
// d_values = value_loss_coef * d value_loss / d v,  d_log_probs = d action_loss / d lp  -- ties of min / max split the
// gradient evenly and clamp passes it on its closed interval, as torch.min / torch.max / torch.clamp do.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ppo_loss_kernel(const float* __restrict__ values, const float* __restrict__ lp,
                                                       const float* __restrict__ value_preds,
                                                       const float* __restrict__ returns, const float* __restrict__ old_lp,
                                                       const float* __restrict__ adv, int64_t n, float clip,
                                                       int use_clipped_value, float value_coef,
                                                       float* __restrict__ losses, float* __restrict__ d_values,
                                                       float* __restrict__ d_lp) {
  __shared__ float s_v[8], s_a[8];
  const float inv_n = 1.f / static_cast<float>(n);
  float acc_v = 0.f, acc_a = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float a = adv[i];
    const float ratio = expf(lp[i] - old_lp[i]);
    const float rc = fminf(fmaxf(ratio, 1.f - clip), 1.f + clip);
    const float s1 = ratio * a, s2 = rc * a;
    acc_a -= fminf(s1, s2);
    const float in_range = (ratio >= 1.f - clip && ratio <= 1.f + clip) ? 1.f : 0.f;
    // d min(s1, s2) / d ratio: s1 carries a, s2 carries a * in_range; a tie splits evenly
    float g = (s1 < s2) ? a : ((s1 > s2) ? a * in_range : 0.5f * (a + a * in_range));
    d_lp[i] = -g * ratio * inv_n;
    const float v = values[i], R = returns[i];
    const float e1 = v - R;
    float dv;
    if (use_clipped_value) {
      const float vp = value_preds[i];
      const float dvp = v - vp;
      const float vc = vp + fminf(fmaxf(dvp, -clip), clip);
      const float e2 = vc - R;
      const float l1 = e1 * e1, l2 = e2 * e2;
      acc_v += 0.5f * fmaxf(l1, l2);
      const float inr = (dvp >= -clip && dvp <= clip) ? 1.f : 0.f;
      dv = (l1 > l2) ? e1 : ((l1 < l2) ? e2 * inr : 0.5f * (e1 + e2 * inr));
    } else {
      acc_v += 0.5f * e1 * e1;
      dv = e1;
    }
    d_values[i] = value_coef * dv * inv_n;
  }
  acc_v = warp_sum(acc_v);
  acc_a = warp_sum(acc_a);
  if ((threadIdx.x & 31) == 0) {
    s_v[threadIdx.x >> 5] = acc_v;
    s_a[threadIdx.x >> 5] = acc_a;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tv = 0.f, ta = 0.f;
    for (int k = 0; k < 8; ++k) { tv += s_v[k]; ta += s_a[k]; }
    atomicAdd(losses, tv * inv_n);
    atomicAdd(losses + 1, ta * inv_n);
  }
}

extern "C" int pnvo_ppo_loss(const float* values, const float* log_probs, const float* value_preds, const float* returns,
                             const float* old_log_probs, const float* adv_targ, int64_t n, float clip_param,
                             int use_clipped_value_loss, float value_loss_coef, float* losses, float* d_values,
                             float* d_log_probs, void* stream) {
  PNVO_REQUIRE(values && log_probs && returns && old_log_probs && adv_targ && losses && d_values && d_log_probs,
               "ppo_loss: null argument");
  PNVO_REQUIRE(!use_clipped_value_loss || value_preds, "ppo_loss: the clipped value loss needs value_preds");
  PNVO_REQUIRE(n > 0, "ppo_loss: empty minibatch");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (zero_launch(losses, 2 * sizeof(float), st)) return -1;
  const int blocks = static_cast<int>(std::min<int64_t>(ceil_div64(n, 256), 148));
  ppo_loss_kernel<<<blocks, 256, 0, st>>>(values, log_probs, value_preds, returns, old_log_probs, adv_targ, n, clip_param,
                                          use_clipped_value_loss, value_loss_coef, losses, d_values, d_log_probs);
  count_launch();
  return check_launch("ppo_loss");
}

extern "C" int pnvo_gae_scan(const float* rewards, float* value_preds, const float* masks, const float* next_value,
                             float* returns, int T, int N, int use_gae, float gamma, float gamma_tau, int mode,
                             void* stream) {
  PNVO_REQUIRE(rewards && value_preds && masks && next_value && returns, "gae_scan: null argument");
  PNVO_REQUIRE(T >= 0 && N >= 0, "gae_scan: negative size");
  if (N == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mode == 0) {
    gae_seq_kernel<<<ceil_div(N, 128), 128, 0, st>>>(rewards, value_preds, masks, next_value, returns, T, N, use_gae,
                                                     gamma, gamma_tau);
  } else {
    gae_scan_kernel<<<ceil_div(N * 32, 128), 128, 0, st>>>(rewards, value_preds, masks, next_value, returns, T, N,
                                                           use_gae, gamma, gamma_tau);
  }
  count_launch();
  return check_launch("gae_scan");
}

extern "C" int pnvo_goal_update(double* goal_xyz, const float* delta, float* polar, int n, void* stream) {
  PNVO_REQUIRE(goal_xyz && delta && polar, "goal_update: null argument");
  if (n <= 0) return 0;
  goal_update_kernel<<<ceil_div(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(goal_xyz, delta, polar, n);
  count_launch();
  return check_launch("goal_update");
}
