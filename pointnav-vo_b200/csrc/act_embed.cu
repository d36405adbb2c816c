// Action-embedding branch of the act-embed VO variants (vo_cnn_act_embed.py:36,65-75): the hidden layer sees
// [flattened visual features | Embedding(action)] behind one Dropout.  The visual part of that Linear runs on the
// tensor cores (fc as a 1x1 conv); the 32 embedding columns are a rank-32 update done here in fp32:
//   fwd : e[b][j] = E[action_b][j] * mask[b][j]            (mask = inverted-dropout keep/scale, 1 in eval mode)
//         z[b][n] += sum_j W[n][col0 + j] * e[b][j]
//   bwd : dW[n][col0 + j] = sum_b dz[b][n] * e[b][j]
//         dE[a][j]       += sum_{b: action_b = a} mask[b][j] * sum_n dz[b][n] * W[n][col0 + j]
#include "common.cuh"
#include "elem.cuh"

namespace pnvo {

__device__ __forceinline__ uint32_t ae_hash_u32(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return static_cast<uint32_t>(x >> 16);
}

__global__ void __launch_bounds__(256) act_embed_fwd_kernel(float* __restrict__ z, const float* __restrict__ W, int w_ld,
                                                            int col0, const float* __restrict__ E, int n_rows,
                                                            const int64_t* __restrict__ actions, int hidden, int dim,
                                                            float* __restrict__ e_used, float* __restrict__ mask,
                                                            const uint64_t* __restrict__ seed, uint32_t thresh,
                                                            float scale) {
  extern __shared__ float s_e[];
  const int b = blockIdx.x;
  int64_t a = actions[b];
  a = a < 0 ? 0 : (a >= n_rows ? n_rows - 1 : a);
  for (int j = threadIdx.x; j < dim; j += blockDim.x) {
    float m = 1.f;
    if (thresh != 0u) {
      const uint64_t s0 = *seed + 2ull * 0x9E3779B97F4A7C15ULL;  // dropout site 2 (0 / 1 = feature map / hidden)
      m = ae_hash_u32(s0 + static_cast<uint64_t>(b * dim + j) * 0xD6E8FEB86659FD93ULL) >= thresh ? scale : 0.f;
    }
    const float e = E[a * dim + j] * m;
    s_e[j] = e;
    e_used[b * dim + j] = e;
    mask[b * dim + j] = m;
  }
  __syncthreads();
  for (int n = threadIdx.x; n < hidden; n += blockDim.x) {
    const float* w = W + static_cast<int64_t>(n) * w_ld + col0;
    float acc = 0.f;
    for (int j = 0; j < dim; ++j) acc = fmaf(w[j], s_e[j], acc);
    z[static_cast<int64_t>(b) * hidden + n] += acc;
  }
}

__global__ void __launch_bounds__(256) act_embed_bwd_w_kernel(const __half* __restrict__ dz, int dz_ld,
                                                              const float* __restrict__ e_used, int B, int hidden,
                                                              int dim, float* __restrict__ dW, int w_ld, int col0) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= hidden * dim) return;
  const int n = idx / dim, j = idx - n * dim;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) acc = fmaf(__half2float(dz[static_cast<int64_t>(b) * dz_ld + n]), e_used[b * dim + j], acc);
  dW[static_cast<int64_t>(n) * w_ld + col0 + j] = acc;
}

__global__ void __launch_bounds__(256) act_embed_bwd_e_kernel(const __half* __restrict__ dz, int dz_ld,
                                                              const float* __restrict__ W, int w_ld, int col0,
                                                              const float* __restrict__ mask,
                                                              const int64_t* __restrict__ actions, int n_rows, int B,
                                                              int hidden, int dim, float* __restrict__ dE) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * dim) return;
  const int b = idx / dim, j = idx - b * dim;
  float acc = 0.f;
  for (int n = 0; n < hidden; ++n)
    acc = fmaf(__half2float(dz[static_cast<int64_t>(b) * dz_ld + n]), W[static_cast<int64_t>(n) * w_ld + col0 + j], acc);
  int64_t a = actions[b];
  a = a < 0 ? 0 : (a >= n_rows ? n_rows - 1 : a);
  atomicAdd(dE + a * dim + j, acc * mask[idx]);
}

int act_embed_op(int code, const int32_t* i, const float* f, void* const* p, cudaStream_t st) {
  // p0 = z fp32 [B][hidden] (fwd) / dz fp16 [B][dz_ld] (bwd), p1 = W fp32 [hidden][w_ld], p2 = E fp32 [n_rows][dim],
  // p3 = actions int64 [B], p4 = e_used [B][dim], p5 = mask [B][dim], p6 = seed (fwd) / dW (bwd), p7 = dE (bwd, pre-zeroed)
  // i0 = B, i1 = hidden, i2 = dim, i3 = n_rows, i4 = w_ld, i5 = col0, i6 = dz_ld; f0 = dropout p
  const int B = i[0], hidden = i[1], dim = i[2], n_rows = i[3], w_ld = i[4], col0 = i[5], dz_ld = i[6];
  PNVO_REQUIRE(B >= 0 && hidden > 0 && dim > 0 && dim <= 1024 && n_rows > 0 && col0 + dim <= w_ld, "act_embed: bad sizes");
  for (int k = 1; k <= 5; ++k) PNVO_REQUIRE(p[k], "act_embed: null pointer %d", k);
  if (B == 0) return 0;
  if (code == PNVO_OP_ACT_EMBED_FWD) {
    const float pdrop = f[0];
    PNVO_REQUIRE(p[0] && (pdrop == 0.f || p[6]) && pdrop >= 0.f && pdrop < 1.f, "act_embed_fwd: bad arguments");
    const uint32_t thresh = static_cast<uint32_t>(static_cast<double>(pdrop) * 4294967296.0);
    act_embed_fwd_kernel<<<B, 256, dim * sizeof(float), st>>>(
        static_cast<float*>(p[0]), static_cast<const float*>(p[1]), w_ld, col0, static_cast<const float*>(p[2]), n_rows,
        static_cast<const int64_t*>(p[3]), hidden, dim, static_cast<float*>(p[4]), static_cast<float*>(p[5]),
        static_cast<const uint64_t*>(p[6]), thresh, 1.f / (1.f - pdrop));
    count_launch();
    return check_launch("act_embed_fwd");
  }
  PNVO_REQUIRE(p[0] && p[6] && p[7] && dz_ld >= hidden, "act_embed_bwd: bad arguments");
  act_embed_bwd_w_kernel<<<ceil_div(hidden * dim, 256), 256, 0, st>>>(
      static_cast<const __half*>(p[0]), dz_ld, static_cast<const float*>(p[4]), B, hidden, dim, static_cast<float*>(p[6]),
      w_ld, col0);
  act_embed_bwd_e_kernel<<<ceil_div(B * dim, 256), 256, 0, st>>>(
      static_cast<const __half*>(p[0]), dz_ld, static_cast<const float*>(p[1]), w_ld, col0, static_cast<const float*>(p[5]),
      static_cast<const int64_t*>(p[3]), n_rows, B, hidden, dim, static_cast<float*>(p[7]));
  count_launch(2);
  return check_launch("act_embed_bwd");
}

}  // namespace pnvo
