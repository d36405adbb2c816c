// Data-parallel gradient exchange fused with the optimiser, over NVLink peer memory (the reference's
// counterpart is DistributedDataParallel's bucketed NCCL all-reduce followed by torch.optim.Adam: rl/ddppo/algo/ddppo.py:55-96,
// vo_cnn_regression_geo_invariance_engine.py:122-133).
//
// Every rank owns one contiguous slice of the flat fp32 parameter bucket.  ONE kernel per rank and step does
//   reduce-scatter : g[i] = sum over ranks p (fixed order) of grad_p[i], read straight from the peers' gradient buckets
//                    (16-byte loads over NVLink, all `world` loads of an element in flight together);
//   Adam           : torch.optim.Adam on the slice (moments m, v are touched on the owning rank only);
//   all-gather     : the updated parameters are stored into EVERY rank's parameter bucket (16-byte stores over NVLink).
// Nothing is staged: no reduced-gradient buffer, no separate optimiser pass, no second collective, and because every
// element is reduced and updated by exactly one rank, the replicas stay bit-identical by construction.
// Traffic per rank: (world-1)/world of the bucket in and the same out -- for the 15.8 MB ResNet-18 bucket on 8 GPUs
// 13.8 MB each way, ~25 us at NVLink 5 rates, against ~250 us for the NCCL all-reduce at this (latency-bound) size.
//
// Synchronisation: two sets of per-peer sequence flags in peer memory, written with system-scope releases.
//   entry : "my gradients are final" -> every peer; wait until every peer said so (the kernel is stream-ordered behind the
//           rank's own backward pass, so the gradients are complete when it starts).
//   exit  : the last CTA of the grid (device-scope ticket) tells every peer "I have read your gradients and written my
//           slice into your parameters", then waits for the same message from every peer before the kernel ends: the next
//           kernel on any rank may overwrite its gradient bucket and read its parameters.
// Flags carry the step sequence number (monotonic), so nothing is ever reset.  A poll that sees no progress for ~4 s sets
// an error word instead of hanging the GPU.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace pnvo {

static constexpr int kMaxPeers = 8;

struct PeerAdamArgs {
  const float* grad[kMaxPeers];   // every rank's gradient bucket (own entry = local pointer)
  float* param[kMaxPeers];        // every rank's parameter bucket
  uint32_t* flags[kMaxPeers];     // every rank's flag block: [0..7] entry flags, [8..15] exit flags, [16] ticket, [17] error
  float* m;
  float* v;
  int64_t n4;                     // bucket length in float4 (the bucket is padded to a multiple of 4 floats)
  int64_t slice4;                 // float4 per rank slice
  int rank, world;
  uint32_t seq;
  float lr, b1, b2, eps, bc1, bc2_sqrt;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(float4* p, const float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// waits until flag >= seq (wrap-safe); false on timeout
__device__ __forceinline__ bool wait_flag(const uint32_t* f, uint32_t seq) {
  const long long t0 = clock64();
  while (static_cast<int32_t>(ld_acquire_sys(f) - seq) < 0) {
    if (clock64() - t0 > 8000000000LL) return false;  // ~4 s at 2 GHz
    __nanosleep(100);
  }
  return true;
}

__device__ __forceinline__ float adam1(float& p, float g, float& m, float& v, const PeerAdamArgs& a) {
  m = a.b1 * m + (1.f - a.b1) * g;
  v = a.b2 * v + (1.f - a.b2) * g * g;
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p -= (a.lr / a.bc1) * (m / denom);
  return p;
}

template <int WORLD>
__global__ void __launch_bounds__(256) peer_reduce_adam_kernel(const PeerAdamArgs a) {
  uint32_t* my_flags = a.flags[a.rank];
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  // ---- entry: announce (one CTA), then every CTA waits on the LOCAL flag block
  if (blockIdx.x == 0 && threadIdx.x < WORLD) st_release_sys(a.flags[threadIdx.x] + a.rank, a.seq);
  if (threadIdx.x < WORLD) {
    if (!wait_flag(my_flags + threadIdx.x, a.seq)) {
      s_ok = 0;
      my_flags[17] = 1;
    }
  }
  __syncthreads();
  if (s_ok) {
    const int64_t lo = a.slice4 * a.rank;
    const int64_t hi = min(lo + a.slice4, a.n4);
    for (int64_t i = lo + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < hi;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
      float4 g[WORLD];
#pragma unroll
      for (int p = 0; p < WORLD; ++p) g[p] = ld_peer(reinterpret_cast<const float4*>(a.grad[p]) + i);
      float4 w = reinterpret_cast<const float4*>(a.param[a.rank])[i];
      float4 m = reinterpret_cast<const float4*>(a.m)[i];
      float4 v = reinterpret_cast<const float4*>(a.v)[i];
      float4 s = g[0];
#pragma unroll
      for (int p = 1; p < WORLD; ++p) {  // fixed order: the result does not depend on which rank owns the slice
        s.x += g[p].x;
        s.y += g[p].y;
        s.z += g[p].z;
        s.w += g[p].w;
      }
      adam1(w.x, s.x, m.x, v.x, a);
      adam1(w.y, s.y, m.y, v.y, a);
      adam1(w.z, s.z, m.z, v.z, a);
      adam1(w.w, s.w, m.w, v.w, a);
      reinterpret_cast<float4*>(a.m)[i] = m;
      reinterpret_cast<float4*>(a.v)[i] = v;
#pragma unroll
      for (int p = 0; p < WORLD; ++p) st_peer(reinterpret_cast<float4*>(a.param[p]) + i, w);
    }
  }
  // ---- exit: every thread's peer stores are ordered before the ticket; the last CTA signals and waits
  __threadfence_system();
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    const uint32_t t = atomicAdd(my_flags + 16, 1u);
    s_last = (t == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (threadIdx.x == 0) my_flags[16] = 0;  // ticket ready for the next step (ordered by the kernel boundary)
  if (threadIdx.x < WORLD) {
    st_release_sys(a.flags[threadIdx.x] + 8 + a.rank, a.seq);
    if (!wait_flag(my_flags + 8 + threadIdx.x, a.seq)) my_flags[17] = 1;
  }
}

int peer_reduce_adam_launch(const PeerAdamArgs& a, cudaStream_t st) {
  PNVO_REQUIRE(a.world >= 2 && a.world <= kMaxPeers, "peer_reduce_adam: world must be 2..8");
  PNVO_REQUIRE(a.rank >= 0 && a.rank < a.world, "peer_reduce_adam: bad rank");
  const int64_t work = std::min<int64_t>(a.slice4, a.n4);
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div64(work, 256), 148 * 4)));
  switch (a.world) {
    case 2: peer_reduce_adam_kernel<2><<<grid, 256, 0, st>>>(a); break;
    case 3: peer_reduce_adam_kernel<3><<<grid, 256, 0, st>>>(a); break;
    case 4: peer_reduce_adam_kernel<4><<<grid, 256, 0, st>>>(a); break;
    case 5: peer_reduce_adam_kernel<5><<<grid, 256, 0, st>>>(a); break;
    case 6: peer_reduce_adam_kernel<6><<<grid, 256, 0, st>>>(a); break;
    case 7: peer_reduce_adam_kernel<7><<<grid, 256, 0, st>>>(a); break;
    default: peer_reduce_adam_kernel<8><<<grid, 256, 0, st>>>(a); break;
  }
  count_launch();
  return check_launch("peer_reduce_adam");
}

// ------------------------------------------------------------------------------------------------
// All-reduce (sum) of a SMALL fp64 vector (the packed RunningMeanAndVar batch statistics: 2 x channels values,
// running_mean_and_var.py:28-38) through peer memory, as one tiny CTA.
// Why not NCCL for 480 bytes: its kernel spins on an SM until the slowest rank arrives, and while it sits there the
// persistent one-CTA-per-SM convolution kernels of the main stream (200 KB of shared memory each) cannot place their 148th
// CTA -- measured 0.21 ms per training step at 2 GPUs (7.37 -> 7.17 ms).  This kernel holds 256 threads and no shared memory,
// so it co-resides with anything.
//   publish : every rank stores its vector into slot [rank][seq & 1] of EVERY rank's exchange area, fences, then raises
//             flag [rank] on every rank to `seq` (release, system scope);
//   combine : waits until every flag of the local block reached `seq` (acquire), then sums the local slots in rank order
//             -- the same order on every rank, so the replicas' statistics stay bit-identical -- and overwrites `data`.
// Two slot parities: a rank can run at most one exchange ahead of the slowest reader (it needs that reader's next flag).
// ------------------------------------------------------------------------------------------------
static constexpr int kSmallMax = 1024;  // doubles per vector

struct PeerSumArgs {
  double* slots[kMaxPeers];     // every rank's exchange area: [world][2][kSmallMax] doubles
  uint32_t* flags[kMaxPeers];   // every rank's flag block: [world] sequence flags, [17] error
  double* data;
  int n, rank, world;
  uint32_t seq;
};

__global__ void __launch_bounds__(256) peer_sum_f64_kernel(const PeerSumArgs a) {
  const int par = static_cast<int>(a.seq & 1u);
  for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
    const double v = a.data[i];
    for (int p = 0; p < a.world; ++p) {
      double* dst = a.slots[p] + (static_cast<int64_t>(a.rank) * 2 + par) * kSmallMax + i;
      asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(dst), "d"(v) : "memory");
    }
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (threadIdx.x < a.world) {
    st_release_sys(a.flags[threadIdx.x] + a.rank, a.seq);
    if (!wait_flag(a.flags[a.rank] + threadIdx.x, a.seq)) {
      s_ok = 0;
      a.flags[a.rank][17] = 1;
    }
  }
  __syncthreads();
  if (!s_ok) return;
  const double* mine = a.slots[a.rank];
  for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
    double s = 0.0;
    for (int p = 0; p < a.world; ++p) {
      double v;
      asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(mine + (static_cast<int64_t>(p) * 2 + par) * kSmallMax + i) : "memory");
      s += v;
    }
    a.data[i] = s;
  }
}

}  // namespace pnvo

// ------------------------------------------------------------------------------------------------
// C ABI (include/pnvo.h, section "e multi-GPU")
// ------------------------------------------------------------------------------------------------
using namespace pnvo;

extern "C" {

int pnvo_peer_alloc(int64_t bytes, void** ptr_out, void* handle64_out) {
  PNVO_REQUIRE(bytes > 0 && ptr_out && handle64_out, "pnvo_peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, static_cast<size_t>(bytes));
  if (e == cudaSuccess) e = cudaMemset(p, 0, static_cast<size_t>(bytes));
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    set_error("pnvo_peer_alloc: %s", cudaGetErrorString(e));
    if (p) cudaFree(p);
    return -2;
  }
  memcpy(handle64_out, &h, 64);
  *ptr_out = p;
  return 0;
}

int pnvo_peer_open(const void* handle64, void** ptr_out) {
  PNVO_REQUIRE(handle64 && ptr_out, "pnvo_peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    set_error("pnvo_peer_open: %s", cudaGetErrorString(e));
    return -2;
  }
  *ptr_out = p;
  return 0;
}

int pnvo_peer_close(void* ptr) {
  const cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) {
    set_error("pnvo_peer_close: %s", cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

int pnvo_peer_free(void* ptr) {
  const cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) {
    set_error("pnvo_peer_free: %s", cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

int pnvo_peer_reduce_adam(const void* const* grads, void* const* params, void* const* flags, float* m, float* v,
                          int64_t n, int rank, int world, uint32_t seq, float lr, float beta1, float beta2, float eps,
                          int step, void* stream) {
  PNVO_REQUIRE(grads && params && flags && m && v, "pnvo_peer_reduce_adam: null argument");
  PNVO_REQUIRE(n > 0 && (n & 3) == 0, "pnvo_peer_reduce_adam: n must be a positive multiple of 4");
  PNVO_REQUIRE(world >= 2 && world <= kMaxPeers, "pnvo_peer_reduce_adam: world must be 2..8");
  PeerAdamArgs a{};
  for (int p = 0; p < world; ++p) {
    PNVO_REQUIRE(grads[p] && params[p] && flags[p], "pnvo_peer_reduce_adam: null peer pointer");
    PNVO_REQUIRE(((reinterpret_cast<uintptr_t>(grads[p]) | reinterpret_cast<uintptr_t>(params[p])) & 15) == 0,
                 "pnvo_peer_reduce_adam: buckets must be 16-byte aligned");
    a.grad[p] = static_cast<const float*>(grads[p]);
    a.param[p] = static_cast<float*>(params[p]);
    a.flags[p] = static_cast<uint32_t*>(flags[p]);
  }
  a.m = m;
  a.v = v;
  a.n4 = n / 4;
  a.slice4 = ceil_div64(a.n4, world);
  a.rank = rank;
  a.world = world;
  a.seq = seq;
  a.lr = lr;
  a.b1 = beta1;
  a.b2 = beta2;
  a.eps = eps;
  a.bc1 = 1.f - powf(beta1, static_cast<float>(step));
  a.bc2_sqrt = sqrtf(1.f - powf(beta2, static_cast<float>(step)));
  return peer_reduce_adam_launch(a, static_cast<cudaStream_t>(stream));
}

int pnvo_peer_sum_f64(void* const* slots, void* const* flags, double* data, int n, int rank, int world, uint32_t seq,
                      void* stream) {
  PNVO_REQUIRE(slots && flags && data, "pnvo_peer_sum_f64: null argument");
  PNVO_REQUIRE(n > 0 && n <= kSmallMax, "pnvo_peer_sum_f64: n must be 1..1024");
  PNVO_REQUIRE(world >= 2 && world <= kMaxPeers && rank >= 0 && rank < world, "pnvo_peer_sum_f64: bad rank / world");
  PeerSumArgs a{};
  for (int p = 0; p < world; ++p) {
    PNVO_REQUIRE(slots[p] && flags[p], "pnvo_peer_sum_f64: null peer pointer");
    a.slots[p] = static_cast<double*>(slots[p]);
    a.flags[p] = static_cast<uint32_t*>(flags[p]);
  }
  a.data = data;
  a.n = n;
  a.rank = rank;
  a.world = world;
  a.seq = seq;
  peer_sum_f64_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  count_launch();
  return check_launch("peer_sum_f64");
}

}  // extern "C"
