#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "tmap.cuh"

namespace pnvo {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

static EncodeTiledFn g_tiled = nullptr;
static EncodeIm2colFn g_im2col = nullptr;
static int g_driver = 0;
static std::mutex g_mu;
static std::map<std::vector<int64_t>, CUtensorMap> g_cache;

static int resolve() {
  if (g_tiled && g_im2col) return 0;
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return -1;
  }
  g_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
    set_error("cuTensorMapEncodeIm2col not available from the driver");
    return -1;
  }
  g_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
  cudaDriverGetVersion(&g_driver);
  return 0;
}

int tmap_im2col(CUtensorMap* out, const void* x, int N, int H, int W, int C, int R, int S, int stride, int pad,
                int channels, int pixels, int row_pitch_px, int pad_w_lo, int pad_h_hi, int pad_w_hi) {
  if (row_pitch_px <= 0) row_pitch_px = W;
  if (pad_w_lo < 0) pad_w_lo = pad;
  if (pad_h_hi < 0) pad_h_hi = pad;
  if (pad_w_hi < 0) pad_w_hi = pad_w_lo;
  std::lock_guard<std::mutex> lk(g_mu);
  if (resolve()) return -1;
  std::vector<int64_t> key = {1, reinterpret_cast<int64_t>(x), N, H, W, C, R, S, stride, pad, channels, pixels, row_pitch_px,
                              pad_w_lo, pad_h_hi, pad_w_hi};
  auto it = g_cache.find(key);
  if (it != g_cache.end()) {
    *out = it->second;
    return 0;
  }
  alignas(64) CUtensorMap m;
  const cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                              static_cast<cuuint64_t>(N)};
  const cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(row_pitch_px) * C * 2,
                                 static_cast<cuuint64_t>(H) * row_pitch_px * C * 2};
  // bounding box of the filter origin: [-pad_lo, dim + pad_hi - (filter - 1)) in W and H
  const int lower[2] = {-pad_w_lo, -pad};
  const int upper[2] = {pad_w_hi - (S - 1), pad_h_hi - (R - 1)};
  const cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(stride), static_cast<cuuint32_t>(stride), 1};
  const CUtensorMapSwizzle sw = (channels * 2 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const CUresult r = g_im2col(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x), dims, strides, lower, upper,
                              static_cast<cuuint32_t>(channels), static_cast<cuuint32_t>(pixels), estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeIm2col failed (%d) for [%d,%d,%d,%d] %dx%d/s%d p%d box %dx%d", static_cast<int>(r), N, H,
              W, C, R, S, stride, pad, pixels, channels);
    return -1;
  }
  // driver <= 13.1 workaround for tensors smaller than 128 KiB (same bit CUTLASS clears in
  // cute/atom/copy_traits_sm90_im2col.hpp)
  if (g_driver <= 13010 && static_cast<int64_t>(N) * H * W * C * 2 < 131072)
    reinterpret_cast<uint64_t*>(&m)[1] &= ~(1ull << 21);
  g_cache[key] = m;
  *out = m;
  return 0;
}

int tmap_tiled4d(CUtensorMap* out, const void* x, int N, int H, int W, int C, int box_w, int swizzle_bytes, int box_h,
                 int box_c) {
  if (box_c <= 0) box_c = C;
  std::lock_guard<std::mutex> lk(g_mu);
  if (resolve()) return -1;
  std::vector<int64_t> key = {3, reinterpret_cast<int64_t>(x), N, H, W, C, box_w, swizzle_bytes, box_h, box_c};
  auto it = g_cache.find(key);
  if (it != g_cache.end()) {
    *out = it->second;
    return 0;
  }
  alignas(64) CUtensorMap m;
  const cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                              static_cast<cuuint64_t>(N)};
  const cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W) * C * 2,
                                 static_cast<cuuint64_t>(H) * W * C * 2};
  const cuuint32_t box[4] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = g_tiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE,
                             swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(4d) failed (%d) for [%d,%d,%d,%d] box %d", static_cast<int>(r), N, H, W, C, box_w);
    return -1;
  }
  g_cache[key] = m;
  *out = m;
  return 0;
}

int tmap_tiled2d(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                 int box_cols) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (resolve()) return -1;
  std::vector<int64_t> key = {2, reinterpret_cast<int64_t>(base), rows, cols, ld, box_rows, box_cols};
  auto it = g_cache.find(key);
  if (it != g_cache.end()) {
    *out = it->second;
    return 0;
  }
  alignas(64) CUtensorMap m;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = (box_cols * 2 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const CUresult r = g_tiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for [%lld x %lld, ld %lld] box %dx%d", static_cast<int>(r),
              static_cast<long long>(rows), static_cast<long long>(cols), static_cast<long long>(ld), box_rows, box_cols);
    return -1;
  }
  g_cache[key] = m;
  *out = m;
  return 0;
}

}  // namespace pnvo
