// Argument structs + launchers of the elementwise / reduction kernels (norm_pool.cu).
#pragma once
#include "common.cuh"

namespace pnvo {

static constexpr int kMaxInC = 32;  // input channels of the first conv (30 for the default VO model)

struct AssembleArgs {
  const float* src[4];  // NHWC fp32 sources (rgb, depth, discretized_depth, top_down_view)
  int nch[4];           // channels per pixel of each source
  float pre_scale[4];   // 1/255 for rgb, 1 otherwise (vo_cnn.py:117-118)
  int n_src;
  int C, Cpad;             // real / padded channel count of the assembled tensor
  signed char src_idx[kMaxInC];  // output channel -> source tensor
  signed char src_ch[kMaxInC];   // output channel -> channel inside that source
  const float* scale;      // per output channel (nullable = 1)
  const float* shift;      // per output channel (nullable = 0)
  __half* out;             // [n_pix][Cpad], or W-padded rows when out_pitch > 0
  __half* out_lo;          // split-fp16 mode (nullable): value - fp16(value), same layout as out
  int64_t n_pix;
  int row_w, out_pitch;    // pixels per image row / pixels per padded output row (image starts at pixel 3)
};
int assemble_launch(const AssembleArgs& a, cudaStream_t st);
int input_stats_launch(const AssembleArgs& a, double* stats, cudaStream_t st);
int rmv_update_launch(const double* stats, double n_batch, double pix_per_sample, float* mean, float* var,
                      float* count, int C, int update, int have_rmv, float* scale, float* shift, cudaStream_t st);
int avgpool2_launch(const float* src, int B, int H, int W, int C, float pre_scale, __half* out, int Cpad, int coff,
                    cudaStream_t st, __half* out_lo = nullptr, float* out32 = nullptr, int ld32 = 0);
int zero_launch(void* p, int64_t bytes, cudaStream_t st);
int upsample2_launch(const __half* src, __half* dst, int B, int OH, int OW, int IH, int IW, int C, cudaStream_t st);
// raw_input.cu: PNVO_OP_RAW_STATS / PNVO_OP_RAW_ASSEMBLE (field layout documented there)
// act_embed.cu: PNVO_OP_ACT_EMBED_FWD / PNVO_OP_ACT_EMBED_BWD
int act_embed_op(int code, const int32_t* i, const float* f, void* const* p, cudaStream_t st);
int raw_op(int code, const int32_t* i, const float* f, void* const* p, cudaStream_t st);
// stem_exact.cu: PNVO_OP_STEM_EXACT_PREP / _PACK / PNVO_OP_STEM_DY_SUMS / PNVO_OP_STEM_EXACT_UNPACK
int stem_exact_op(int code, const int32_t* i, const float* f, void* const* p, cudaStream_t st);

struct GnArgs {
  const void* x;       // raw conv output [B*HW][C] fp16 (or fp32 when x_fp32)
  const __half* x_lo;  // optional residual plane of x (raw outputs of split-precision convs: x = value + residual)
  int x_fp32;
  const double* stats; // [B][G][2] fp64 (sum, sum of squares)
  const float* gamma;  // [C] (padded channels: 0)
  const float* beta;   // [C]
  const __half* res;   // optional residual [B*HW][C]
  __half* y;           // [B*HW][C] (or pooled)
  __half* y_lo;        // split-fp16 mode (nullable): residual plane y_fp32 - fp16(y_fp32), same layout as y
  const __half* res_lo; // residual plane of `res` (nullable)
  int C, C_real, G, cpg, HW;
  float cnt;           // elements per group = cpg_real * HW
  float eps;
  int relu;
};
int gn_apply_launch(const GnArgs& a, int B, cudaStream_t st);
int gn_pool_launch(const GnArgs& a, int B, int H, int W, int PH, int PW, uint8_t* argmax, cudaStream_t st);
int pool_bwd_launch(const __half* g, const __half* pooled, const uint8_t* argmax, __half* dy, int B, int H, int W,
                    int PH, int PW, int C, cudaStream_t st);

struct GnBwdArgs {
  const __half* g;         // gradient w.r.t. the GN(+ReLU) output [B*HW][C]
  const __half* relu_ref;  // saved post-ReLU output (mask = relu_ref > 0); null = no ReLU
  const void* x;           // raw conv output (GN input)
  int x_fp32;
  const double* stats;
  const float* gamma;
  float* sums;             // [B][C][2] (sum dy, sum dy*xhat), pre-zeroed for the reduce pass
  __half* dx;              // apply pass: gradient w.r.t. the raw conv output
  __half* dy_out;          // apply pass, optional: masked g (identity-branch gradient)
  int C, C_real, G, cpg, HW;
  float cnt, eps;
  float g_scale;           // multiplies the incoming gradient where the ReLU mask is set (inverted dropout)
  // apply pass, optional (exact-input stem, C == 32): border-class sums S[5][5][32] of the dx it writes, accumulated with
  // atomics -- the sums stem_exact_unpack needs (stem_exact.cu), without a separate pass over dx
  float* class_sums;
  int OH, OW;
};
int gn_bwd_reduce_launch(const GnBwdArgs& a, int B, cudaStream_t st);
int gn_bwd_apply_launch(const GnBwdArgs& a, int B, cudaStream_t st);
int gn_bwd_fused_supported(const GnBwdArgs& a);
int gn_bwd_fused_launch(const GnBwdArgs& a, int B, cudaStream_t st);
int gn_param_grad_launch(const float* sums, int B, int C, int C_real, float* dgamma, float* dbeta, int accumulate,
                         cudaStream_t st);

// descriptor tables of the batched (multi) ops; mirrored field by field in pointnav_vo_b200/lib.py
struct PackDesc {
  const float* w;
  __half* wp;
  __half* wt;
  int Cout, Cin, R, S, cin_pad, ld_p, cout_pad, ld_t, t_mode, src_ld;
};
struct UnpackDesc {
  const float* dwp;
  float* grad;
  int Cout, Cin, R, S, cin_pad, ld_p, accumulate, dst_ld;
};
struct GnParamDesc {
  const float* sums;
  float* dgamma;
  float* dbeta;
  int C, C_real;
};
static_assert(sizeof(PackDesc) == 64 && sizeof(UnpackDesc) == 48 && sizeof(GnParamDesc) == 32,
              "descriptor layouts are mirrored by ctypes structs in lib.py");
int multi_launch(int code, const void* table, int n, int B, cudaStream_t st);
int pack_w_launch(const float* w, int Cout, int Cin, int R, int S, __half* wp, int cin_pad, int ld_p, __half* wt,
                  int cout_pad, int ld_t, int t_mode, cudaStream_t st, int src_ld = 0);
int unpack_dw_launch(const float* dwp, int Cout, int Cin, int R, int S, int cin_pad, int ld_p, float* grad,
                     int accumulate, cudaStream_t st, int dst_ld = 0);
int bias_relu_launch(const float* z, const float* bias, int B, int N, int relu, float* h32, __half* h16,
                     cudaStream_t st);
int bias_relu_bwd_launch(const float* dh, const float* h, int B, int N, __half* dz16, float* db, int accumulate,
                         cudaStream_t st);
int head_fwd_launch(const float* h, const float* W, const float* bias, int B, int K, int O, float* out,
                    cudaStream_t st);
int head_bwd_launch(const float* dout, const float* h, const float* W, int B, int K, int O, float* dW, float* db2,
                    __half* dz16, float* db1, int accumulate, float dh_scale, cudaStream_t st);
int dropout_launch(void* buf, int64_t n, int is_fp16, uint64_t* seed, int site, float p, int advance, cudaStream_t st);
int mse_loss_launch(const float* pred, const float* tgt, const float* dz_mask, const int64_t* data_types, int B, int O,
                    float w0, float w1, float w2, float grad_scale, float* dout, float* loss, cudaStream_t st);
int geo_inv_loss_launch(const float* pred, const int64_t* actions, const int64_t* data_types, int B, int O,
                        int move_forward, int turn_left, int turn_right, float weight, float grad_scale, float* dout,
                        float* loss, int* err, cudaStream_t st);
int adam_launch(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps,
                int step, float grad_scale, cudaStream_t st);

}  // namespace pnvo
