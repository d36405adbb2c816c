// C ABI of libpnvo.so (include/pnvo.h): error plumbing and the op-program interpreter.
// Field layout of every op (i = int slots, f = float slots, p = pointer slots) is mirrored by
// pointnav_vo_b200/lib.py, which is the only producer of pnvo_op records.
#include <cstdarg>
#include <cstdio>
#include <atomic>

#include "common.cuh"
#include "elem.cuh"
#include "ops.cuh"

namespace pnvo {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

static int run_op(const pnvo_op& op, cudaStream_t st) {
  const int32_t* i = op.i;
  const float* f = op.f;
  void* const* p = op.p;
  const int code = op.code & 0xffff;  // bits 16+: lane (PNVO_OP_SIDE_LANE)
  switch (code) {
    case PNVO_OP_JOIN:
      return 0;  // only meaningful to the graph capture
    case PNVO_OP_ZERO:
      // p0 = buffer; i0|i1 = byte count (lo, hi)
      return zero_launch(p[0], (static_cast<int64_t>(static_cast<uint32_t>(i[1])) << 32) | static_cast<uint32_t>(i[0]), st);
    case PNVO_OP_ASSEMBLE:
    case PNVO_OP_INPUT_STATS: {
      // p0..p3 = sources, p4 = scale, p5 = shift, p6 = out (assemble) / fp64 stats (input_stats)
      // i0 = n_src, i1..i4 = nch, i5 = C, i6 = Cpad, i7|i8 = n_pix, i9..i16 = packed LUT (4 x int8 per word:
      // 8 words of src_idx<<4|... see below), f0..f3 = pre_scale
      AssembleArgs a{};
      a.n_src = i[0];
      for (int t = 0; t < 4; ++t) {
        a.src[t] = static_cast<const float*>(p[t]);
        a.nch[t] = i[1 + t];
        a.pre_scale[t] = f[t];
      }
      a.C = i[5];
      a.Cpad = i[6];
      a.row_w = i[25];
      a.out_pitch = i[26];
      a.n_pix = (static_cast<int64_t>(static_cast<uint32_t>(i[8])) << 32) | static_cast<uint32_t>(i[7]);
      PNVO_REQUIRE(a.C >= 1 && a.C <= kMaxInC, "assemble: C=%d", a.C);
      for (int c = 0; c < kMaxInC; ++c) {
        // word 9 + c/2: two channels per word, each 16 bits = (src_idx << 8) | src_ch
        const uint32_t w = static_cast<uint32_t>(i[9 + c / 2]);
        const uint32_t h = (c & 1) ? (w >> 16) : (w & 0xffff);
        a.src_idx[c] = static_cast<signed char>(h >> 8);
        a.src_ch[c] = static_cast<signed char>(h & 0xff);
        if (c < a.C) PNVO_REQUIRE(a.src_idx[c] >= 0 && a.src_idx[c] < a.n_src && a.src_ch[c] < a.nch[a.src_idx[c]],
                                  "assemble: bad LUT entry for channel %d", c);
      }
      a.scale = static_cast<const float*>(p[4]);
      a.shift = static_cast<const float*>(p[5]);
      if (code == PNVO_OP_ASSEMBLE) {
        a.out = static_cast<__half*>(p[6]);
        a.out_lo = static_cast<__half*>(p[7]);  // split-fp16 residual plane (nullable)
        return assemble_launch(a, st);
      }
      return input_stats_launch(a, static_cast<double*>(p[6]), st);
    }
    case PNVO_OP_PACK_W_MULTI:
    case PNVO_OP_UNPACK_DW_MULTI:
    case PNVO_OP_GN_PARAM_GRAD_MULTI:
      // p0 = device table of PackDesc / UnpackDesc / GnParamDesc (elem.cuh); i0 = entries, i1 = B (param grad)
      return multi_launch(code, p[0], i[0], i[1], st);
    case PNVO_OP_GEO_INV_LOSS:
      // p0 = pred [B][O], p1 = actions int64 [B], p2 = dout (nullable, accumulated), p3 = loss[3] (total+=, rot, pos),
      // p4 = data types int64 [B] (nullable: every row, interleaved pairs), p5 = int32 error flag (nullable)
      // i0 = B, i1 = O, i2 = MOVE_FORWARD id, i3 = TURN_LEFT id, i4 = TURN_RIGHT id; f0 = loss_inv_weight, f1 = gradient scale
      return geo_inv_loss_launch(static_cast<const float*>(p[0]), static_cast<const int64_t*>(p[1]),
                                 static_cast<const int64_t*>(p[4]), i[0], i[1], i[2], i[3], i[4], f[0], f[1],
                                 static_cast<float*>(p[2]), static_cast<float*>(p[3]), static_cast<int*>(p[5]), st);
    case PNVO_OP_UPSAMPLE2:
      // p0 = src [B,OH,OW,C] fp16, p1 = dst [B,IH,IW,C] fp16; i0 = B, i1 = OH, i2 = OW, i3 = IH, i4 = IW, i5 = C
      return upsample2_launch(static_cast<const __half*>(p[0]), static_cast<__half*>(p[1]), i[0], i[1], i[2], i[3], i[4],
                              i[5], st);
    case PNVO_OP_ACT_EMBED_FWD:
    case PNVO_OP_ACT_EMBED_BWD:
      return act_embed_op(code, i, f, p, st);
    case PNVO_OP_RAW_STATS:
    case PNVO_OP_RAW_ASSEMBLE:
      return raw_op(code, i, f, p, st);
    case PNVO_OP_STEM_EXACT_PREP:
    case PNVO_OP_STEM_EXACT_PACK:
    case PNVO_OP_STEM_DY_SUMS:
    case PNVO_OP_STEM_EXACT_UNPACK:
      return stem_exact_op(code, i, f, p, st);
    case PNVO_OP_RMV_UPDATE:
      // p0 = fp64 stats, p1 = _mean, p2 = _var, p3 = _count, p4 = scale, p5 = shift
      // i0 = C, i1 = update, i2 = have_rmv; f0 = batch samples (all ranks), f1 = pixels per sample
      return rmv_update_launch(static_cast<const double*>(p[0]), f[0], f[1], static_cast<float*>(p[1]),
                               static_cast<float*>(p[2]), static_cast<float*>(p[3]), i[0], i[1], i[2],
                               static_cast<float*>(p[4]), static_cast<float*>(p[5]), st);
    case PNVO_OP_CONV: {
      // p0 = x, p1 = w, p2 = y, p3 = add, p4 = stats, p5 = x_lo, p6 = w_lo (split-fp16 mode), p7 = y_lo (output as value +
      // residual fp16 planes)
      ConvArgs a{};
      a.x_lo = static_cast<const __half*>(p[5]);
      a.w_lo = static_cast<const __half*>(p[6]);
      a.y_lo = static_cast<__half*>(p[7]);
      a.x_c = static_cast<const float*>(p[8]);   // compact fp32 image of a one-channel stem (nullable)
      a.x = static_cast<const __half*>(p[0]);
      a.w = static_cast<const __half*>(p[1]);
      a.y = p[2];
      a.add = static_cast<const __half*>(p[3]);
      a.stats = static_cast<double*>(p[4]);
      a.B = i[0]; a.IH = i[1]; a.IW = i[2]; a.Cin = i[3]; a.OH = i[4]; a.OW = i[5]; a.R = i[6]; a.S = i[7];
      a.mul = i[8]; a.pad = i[9]; a.div = i[10]; a.w_ld = i[11]; a.n_total = i[12]; a.n_store = i[13]; a.ldo = i[14];
      a.cpg = i[15]; a.G = i[16]; a.out_fp32 = i[17]; a.pad_w = i[18]; a.force_generic = i[19]; a.cin_real = i[21];
      // i22 = output pixel stride (0 = dense), i23 = output offsets (h | w << 8), i24 / i25 = full output height / width,
      // i26 = 0 (symmetric padding) or (pad_hi_h + 1) | (pad_hi_w + 1) << 8
      a.o_mul = i[22]; a.o_off_h = i[23] & 0xff; a.o_off_w = (i[23] >> 8) & 0xff; a.o_H = i[24]; a.o_W = i[25];
      a.asym = i[26] != 0; a.pad_hi_h = (i[26] & 0xff) - 1; a.pad_hi_w = ((i[26] >> 8) & 0xff) - 1;
      return conv_launch(a, st);
    }
    case PNVO_OP_WGRAD: {
      // p0 = x, p1 = dy, p2 = dw (packed fp32)
      WgradArgs a{};
      a.x = static_cast<const __half*>(p[0]);
      a.dy = static_cast<const __half*>(p[1]);
      a.dw = static_cast<float*>(p[2]);
      a.x_c = static_cast<const float*>(p[3]);
      a.B = i[0]; a.IH = i[1]; a.IW = i[2]; a.Cin = i[3]; a.OH = i[4]; a.OW = i[5]; a.R = i[6]; a.S = i[7];
      a.mul = i[8]; a.pad = i[9]; a.w_ld = i[11]; a.n_total = i[12]; a.ld_dy = i[14]; a.pad_w = i[18]; a.force_generic = i[19]; a.x_row_pitch = i[20];
      a.cin_real = i[21];
      return wgrad_launch(a, st);
    }
    case PNVO_OP_GN_APPLY:
    case PNVO_OP_GN_POOL: {
      // p0 = x, p1 = stats, p2 = gamma, p3 = beta, p4 = res, p5 = y, p6 = argmax
      // i0 = B, i1 = C, i2 = G, i3 = cpg, i4 = HW, i5 = relu, i6 = x_fp32, i7..i10 = H, W, PH, PW; f0 = cnt, f1 = eps
      GnArgs a{};
      a.x = p[0]; a.stats = static_cast<const double*>(p[1]); a.gamma = static_cast<const float*>(p[2]);
      a.beta = static_cast<const float*>(p[3]); a.res = static_cast<const __half*>(p[4]);
      a.y = static_cast<__half*>(p[5]);
      a.y_lo = static_cast<__half*>(p[7]); a.res_lo = static_cast<const __half*>(p[8]);  // split-fp16 planes (nullable)
      a.x_lo = static_cast<const __half*>(p[9]);  // residual plane of a raw conv output stored as value + residual
      a.C = i[1]; a.G = i[2]; a.cpg = i[3]; a.HW = i[4]; a.relu = i[5]; a.x_fp32 = i[6]; a.C_real = i[11];
      a.cnt = f[0]; a.eps = f[1];
      if (code == PNVO_OP_GN_APPLY) return gn_apply_launch(a, i[0], st);
      return gn_pool_launch(a, i[0], i[7], i[8], i[9], i[10], static_cast<uint8_t*>(p[6]), st);
    }
    case PNVO_OP_GN_POOL_BWD:
      // p0 = g (pooled grad), p1 = pooled output, p2 = argmax, p3 = dy out; i0 = B, i1 = C, i7..i10 = H, W, PH, PW
      return pool_bwd_launch(static_cast<const __half*>(p[0]), static_cast<const __half*>(p[1]),
                             static_cast<const uint8_t*>(p[2]), static_cast<__half*>(p[3]), i[0], i[7], i[8], i[9],
                             i[10], i[1], st);
    case PNVO_OP_GN_BWD_REDUCE:
    case PNVO_OP_GN_BWD_FUSED:
    case PNVO_OP_GN_BWD_APPLY: {
      // p0 = g, p1 = relu_ref, p2 = x, p3 = stats, p4 = gamma, p5 = sums, p6 = dx, p7 = dy_out
      GnBwdArgs a{};
      a.g = static_cast<const __half*>(p[0]); a.relu_ref = static_cast<const __half*>(p[1]); a.x = p[2];
      a.stats = static_cast<const double*>(p[3]); a.gamma = static_cast<const float*>(p[4]);
      a.sums = static_cast<float*>(p[5]); a.dx = static_cast<__half*>(p[6]); a.dy_out = static_cast<__half*>(p[7]);
      a.C = i[1]; a.G = i[2]; a.cpg = i[3]; a.HW = i[4]; a.x_fp32 = i[6]; a.C_real = i[11];
      a.cnt = f[0]; a.eps = f[1]; a.g_scale = (f[2] == 0.f) ? 1.f : f[2];
      a.class_sums = static_cast<float*>(p[8]); a.OH = i[7]; a.OW = i[8];   // apply pass only (exact-input stem)
      if (code == PNVO_OP_GN_BWD_REDUCE) return gn_bwd_reduce_launch(a, i[0], st);
      if (code == PNVO_OP_GN_BWD_FUSED) return gn_bwd_fused_launch(a, i[0], st);
      return gn_bwd_apply_launch(a, i[0], st);
    }
    case PNVO_OP_GN_PARAM_GRAD:
      // p0 = sums, p1 = dgamma, p2 = dbeta; i0 = B, i1 = C, i2 = C_real, i3 = accumulate
      return gn_param_grad_launch(static_cast<const float*>(p[0]), i[0], i[1], i[2], static_cast<float*>(p[1]),
                                  static_cast<float*>(p[2]), i[3], st);
    case PNVO_OP_PACK_W:
      // p0 = w (OIHW fp32), p1 = wp, p2 = wt; i0..i3 = Cout, Cin, R, S; i4 = cin_pad, i5 = ld_p, i6 = cout_pad, i7 = ld_t,
      // i8 = t_mode (0: conv dgrad layout, 1: plain transpose of wp)
      return pack_w_launch(static_cast<const float*>(p[0]), i[0], i[1], i[2], i[3], static_cast<__half*>(p[1]), i[4],
                           i[5], static_cast<__half*>(p[2]), i[6], i[7], i[8], st, i[9]);
    case PNVO_OP_UNPACK_DW:
      // p0 = dwp, p1 = grad; i0..i3 = Cout, Cin, R, S; i4 = cin_pad, i5 = ld_p, i6 = accumulate
      return unpack_dw_launch(static_cast<const float*>(p[0]), i[0], i[1], i[2], i[3], i[4], i[5],
                              static_cast<float*>(p[1]), i[6], st, i[7]);
    case PNVO_OP_BIAS_RELU:
      // p0 = z, p1 = bias, p2 = h32, p3 = h16; i0 = B, i1 = N, i2 = relu
      return bias_relu_launch(static_cast<const float*>(p[0]), static_cast<const float*>(p[1]), i[0], i[1], i[2],
                              static_cast<float*>(p[2]), static_cast<__half*>(p[3]), st);
    case PNVO_OP_BIAS_RELU_BWD:
      // p0 = dh, p1 = h, p2 = dz16, p3 = db; i0 = B, i1 = N, i2 = accumulate
      return bias_relu_bwd_launch(static_cast<const float*>(p[0]), static_cast<const float*>(p[1]), i[0], i[1],
                                  static_cast<__half*>(p[2]), static_cast<float*>(p[3]), i[2], st);
    case PNVO_OP_HEAD_FWD:
      // p0 = h, p1 = W, p2 = bias, p3 = out; i0 = B, i1 = K, i2 = O
      return head_fwd_launch(static_cast<const float*>(p[0]), static_cast<const float*>(p[1]),
                             static_cast<const float*>(p[2]), i[0], i[1], i[2], static_cast<float*>(p[3]), st);
    case PNVO_OP_HEAD_BWD:
      // p0 = dout, p1 = h, p2 = W, p3 = dW, p4 = db2, p5 = dz16, p6 = db1; i0 = B, i1 = K, i2 = O, i3 = accumulate
      return head_bwd_launch(static_cast<const float*>(p[0]), static_cast<const float*>(p[1]),
                             static_cast<const float*>(p[2]), i[0], i[1], i[2], static_cast<float*>(p[3]),
                             static_cast<float*>(p[4]), static_cast<__half*>(p[5]), static_cast<float*>(p[6]), i[3],
                             (f[0] == 0.f) ? 1.f : f[0], st);
    case PNVO_OP_MSE_LOSS:
      // p0 = pred, p1 = target, p2 = dz mask (nullable), p3 = dout (nullable), p4 = loss, p5 = data types int64 [B]
      // (nullable; with them: one mean per data type, summed); i0 = B, i1 = O; f0..f2 = loss weights, f3 = gradient scale
      return mse_loss_launch(static_cast<const float*>(p[0]), static_cast<const float*>(p[1]),
                             static_cast<const float*>(p[2]), static_cast<const int64_t*>(p[5]), i[0], i[1], f[0], f[1],
                             f[2], f[3], static_cast<float*>(p[3]), static_cast<float*>(p[4]), st);
    case PNVO_OP_CONV_STEM:
      // p0 = x [B,IH,IW,32] fp16, p1 = stem-packed weights, p2 = y, p3 = stats; i0 = B, i1 = IH, i2 = IW, i3 = G, i4 = cpg, i5 = stages
      return conv_stem_fwd_launch(static_cast<const __half*>(p[0]), static_cast<const __half*>(p[1]), p[2],
                                  static_cast<double*>(p[3]), i[0], i[1], i[2], i[3], i[4], i[5], st);
    case PNVO_OP_CONV_STEM2:
      // p0 = W-padded x, p1 = stem2-packed weights, p2 = y, p3 = stats, p4 = x_lo (split mode), p5 = fp16 tensor added in
      // the epilogue, p6 = border-class bias table [5][5][32] fp32 (exact-input stem); i0 = B, i1 = IH, i2 = IW, i3 = G,
      // i4 = cpg, i5 = fp32 output
      return conv_stem2_fwd_launch(static_cast<const __half*>(p[0]), static_cast<const __half*>(p[1]), p[2],
                                   static_cast<double*>(p[3]), i[0], i[1], i[2], i[3], i[4], st,
                                   static_cast<const __half*>(p[4]), static_cast<const __half*>(p[5]), i[5],
                                   static_cast<const float*>(p[6]), static_cast<__half*>(p[7]));
    case PNVO_OP_WGRAD_STEM2:
      // p0 = W-padded x, p1 = dy [B,OH,OW,32], p2 = packed fp32 dW; i0 = B, i1 = IH, i2 = IW, i3 = w_ld
      return conv_stem_wgrad2_launch(static_cast<const __half*>(p[0]), static_cast<const __half*>(p[1]),
                                     static_cast<float*>(p[2]), i[3], i[0], i[1], i[2], st);
    case PNVO_OP_PACK_W_STEM2:
      // p0 = w OIHW fp32, p1 = packed; i0 = Cin, i1 = residual plane (w - fp16(w)) instead of the value plane
      return pack_w_stem2_launch(static_cast<const float*>(p[0]), i[0], static_cast<__half*>(p[1]), i[1], st);
    case PNVO_OP_WGRAD_STEM:
      // p0 = W-padded x, p1 = dy [B,OH,OW,32], p2 = packed fp32 dW; i0 = B, i1 = IH, i2 = IW, i3 = w_ld, i4 = rows per CTA
      return conv_stem_wgrad_launch(static_cast<const __half*>(p[0]), static_cast<const __half*>(p[1]),
                                    static_cast<float*>(p[2]), i[3], i[0], i[1], i[2], i[4], st);
    case PNVO_OP_PACK_W_STEM:
      // p0 = w OIHW fp32 [32][Cin][7][7], p1 = packed; i0 = Cin
      return pack_w_stem_launch(static_cast<const float*>(p[0]), i[0], static_cast<__half*>(p[1]), st);
    case PNVO_OP_DROPOUT:
      // p0 = buffer (in place), p1 = uint64 seed on device; i0|i1 = n, i2 = is_fp16, i3 = site, i4 = advance seed; f0 = p
      return dropout_launch(p[0], (static_cast<int64_t>(static_cast<uint32_t>(i[1])) << 32) | static_cast<uint32_t>(i[0]), i[2],
                            static_cast<uint64_t*>(p[1]), i[3], f[0], i[4], st);
    case PNVO_OP_ADAM:
      // p0 = param, p1 = grad, p2 = m, p3 = v; i0|i1 = n, i2 = step; f0 = lr, f1 = beta1, f2 = beta2, f3 = eps
      return adam_launch(static_cast<float*>(p[0]), static_cast<const float*>(p[1]), static_cast<float*>(p[2]),
                         static_cast<float*>(p[3]),
                         (static_cast<int64_t>(static_cast<uint32_t>(i[1])) << 32) | static_cast<uint32_t>(i[0]), f[0],
                         f[1], f[2], f[3], i[2], 1.0f, st);
    case PNVO_OP_AVGPOOL2:
      // p0 = src fp32 NHWC, p1 = out fp16 (nullable), p2 = out_lo (split-fp16 residual plane, nullable), p3 = fp32 pooled
      // output [.., ld32] (nullable); i0 = B, i1 = H, i2 = W, i3 = C, i4 = Cpad, i5 = coff, i6 = ld32; f0 = pre_scale
      return avgpool2_launch(static_cast<const float*>(p[0]), i[0], i[1], i[2], i[3], f[0], static_cast<__half*>(p[1]),
                             i[4], i[5], st, static_cast<__half*>(p[2]), static_cast<float*>(p[3]), i[6]);
    default:
      set_error("run_ops: unknown opcode %d", code);
      return -3;
  }
}

}  // namespace pnvo

using namespace pnvo;

extern "C" const char* pnvo_last_error(void) { return g_err; }
extern "C" int pnvo_abi_version(void) { return PNVO_ABI_VERSION; }
extern "C" int64_t pnvo_launch_count(void) { return g_launches.load(); }
extern "C" int pnvo_stem_padded_width(int IW) { return stem_padded_width(IW); }
extern "C" int pnvo_conv_stem2_supported(int IH, int IW) { return conv_stem2_supported(IH, IW); }
extern "C" int pnvo_conv_stem_wgrad2_supported(int IH, int IW) { return conv_stem_wgrad2_supported(IH, IW); }
extern "C" int pnvo_gn_bwd_fused_supported(int C, int HW, int x_fp32) {
  GnBwdArgs a{};
  a.C = C; a.HW = HW; a.x_fp32 = x_fp32;
  return gn_bwd_fused_supported(a);
}

extern "C" int pnvo_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no CUDA device: %s", cudaGetErrorString(e));
    return -1;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    return -1;
  }
  if (prop.major != 10) {
    set_error("libpnvo is built for sm_100a (B200); device %d is sm_%d%d", dev, prop.major, prop.minor);
    return -2;
  }
  return 0;
}

extern "C" int pnvo_run_ops(const pnvo_op* ops, int n_ops, void* stream) {
  PNVO_REQUIRE(ops || n_ops == 0, "run_ops: null program");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int k = 0; k < n_ops; ++k) {
    const int rc = run_op(ops[k], st);
    if (rc != 0) {
      char tmp[400];
      snprintf(tmp, sizeof(tmp), "%s", g_err);
      set_error("op %d (code %d): %s", k, ops[k].code, tmp);
      return rc;
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// CUDA-graph replay of an op program: the ~215 launches of a training step are mostly 3-60 us kernels, so the
// host-side launch cost (tensor-map lookup, argument marshalling, ~4 us per cudaLaunchKernel) leaves the GPU idle
// between them.  A program whose buffers are fixed (the plan owns them) is captured once and replayed with one
// cudaGraphLaunch.
// ---------------------------------------------------------------------------------------------------------
struct PnvoGraph {
  cudaGraphExec_t exec;
  int n_kernels;
};

extern "C" int pnvo_graph_capture(const pnvo_op* ops, int n_ops, void** handle_out) {
  PNVO_REQUIRE(ops && n_ops > 0 && handle_out, "graph_capture: bad arguments");
  cudaStream_t s;
  // Kernel nodes inherit the priority of the stream they were captured on.  Captured programs (weight pack, forward,
  // backward) get the GREATEST priority: when the input pipeline of the next batch runs on a side stream (default = lowest
  // priority), free SMs go to the training step's kernels first and the input pipeline fills the SMs they leave idle
  // (tails of the persistent kernels, small grids), instead of holding SMs the one-CTA-per-SM kernels are waiting for.
  // PNVO_GRAPH_PRIORITY=0 captures at the default priority (A/B measurements).
  int prio_least = 0, prio_greatest = 0;
  cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
  const char* pe = getenv("PNVO_GRAPH_PRIORITY");
  const int prio = (pe && atoi(pe) == 0) ? prio_least : prio_greatest;
  cudaError_t e = cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, prio);
  PNVO_REQUIRE(e == cudaSuccess, "graph_capture: cudaStreamCreate: %s", cudaGetErrorString(e));
  e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) {
    cudaStreamDestroy(s);
    set_error("graph_capture: cudaStreamBeginCapture: %s", cudaGetErrorString(e));
    return -2;
  }
  const int64_t before = g_launches.load();
  int rc = 0;
  // side lane: a forked branch of the graph.  Every side op depends on all main-lane ops issued before it and on the
  // previous side op; the main lane picks the branch up again at PNVO_OP_JOIN / the end of the program.
  cudaStream_t s2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool side_open = false;
  auto join = [&]() {
    if (!side_open) return;
    cudaEventRecord(ev_join, s2);
    cudaStreamWaitEvent(s, ev_join, 0);
    side_open = false;
  };
  for (int k = 0; k < n_ops && rc == 0; ++k) {
    const int code = ops[k].code & 0xffff;
    if (code == PNVO_OP_JOIN) {
      join();
      continue;
    }
    if (ops[k].code & PNVO_OP_SIDE_LANE) {
      if (!s2) {
        cudaStreamCreateWithPriority(&s2, cudaStreamNonBlocking, prio);
        cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming);
      }
      cudaEventRecord(ev_fork, s);
      cudaStreamWaitEvent(s2, ev_fork, 0);
      side_open = true;
      rc = run_op(ops[k], s2);
    } else {
      rc = run_op(ops[k], s);
    }
  }
  join();
  const int n_kernels = static_cast<int>(g_launches.load() - before);
  g_launches.fetch_sub(n_kernels);  // nothing ran yet: launches are counted per replay
  cudaGraph_t graph = nullptr;
  e = cudaStreamEndCapture(s, &graph);
  cudaStreamDestroy(s);
  if (s2) {
    cudaStreamDestroy(s2);
    cudaEventDestroy(ev_fork);
    cudaEventDestroy(ev_join);
  }
  if (rc != 0) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  PNVO_REQUIRE(e == cudaSuccess && graph, "graph_capture: cudaStreamEndCapture: %s", cudaGetErrorString(e));
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  PNVO_REQUIRE(e == cudaSuccess, "graph_capture: cudaGraphInstantiate: %s", cudaGetErrorString(e));
  *handle_out = new PnvoGraph{exec, n_kernels};
  return 0;
}

extern "C" int pnvo_graph_launch(void* handle, void* stream) {
  PNVO_REQUIRE(handle, "graph_launch: null handle");
  PnvoGraph* g = static_cast<PnvoGraph*>(handle);
  const cudaError_t e = cudaGraphLaunch(g->exec, static_cast<cudaStream_t>(stream));
  PNVO_REQUIRE(e == cudaSuccess, "graph_launch: %s", cudaGetErrorString(e));
  count_launch(g->n_kernels);
  return 0;
}

extern "C" int pnvo_graph_destroy(void* handle) {
  if (!handle) return 0;
  PnvoGraph* g = static_cast<PnvoGraph*>(handle);
  cudaGraphExecDestroy(g->exec);
  delete g;
  return 0;
}

extern "C" int pnvo_conv_launch_info(const pnvo_op* op, int32_t* grid_x, int32_t* grid_y, int32_t* smem_bytes,
                                     int32_t* tmem_cols, int32_t* stages) {
  const int code = op ? (op->code & 0xffff) : 0;
  PNVO_REQUIRE(code == PNVO_OP_CONV || code == PNVO_OP_WGRAD, "conv_launch_info: not a conv op");
  const int32_t* i = op->i;
  if (code == PNVO_OP_CONV) {
    ConvArgs a{};
    a.B = i[0]; a.IH = i[1]; a.IW = i[2]; a.Cin = i[3]; a.OH = i[4]; a.OW = i[5]; a.R = i[6]; a.S = i[7];
    a.mul = i[8]; a.pad = i[9]; a.div = i[10]; a.w_ld = i[11]; a.n_total = i[12]; a.n_store = i[13]; a.ldo = i[14];
    a.cpg = i[15]; a.G = i[16]; a.out_fp32 = i[17]; a.pad_w = i[18];
    a.stats = static_cast<double*>(op->p[4]);
    if (conv_plan(a)) return -1;
    *grid_x = a.grid_x; *grid_y = a.grid_y; *smem_bytes = a.smem_bytes; *tmem_cols = a.tmem_cols; *stages = a.stages;
  } else {
    WgradArgs a{};
    a.B = i[0]; a.IH = i[1]; a.IW = i[2]; a.Cin = i[3]; a.OH = i[4]; a.OW = i[5]; a.R = i[6]; a.S = i[7];
    a.mul = i[8]; a.pad = i[9]; a.w_ld = i[11]; a.n_total = i[12]; a.ld_dy = i[14]; a.pad_w = i[18];
    if (wgrad_plan(a)) return -1;
    *grid_x = a.grid_x * a.grid_y; *grid_y = a.grid_z; *smem_bytes = a.smem_bytes; *tmem_cols = a.tmem_cols;
    *stages = a.stages;
  }
  return 0;
}
