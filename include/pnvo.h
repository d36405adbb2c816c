/*
 * pnvo.h -- C ABI of libpnvo.so: the B200 (sm_100a) kernels behind PointNav-VO's data-parallel hot path.
 *
 * The reference (Xiaoming-Zhao/PointNav-VO) is 100% Python/PyTorch and has no native/FFI boundary
 * (SURVEY.md section 8b); these entry points are what a ctypes binding of its hot path binds.  Each
 * one cites the reference code it replaces.  Conventions:
 *   - plain pointers and sizes only; every buffer is caller-allocated DEVICE memory (the library
 *     keeps no global state besides the last-error string and owns no memory, except the
 *     peer-mappable regions handed out by pnvo_peer_alloc for the multi-GPU exchange);
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - return value 0 = launched; negative = error, text via pnvo_last_error();
 *   - kernels are asynchronous with respect to the host.
 */
#ifndef PNVO_H_
#define PNVO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNVO_ABI_VERSION 1

const char* pnvo_last_error(void);
int pnvo_abi_version(void);
/* compute capability check: returns 0 when the current device is sm_100 (B200), <0 otherwise. */
int pnvo_check_device(void);

/* ---------------------------------------------------------------------------------------------
 * a7  depth discretisation -- BaseRLTrainerWithVO._discretize_depth_func
 *     (pointnav_vo/rl/common/base_trainer_with_vo.py:135-167; NumPy twin
 *     pointnav_vo/vo/dataset/regression_iter_dataset.py:32-69).
 * depth: n_pix fp32 in [0,1].  edges: n_channels+1 fp32 bin edges on the device (float(i/n), last 1.0).
 * onehot (nullable): fp32, element (p, c) written at onehot[p*onehot_stride + c], c < n_channels.
 * index (nullable): uint8 bin per pixel (255 = outside [0,1], which the reference asserts against).
 * err_count (nullable): int32 device counter incremented per out-of-range pixel.
 */
int pnvo_discretize_depth(const float* depth, int64_t n_pix, const float* edges, int n_channels,
                          float* onehot, int64_t onehot_stride, uint8_t* index, int32_t* err_count,
                          void* stream);

/* ---------------------------------------------------------------------------------------------
 * a8  egocentric top-down projection -- NormalizedDepth2TopDownViewHabitatTorch.gen_top_down_view
 *     (pointnav_vo/utils/geometry_utils.py:491-721), fp32 variant, batched over frames.
 * depth: [n_frames, H, W] fp32 (frame stride in_stride floats).  out: [n_frames, H, W] fp32 written at
 * out[f*out_frame_stride + (r*W + c)*out_pix_stride].  count (nullable): int32 [n_frames, H, W] raw
 * per-cell point counts (the bit-exact index map).  ray: W fp32 = (Kinv @ [u+.5, ., 1])[0].
 * consts: {min_x, x_den, z_den, depth_scale, depth_off} as produced by the host mirror of
 * geometry_utils.py:558-583,678-682.
 */
typedef struct {
  float min_x, x_den, z_den, depth_scale, depth_off;
  int rows_around_center; /* 50 */
  int center_crop;        /* 1 */
} pnvo_topdown_consts;

int pnvo_topdown_project(const float* depth, int64_t in_stride, int n_frames, int H, int W,
                         const float* ray, const pnvo_topdown_consts* consts, float* out,
                         int64_t out_frame_stride, int64_t out_pix_stride, int32_t* count, void* stream);

/* Same, reading frame f, pixel i at depth[f*in_frame_stride + i*in_pix_stride]: projects both frames of a
 * [B, H, W, 2] depth-pair tensor in place (n_frames = 2B, in_frame_stride = ... see the host mirror). */
int pnvo_topdown_project_strided(const float* depth, int64_t in_frame_stride, int64_t in_pix_stride, int n_frames,
                                 int H, int W, const float* ray, const pnvo_topdown_consts* consts, float* out,
                                 int64_t out_frame_stride, int64_t out_pix_stride, int32_t* count, void* stream);

/* Same with fp16 depth (the type the reference's HDF5 datasets store, vo/dataset/regression_iter_dataset.py:47-58 and
 * regression_geo_invariance_iter_dataset.py:229-236): widened exactly to fp32 on load, so the result equals
 * pnvo_topdown_project_strided on the widened tensor.  Strides are in elements. */
int pnvo_topdown_project_strided_f16(const uint16_t* depth, int64_t in_frame_stride, int64_t in_pix_stride, int n_frames,
                                     int H, int W, const float* ray, const pnvo_topdown_consts* consts, float* out,
                                     int64_t out_frame_stride, int64_t out_pix_stride, int32_t* count, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a13 GAE / discounted returns -- RolloutStorage.compute_returns
 *     (pointnav_vo/rl/common/rollout_storage.py:102-120).
 * rewards [T,N], value_preds [T+1,N] (row T is overwritten with next_value when use_gae, as the reference
 * does), masks [T+1,N], next_value [N], returns [T+1,N].  mode 0: one lane per env, sequential in t,
 * same rounding order as the reference (bit-exact); mode 1: warp-scan over t (affine-map composition,
 * equal within fp32 rounding).
 */
int pnvo_gae_scan(const float* rewards, float* value_preds, const float* masks, const float* next_value,
                  float* returns, int T, int N, int use_gae, float gamma, float gamma_tau, int mode,
                  void* stream);

/* ---------------------------------------------------------------------------------------------
 * a14 / 8f-3  PPO clipped-surrogate + clipped-value loss with its gradient, one pass -- PPO.update
 *     (pointnav_vo/rl/ppo/ppo.py:101-126).  All arrays [n] fp32 on the device (n = T * N' rows of a minibatch).
 * losses[0] = value_loss, losses[1] = action_loss (written; no pre-zeroing needed);
 * d_values = value_loss_coef * d(value_loss)/d(values), d_log_probs = d(action_loss)/d(log_probs): the gradient of
 * `value_loss * value_loss_coef + action_loss`; the entropy term of the total loss stays with the caller.
 */
int pnvo_ppo_loss(const float* values, const float* log_probs, const float* value_preds, const float* returns,
                  const float* old_log_probs, const float* adv_targ, int64_t n, float clip_param,
                  int use_clipped_value_loss, float value_loss_coef, float* losses, float* d_values,
                  float* d_log_probs, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a10 batched goal update -- compute_goal_pos (pointnav_vo/utils/geometry_utils.py:115-144), fp64.
 * goal_xyz [n,3] f64 in/out (cartesian), delta [n,3] f32 (dx, dz, dyaw), polar [n,2] f32 out (rho, -phi).
 */
int pnvo_goal_update(double* goal_xyz, const float* delta, float* polar, int n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The CNN path (a1-a6, a11): a flat program of ops executed in order on one stream.  The host mirror
 * of the reference modules (pointnav_vo_b200/vo/models, rl/policies) builds the program once per
 * (architecture, batch) and replays it; per-op wrappers below take the same struct.
 *   reference: vo/models/vo_cnn.py:110-233, model_utils/visual_encoders/resnet.py:29-223,
 *   model_utils/running_mean_and_var.py:22-63, rl/policies/resnet_policy.py:146-174.
 */
enum pnvo_opcode {
  PNVO_OP_ZERO = 1,           /* p0 <- 0 (i0 bytes) */
  PNVO_OP_ASSEMBLE = 2,       /* NHWC fp32 sources -> normalised fp16 NHWC (vo_cnn.py:110-176) */
  PNVO_OP_INPUT_STATS = 3,    /* RunningMeanAndVar batch statistics (running_mean_and_var.py:24-38) */
  PNVO_OP_RMV_UPDATE = 4,     /* RunningMeanAndVar Chan merge + scale/shift (:41-63) */
  PNVO_OP_CONV = 5,           /* implicit-GEMM conv (fprop or dgrad) + GroupNorm partial sums, tcgen05 */
  PNVO_OP_WGRAD = 6,          /* weight gradient, tcgen05, split over pixels */
  PNVO_OP_GN_APPLY = 7,       /* GroupNorm affine (+residual) (+ReLU)  (resnet.py:39-55) */
  PNVO_OP_GN_POOL = 8,        /* GroupNorm + ReLU + MaxPool 3x3/s2/p1 (resnet.py:165-168) */
  PNVO_OP_GN_BWD_REDUCE = 9,  /* sum dy, sum dy*xhat per (sample, channel) */
  PNVO_OP_GN_BWD_APPLY = 10,  /* dx of GroupNorm(+ReLU) */
  PNVO_OP_GN_POOL_BWD = 11,   /* gradient through MaxPool+ReLU to the GN output, reduce + apply share it */
  PNVO_OP_PACK_W = 12,        /* OIHW fp32 -> [Cout][R][S][Cin_pad] fp16 (+ flipped/transposed for dgrad) */
  PNVO_OP_UNPACK_DW = 13,     /* packed fp32 dW -> OIHW fp32 grad */
  PNVO_OP_HEAD_FWD = 14,      /* Linear(hidden, out_dim) with warp-shuffle dot products (vo_cnn.py:223-227) */
  PNVO_OP_HEAD_BWD = 15,
  PNVO_OP_BIAS_RELU = 16,     /* fc bias (+ReLU) fwd on fp32 GEMM output */
  PNVO_OP_BIAS_RELU_BWD = 17,
  PNVO_OP_MSE_LOSS = 18,      /* vo_cnn_engine.py:135-198 losses + d(loss)/d(pred) */
  PNVO_OP_ADAM = 19,          /* Adam step over a flat fp32 bucket */
  PNVO_OP_AVGPOOL2 = 20,      /* F.avg_pool2d(x, 2) -> fp16 NHWC (resnet_policy.py:168) */
  PNVO_OP_GN_PARAM_GRAD = 21, /* dgamma/dbeta from the per-(sample,channel) sums */
  PNVO_OP_CAST = 22,          /* fp32 <-> fp16 copies with channel padding */
  PNVO_OP_DROPOUT = 23,       /* in-place inverted dropout, counter-based generator (vo_cnn.py:218,224) */
  PNVO_OP_CONV_STEM = 24,     /* 7x7/s2 stem conv, row-raster operands (no im2col expansion), tcgen05 */
  PNVO_OP_PACK_W_STEM = 25,   /* OIHW fp32 -> [r][tap pair][cout][64] fp16 for the stem kernel */
  PNVO_OP_WGRAD_STEM = 26,    /* stem weight gradient on the row raster */
  PNVO_OP_GN_BWD_FUSED = 27, /* GroupNorm(+ReLU) backward in one pass: a thread-block cluster per sample, DSMEM reduce */
  PNVO_OP_RAW_STATS = 28,    /* RunningMeanAndVar batch statistics straight from uint8 rgb / fp32 depth / top-down pairs */
  PNVO_OP_RAW_ASSEMBLE = 29, /* raw pairs -> [rgb/255, depth, one-hot depth bins, top-down] x {prev, cur} -> normalised fp16 NHWC */
  PNVO_OP_ACT_EMBED_FWD = 30, /* z += W[:, col0:col0+dim] . (Embedding(action) * dropout mask)  (vo_cnn_act_embed.py:65-75) */
  PNVO_OP_ACT_EMBED_BWD = 31, /* gradients of the embedding columns of the hidden Linear and of the embedding table */
  PNVO_OP_UPSAMPLE2 = 32,     /* zero-insertion x2 upsampling (stride-2 data gradients as stride-1 convolutions) */
  PNVO_OP_GEO_INV_LOSS = 33,  /* geometric-inversion loss + gradient (vo_cnn_regression_geo_invariance_engine.py:367-449) */
  PNVO_OP_CONV_STEM2 = 34,    /* stem conv, pixels-as-N formulation: D[(4 rows x cout), ow], resident weights, persistent */
  PNVO_OP_PACK_W_STEM2 = 35,  /* OIHW fp32 -> [tap pair][descending filter rows by parity][cout][64] fp16 */
  PNVO_OP_WGRAD_STEM2 = 36,   /* stem weight gradient with a window of four dy rows as the UMMA N dimension (N = 128) */
  PNVO_OP_PACK_W_MULTI = 37,  /* PACK_W over a device table of descriptors (one launch for every layer) */
  PNVO_OP_UNPACK_DW_MULTI = 38,
  PNVO_OP_GN_PARAM_GRAD_MULTI = 39,
  PNVO_OP_STEM_EXACT_PREP = 40,   /* exact-input stem: per-channel (scale, shift, a, b, weight source) from the input statistics */
  PNVO_OP_STEM_EXACT_PACK = 41,   /* W' = a_c W in the stem layout (value + residual planes) + the [5][5][32] border bias table */
  PNVO_OP_STEM_DY_SUMS = 42,      /* per-border-class sums of the stem's output gradient */
  PNVO_OP_STEM_EXACT_UNPACK = 43, /* conv1 weight gradient from the exact-tensor gradient and the border-class sums */
  PNVO_OP_JOIN = 44,              /* graph replay: everything issued on the side lane so far completes before the next op */
  PNVO_OP_MAX = 45
};

/* `code` may carry a lane in bits 16+: PNVO_OP_SIDE_LANE marks an op whose result is not needed until the next
 * PNVO_OP_JOIN (or the end of the program).  pnvo_run_ops ignores lanes (strictly sequential); a captured graph puts such
 * ops on a forked branch that depends on every op before them, so they overlap the ops that follow (weight gradients
 * under the data-gradient / GroupNorm-backward chain). */
#define PNVO_OP_SIDE_LANE (1 << 16)

typedef struct {
  int32_t code;
  int32_t i[27];
  float f[4];
  void* p[10];
} pnvo_op;

/* Executes ops[0..n_ops) in order on `stream`.  `ops` is HOST memory. */
int pnvo_run_ops(const pnvo_op* ops, int n_ops, void* stream);

/* CUDA-graph replay of a program whose device buffers do not change between runs: capture once (nothing executes),
 * then every pnvo_graph_launch replays all of its kernels with a single cudaGraphLaunch on `stream`. */
int pnvo_graph_capture(const pnvo_op* ops, int n_ops, void** handle_out);
int pnvo_graph_launch(void* handle, void* stream);
int pnvo_graph_destroy(void* handle);

/* Dynamic shared memory / TMEM columns / grid the conv op would use (for tests and DESIGN.md tables). */
int pnvo_conv_launch_info(const pnvo_op* op, int32_t* grid_x, int32_t* grid_y, int32_t* smem_bytes,
                          int32_t* tmem_cols, int32_t* stages);

/* Width (pixels) of the zero-padded input rows the stem kernel (PNVO_OP_CONV_STEM) expects: 3 zero pixels left of the
 * image, zeros on the right; the image starts at pixel 3 of every row. */
int pnvo_stem_padded_width(int IW);

/* 1 when PNVO_OP_CONV_STEM2 can take an [IH, IW] input (output width <= 240). */
int pnvo_conv_stem2_supported(int IH, int IW);
/* 1 when PNVO_OP_WGRAD_STEM2 can take it (output width <= 176). */
int pnvo_conv_stem_wgrad2_supported(int IH, int IW);

/* 1 when PNVO_OP_GN_BWD_FUSED can take a [HW, C] fp16 sample (else use GN_BWD_REDUCE + GN_BWD_APPLY). */
int pnvo_gn_bwd_fused_supported(int C, int HW, int x_fp32);

/* ---------------------------------------------------------------------------------------------
 * e  multi-GPU (one process per GPU on one NVLink / NVSwitch node): the data-parallel gradient exchange fused with the
 *    optimiser step over peer memory.  Replaces, for the flat parameter / gradient buckets of the VO and PPO trainers,
 *    DistributedDataParallel's bucketed all-reduce + torch.optim.Adam.step()
 *    (pointnav_vo/rl/ddppo/algo/ddppo.py:55-96; optimiser: vo/engine/vo_cnn_regression_geo_invariance_engine.py:122-133,
 *    rl/ppo/ppo.py:53-58).
 * pnvo_peer_alloc: the one place the library allocates: a zero-filled cudaMalloc region that other processes can map,
 *   and its 64-byte CUDA IPC handle (exchanged by the caller, e.g. with torch.distributed.all_gather).
 * pnvo_peer_open / pnvo_peer_close: map / unmap a peer's region in this process (enables peer access lazily).
 * pnvo_peer_reduce_adam: ONE kernel per rank and step.  grads / params / flags are HOST arrays of `world` device
 *   pointers (entry `rank` = this rank's own buffers, the others = mapped peer regions).  Rank r reduces slice r of the
 *   gradient buckets of all ranks in rank order (loads over NVLink), applies torch.optim.Adam (weight_decay 0) to its
 *   slice of m / v / params and stores the new parameters into every rank's parameter bucket (stores over NVLink).
 *   n: bucket length in floats (multiple of 4; buckets 16-byte aligned); flags: >= 128 zero-initialised bytes per rank;
 *   seq: 1, 2, 3, ... (the same on every rank per call; sequence flags are never reset); step: Adam's t.
 *   On return of the KERNEL every rank's parameters are updated and every rank's gradient bucket may be overwritten.
 *   A peer that does not arrive within ~4 s sets flags[rank][17] != 0 instead of hanging the device.
 */
int pnvo_peer_alloc(int64_t bytes, void** ptr_out, void* handle64_out);
int pnvo_peer_open(const void* handle64, void** ptr_out);
int pnvo_peer_close(void* ptr);
int pnvo_peer_free(void* ptr);
int pnvo_peer_reduce_adam(const void* const* grads, void* const* params, void* const* flags, float* m, float* v,
                          int64_t n, int rank, int world, uint32_t seq, float lr, float beta1, float beta2, float eps,
                          int step, void* stream);

/* All-reduce (sum, fixed rank order) of a small fp64 vector through peer memory -- the packed RunningMeanAndVar batch
 * statistics (pointnav_vo/model_utils/running_mean_and_var.py:28-38: three distrib.all_reduce calls there).
 * slots / flags: HOST arrays of `world` device pointers; every rank provides a zero-initialised exchange area of
 * world * 2 * 1024 doubles and a flag block of >= 128 bytes (peer-mapped for the other ranks).  data: n <= 1024 doubles on
 * this rank, replaced by the sum over ranks.  seq: 1, 2, 3, ... per call, equal on every rank. */
int pnvo_peer_sum_f64(void* const* slots, void* const* flags, double* data, int n, int rank, int world, uint32_t seq,
                      void* stream);

/* Number of kernels launched by this library since load (bench.py's gpu_launches). */
int64_t pnvo_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PNVO_H_ */
