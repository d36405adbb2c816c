"""ctypes binding of libpnvo.so (include/pnvo.h) and builders for `pnvo_op` records.

There is NO fallback: if the library is missing, or a CUDA tensor is handed to an op on a device
that is not sm_100, the call raises.  PyTorch is used only to own device memory and streams.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpnvo.so")

# opcodes (include/pnvo.h: enum pnvo_opcode)
OP_ZERO, OP_ASSEMBLE, OP_INPUT_STATS, OP_RMV_UPDATE, OP_CONV, OP_WGRAD, OP_GN_APPLY, OP_GN_POOL = range(1, 9)
OP_GN_BWD_REDUCE, OP_GN_BWD_APPLY, OP_GN_POOL_BWD, OP_PACK_W, OP_UNPACK_DW, OP_HEAD_FWD, OP_HEAD_BWD = range(9, 16)
OP_BIAS_RELU, OP_BIAS_RELU_BWD, OP_MSE_LOSS, OP_ADAM, OP_AVGPOOL2, OP_GN_PARAM_GRAD, OP_CAST, OP_DROPOUT, OP_CONV_STEM, OP_PACK_W_STEM, OP_WGRAD_STEM, OP_GN_BWD_FUSED, OP_RAW_STATS, OP_RAW_ASSEMBLE, OP_ACT_EMBED_FWD, OP_ACT_EMBED_BWD, OP_UPSAMPLE2, OP_GEO_INV_LOSS, OP_CONV_STEM2, OP_PACK_W_STEM2, OP_WGRAD_STEM2, OP_PACK_W_MULTI, OP_UNPACK_DW_MULTI, OP_GN_PARAM_GRAD_MULTI = range(16, 40)
OP_STEM_EXACT_PREP, OP_STEM_EXACT_PACK, OP_STEM_DY_SUMS, OP_STEM_EXACT_UNPACK = range(40, 44)
OP_JOIN = 44
SIDE_LANE = 1 << 16  # include/pnvo.h: PNVO_OP_SIDE_LANE


def side(op):
    """Marks an op for the side lane of a captured graph: it may overlap the ops that follow until the next op_join()."""
    op.code |= SIDE_LANE
    return op


def op_join():
    return _op(OP_JOIN)


class PnvoOp(ctypes.Structure):
    _fields_ = [("code", ctypes.c_int32), ("i", ctypes.c_int32 * 27), ("f", ctypes.c_float * 4),
                ("p", ctypes.c_void_p * 10)]


class TopdownConsts(ctypes.Structure):
    _fields_ = [("min_x", ctypes.c_float), ("x_den", ctypes.c_float), ("z_den", ctypes.c_float),
                ("depth_scale", ctypes.c_float), ("depth_off", ctypes.c_float),
                ("rows_around_center", ctypes.c_int), ("center_crop", ctypes.c_int)]


class PackDesc(ctypes.Structure):  # csrc/elem.cuh
    _fields_ = [("w", ctypes.c_void_p), ("wp", ctypes.c_void_p), ("wt", ctypes.c_void_p)] + \
               [(n, ctypes.c_int32) for n in ("Cout", "Cin", "R", "S", "cin_pad", "ld_p", "cout_pad", "ld_t", "t_mode", "src_ld")]


class UnpackDesc(ctypes.Structure):
    _fields_ = [("dwp", ctypes.c_void_p), ("grad", ctypes.c_void_p)] + \
               [(n, ctypes.c_int32) for n in ("Cout", "Cin", "R", "S", "cin_pad", "ld_p", "accumulate", "dst_ld")]


class GnParamDesc(ctypes.Structure):
    _fields_ = [("sums", ctypes.c_void_p), ("dgamma", ctypes.c_void_p), ("dbeta", ctypes.c_void_p),
                ("C", ctypes.c_int32), ("C_real", ctypes.c_int32)]


def device_table(descs, device):
    """Uploads a list of ctypes descriptor structs as one device byte tensor (kept alive by the caller)."""
    arr = (type(descs[0]) * len(descs))(*descs)
    host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    return host.to(device)


def op_multi(code, table, n, B=0):
    return _op(code, [n, B], (), [table])


class PnvoError(RuntimeError):
    pass


_lib = None

EXPORTS = ["pnvo_last_error", "pnvo_abi_version", "pnvo_check_device", "pnvo_discretize_depth",
           "pnvo_topdown_project", "pnvo_gae_scan", "pnvo_goal_update", "pnvo_run_ops", "pnvo_conv_launch_info",
           "pnvo_launch_count", "pnvo_stem_padded_width", "pnvo_gn_bwd_fused_supported", "pnvo_topdown_project_strided", "pnvo_topdown_project_strided_f16",
           "pnvo_conv_stem2_supported", "pnvo_conv_stem_wgrad2_supported", "pnvo_graph_capture", "pnvo_graph_launch",
           "pnvo_graph_destroy", "pnvo_ppo_loss", "pnvo_peer_alloc", "pnvo_peer_open", "pnvo_peer_close", "pnvo_peer_free",
           "pnvo_peer_reduce_adam", "pnvo_peer_sum_f64"]


def load():
    """Loads libpnvo.so (building is __graft_entry__.build()'s job); raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PnvoError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU / PyTorch fallback for the CUDA path)")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
    lib.pnvo_last_error.restype = ctypes.c_char_p
    lib.pnvo_abi_version.restype = i32
    lib.pnvo_check_device.restype = i32
    lib.pnvo_launch_count.restype = i64
    lib.pnvo_stem_padded_width.argtypes = [i32]
    lib.pnvo_stem_padded_width.restype = i32
    lib.pnvo_conv_stem2_supported.argtypes = [i32, i32]
    lib.pnvo_conv_stem_wgrad2_supported.argtypes = [i32, i32]
    lib.pnvo_conv_stem_wgrad2_supported.restype = i32
    lib.pnvo_conv_stem2_supported.restype = i32
    lib.pnvo_gn_bwd_fused_supported.argtypes = [i32, i32, i32]
    lib.pnvo_gn_bwd_fused_supported.restype = i32
    lib.pnvo_discretize_depth.argtypes = [vp, i64, vp, i32, vp, i64, vp, vp, vp]
    lib.pnvo_topdown_project.argtypes = [vp, i64, i32, i32, i32, vp, ctypes.POINTER(TopdownConsts), vp, i64, i64, vp, vp]
    lib.pnvo_topdown_project_strided.argtypes = [vp, i64, i64, i32, i32, i32, vp, ctypes.POINTER(TopdownConsts), vp, i64,
                                                 i64, vp, vp]
    lib.pnvo_topdown_project_strided.restype = i32
    lib.pnvo_topdown_project_strided_f16.argtypes = lib.pnvo_topdown_project_strided.argtypes
    lib.pnvo_topdown_project_strided_f16.restype = i32
    lib.pnvo_gae_scan.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, f32, f32, i32, vp]
    lib.pnvo_goal_update.argtypes = [vp, vp, vp, i32, vp]
    lib.pnvo_ppo_loss.argtypes = [vp, vp, vp, vp, vp, vp, i64, f32, i32, f32, vp, vp, vp, vp]
    lib.pnvo_ppo_loss.restype = i32
    lib.pnvo_run_ops.argtypes = [ctypes.POINTER(PnvoOp), i32, vp]
    lib.pnvo_graph_capture.argtypes = [ctypes.POINTER(PnvoOp), i32, ctypes.POINTER(ctypes.c_void_p)]
    lib.pnvo_graph_launch.argtypes = [vp, vp]
    lib.pnvo_graph_destroy.argtypes = [vp]
    for n in ("pnvo_graph_capture", "pnvo_graph_launch", "pnvo_graph_destroy"):
        getattr(lib, n).restype = i32
    lib.pnvo_conv_launch_info.argtypes = [ctypes.POINTER(PnvoOp)] + [ctypes.POINTER(ctypes.c_int32)] * 5
    for n in ("pnvo_discretize_depth", "pnvo_topdown_project", "pnvo_gae_scan", "pnvo_goal_update", "pnvo_run_ops",
              "pnvo_conv_launch_info"):
        getattr(lib, n).restype = i32
    lib.pnvo_peer_alloc.argtypes = [i64, ctypes.POINTER(ctypes.c_void_p), vp]
    lib.pnvo_peer_open.argtypes = [vp, ctypes.POINTER(ctypes.c_void_p)]
    lib.pnvo_peer_close.argtypes = [vp]
    lib.pnvo_peer_free.argtypes = [vp]
    lib.pnvo_peer_reduce_adam.argtypes = [ctypes.POINTER(ctypes.c_void_p)] * 3 + [vp, vp, i64, i32, i32, ctypes.c_uint32,
                                          f32, f32, f32, f32, i32, vp]
    lib.pnvo_peer_sum_f64.argtypes = [ctypes.POINTER(ctypes.c_void_p)] * 2 + [vp, i32, i32, i32, ctypes.c_uint32, vp]
    for n in ("pnvo_peer_alloc", "pnvo_peer_open", "pnvo_peer_close", "pnvo_peer_free", "pnvo_peer_reduce_adam",
              "pnvo_peer_sum_f64"):
        getattr(lib, n).restype = i32
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise PnvoError(load().pnvo_last_error().decode())


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise PnvoError("libpnvo ops need CUDA tensors (no CPU fallback)")
    return ctypes.c_void_p(t.data_ptr())


def launch_count():
    return int(load().pnvo_launch_count())


# ------------------------------------------------------------------------------------------------
# op record builders
# ------------------------------------------------------------------------------------------------
def _op(code, ints=(), floats=(), ptrs=()):
    op = PnvoOp()
    op.code = code
    for k, v in enumerate(ints):
        op.i[k] = int(v)
    for k, v in enumerate(floats):
        op.f[k] = float(v)
    for k, v in enumerate(ptrs):
        op.p[k] = (v.data_ptr() if isinstance(v, torch.Tensor) else (v or 0)) or None
    return op


def _lohi(n):
    n = int(n)
    lo, hi = n & 0xFFFFFFFF, (n >> 32) & 0xFFFFFFFF
    return (lo - (1 << 32) if lo >= (1 << 31) else lo), (hi - (1 << 32) if hi >= (1 << 31) else hi)


def op_zero(t):
    lo, hi = _lohi(t.numel() * t.element_size())
    return _op(OP_ZERO, [lo, hi], (), [t])


def _assemble_fields(srcs, nch, pre_scale, lut, C, Cpad, n_pix):
    ints = [len(srcs)] + list(nch) + [0] * (4 - len(nch)) + [C, Cpad, *_lohi(n_pix)]
    words = []
    for c in range(0, 32, 2):
        h = []
        for cc in (c, c + 1):
            si, sc = lut[cc] if cc < len(lut) else (0, 0)
            h.append(((si & 0xFF) << 8) | (sc & 0xFF))
        w = h[0] | (h[1] << 16)
        words.append(w - (1 << 32) if w >= (1 << 31) else w)
    return ints + words, list(pre_scale) + [1.0] * (4 - len(pre_scale))


def op_assemble(srcs, nch, pre_scale, lut, C, Cpad, n_pix, scale, shift, out, row_w=0, out_pitch=0, out_lo=None):
    ints, fl = _assemble_fields(srcs, nch, pre_scale, lut, C, Cpad, n_pix)
    ints = ints + [0] * (25 - len(ints)) + [row_w, out_pitch]
    return _op(OP_ASSEMBLE, ints, fl, list(srcs) + [None] * (4 - len(srcs)) + [scale, shift, out, out_lo])


def op_input_stats(srcs, nch, pre_scale, lut, C, Cpad, n_pix, stats_f64):
    ints, fl = _assemble_fields(srcs, nch, pre_scale, lut, C, Cpad, n_pix)
    return _op(OP_INPUT_STATS, ints, fl, list(srcs) + [None] * (4 - len(srcs)) + [None, None, stats_f64])


def _raw_fields(use_rgb, use_depth, n_dd, use_td, C, Cpad, n_pix, row_w=0, out_pitch=0, hw=0, depth=None, n_lo=0):
    depth_fp16 = int(depth is not None and depth.dtype == torch.float16)
    return [int(use_rgb), int(use_depth), int(n_dd), int(use_td), C, Cpad, *_lohi(n_pix), row_w, out_pitch, hw, depth_fp16,
            int(n_lo)]


def op_raw_stats(rgb_u8, depth, td, edges, use_rgb, use_depth, n_dd, use_td, C, Cpad, n_pix, stats_f64, pair_map=None,
                 hw=0):
    """Batch statistics of the assembled input straight from the raw pairs (csrc/raw_input.cu).  pair_map (int32 per
    OUTPUT sample = 2 * source pair + swap flag) expands / swaps pairs on the fly; n_pix then counts output pixels."""
    return _op(OP_RAW_STATS, _raw_fields(use_rgb, use_depth, n_dd, use_td, C, Cpad, n_pix, hw=hw, depth=depth), (),
               [rgb_u8, depth, td, edges, None, None, stats_f64, None, pair_map])


def op_raw_assemble(rgb_u8, depth, td, edges, use_rgb, use_depth, n_dd, use_td, C, Cpad, n_pix, scale, shift, out,
                    row_w=0, out_pitch=0, out_lo=None, pair_map=None, hw=0, n_lo=0, exact=False, stats_f64=None):
    """n_lo = 2 (exact-input stem): channels C, C+1 carry the fp16 residuals of the two top-down values; scale / shift then
    have C + 2 entries (op_stem_exact_prep).  exact: the exact-input stem's constant storage maps ((byte - 128) / 256 for rgb,
    raw values otherwise; scale / shift ignored).  stats_f64 (needs exact): the batch statistics are accumulated in the
    same pass (replaces op_raw_stats)."""
    return _op(OP_RAW_ASSEMBLE, _raw_fields(use_rgb, use_depth, n_dd, use_td, C, Cpad, n_pix, row_w, out_pitch, hw, depth,
                                            n_lo) + [int(bool(exact))], (),
               [rgb_u8, depth, td, edges, scale, shift, out, out_lo, pair_map, stats_f64])


def op_stem_exact_prep(scale, shift, xp, use_rgb, use_depth, n_dd, use_td):
    """xp [6][32] fp32 <- per-channel constants of the exact-input stem (csrc/stem_exact.cu) from the normaliser's
    scale = 1/std, shift = -mean/std (None: identity)."""
    return _op(OP_STEM_EXACT_PREP, [int(use_rgb), int(use_depth), int(n_dd), int(use_td)], (), [scale, shift, xp])


def op_stem_exact_pack(w, xp, wr, wr_lo, bias5, Cin, IH, IW):
    return _op(OP_STEM_EXACT_PACK, [Cin, IH, IW], (), [w, xp, wr, wr_lo, bias5])


def op_stem_dy_sums(dy, S, B, OH, OW):
    return _op(OP_STEM_DY_SUMS, [B, OH, OW], (), [dy, S])


def op_stem_exact_unpack(dwp, S, xp, grad, w_ld, Cin, IH, IW):
    return _op(OP_STEM_EXACT_UNPACK, [w_ld, Cin, IH, IW], (), [dwp, S, xp, grad])


def op_rmv_update(stats_f64, mean, var, count, scale, shift, C, update, have_rmv, n_batch, pix_per_sample):
    return _op(OP_RMV_UPDATE, [C, int(update), int(have_rmv)], [n_batch, pix_per_sample],
               [stats_f64, mean, var, count, scale, shift])


def op_conv(x, w, y, B, IH, IW, Cin, OH, OW, R, S, mul, pad, div, w_ld, n_total, n_store, ldo, add=None, stats=None,
            cpg=0, G=0, out_fp32=False, pad_w=None, x_lo=None, w_lo=None, y_lo=None, cin_real=0, o_mul=0, o_off=(0, 0),
            o_hw=(0, 0), pad_hi=None, x_c=None):
    """x_lo / w_lo: residual planes of the split-fp16 representation (x - fp16(x)); with them the kernel accumulates
    x*w + x_lo*w + x*w_lo, i.e. products of ~fp32-precision operands (3 MMAs per product).
    cin_real: input channels that are not zero padding (0 = unknown); 1 selects the direct fp32 stem kernel.
    o_mul / o_off / o_hw: strided output -- GEMM pixel (b, oh, ow) is stored at (b, oh * o_mul + o_off[0], ow * o_mul +
    o_off[1]) of a [B, o_hw[0], o_hw[1]] tensor (and `add` is read there); pad_hi = (h, w): high-side padding when it differs
    from the low-side `pad` / `pad_w` (parity-class data gradient of stride-2 convolutions)."""
    asym = 0 if pad_hi is None else ((pad_hi[0] + 1) | ((pad_hi[1] + 1) << 8))
    return _op(OP_CONV, [B, IH, IW, Cin, OH, OW, R, S, mul, pad, div, w_ld, n_total, n_store, ldo, cpg, G,
                         int(out_fp32), pad if pad_w is None else pad_w, 0, 0, int(cin_real), int(o_mul),
                         o_off[0] | (o_off[1] << 8), o_hw[0], o_hw[1], asym], (),
               [x, w, y, add, stats, x_lo, w_lo, y_lo, x_c])


def op_wgrad(x, dy, dw, B, IH, IW, Cin, OH, OW, R, S, mul, pad, w_ld, n_total, ld_dy, pad_w=None, x_row_pitch=0,
             cin_real=0, x_c=None):
    return _op(OP_WGRAD, [B, IH, IW, Cin, OH, OW, R, S, mul, pad, 1, w_ld, n_total, 0, ld_dy, 0, 0, 0,
                          pad if pad_w is None else pad_w, 0, x_row_pitch, int(cin_real)], (), [x, dy, dw, x_c])


def op_gn_apply(x, stats, gamma, beta, y, B, C, G, cpg, HW, cnt, relu=True, res=None, x_fp32=False, eps=1e-5,
                C_real=None, y_lo=None, res_lo=None, x_lo=None):
    return _op(OP_GN_APPLY, [B, C, G, cpg, HW, int(relu), int(x_fp32), 0, 0, 0, 0, C if C_real is None else C_real],
               [cnt, eps], [x, stats, gamma, beta, res, y, None, y_lo, res_lo, x_lo])


def op_gn_pool(x, stats, gamma, beta, y, argmax, B, C, G, cpg, H, W, PH, PW, cnt, x_fp32=False, eps=1e-5,
               C_real=None, y_lo=None, x_lo=None):
    return _op(OP_GN_POOL, [B, C, G, cpg, H * W, 1, int(x_fp32), H, W, PH, PW, C if C_real is None else C_real],
               [cnt, eps],
               [x, stats, gamma, beta, None, y, argmax, y_lo, None, x_lo])


def op_pool_bwd(g, pooled, argmax, dy, B, C, H, W, PH, PW):
    return _op(OP_GN_POOL_BWD, [B, C, 0, 0, 0, 0, 0, H, W, PH, PW], (), [g, pooled, argmax, dy])


def op_gn_bwd(reduce, g, relu_ref, x, stats, gamma, sums, dx, dy_out, B, C, G, cpg, HW, cnt, x_fp32=False, eps=1e-5,
              C_real=None, g_scale=1.0, class_sums=None, ohw=(0, 0)):
    """class_sums (apply pass, exact-input stem): fp32 [5][5][32] border-class sums of dx, accumulated (pre-zeroed)."""
    code = OP_GN_BWD_FUSED if reduce == "fused" else (OP_GN_BWD_REDUCE if reduce else OP_GN_BWD_APPLY)
    return _op(code,
               [B, C, G, cpg, HW, 0, int(x_fp32), ohw[0], ohw[1], 0, 0, C if C_real is None else C_real], [cnt, eps, g_scale],
               [g, relu_ref, x, stats, gamma, sums, dx, dy_out, class_sums if code == OP_GN_BWD_APPLY else None])


def op_gn_param_grad(sums, dgamma, dbeta, B, C, C_real, accumulate=False):
    return _op(OP_GN_PARAM_GRAD, [B, C, C_real, int(accumulate)], (), [sums, dgamma, dbeta])


def op_pack_w(w, wp, wt, Cout, Cin, R, S, cin_pad, ld_p, cout_pad=0, ld_t=0, t_mode=0, src_ld=0):
    return _op(OP_PACK_W, [Cout, Cin, R, S, cin_pad, ld_p, cout_pad, ld_t, t_mode, src_ld], (), [w, wp, wt])


def op_unpack_dw(dwp, grad, Cout, Cin, R, S, cin_pad, ld_p, accumulate=False, dst_ld=0):
    return _op(OP_UNPACK_DW, [Cout, Cin, R, S, cin_pad, ld_p, int(accumulate), dst_ld], (), [dwp, grad])


def op_geo_inv_loss(pred, actions, dout, loss3, B, O, weight, grad_scale=1.0, move_forward=1, data_types=None, err=None,
                    turn_left=2, turn_right=3):
    """loss3[0] += weight * (rot + pos), loss3[1] = rot, loss3[2] = pos; dout += weight * grad_scale * gradient.
    data_types (int64 per row): only TURN_LEFT / TURN_RIGHT rows pair up, in batch order, and must alternate
    [cur-rel-to-prev, prev-rel-to-cur] (else *err = 1); None: every row, interleaved."""
    return _op(OP_GEO_INV_LOSS, [B, O, move_forward, turn_left, turn_right], [weight, grad_scale],
               [pred, actions, dout, loss3, data_types, err])


def op_upsample2(src, dst, B, OH, OW, IH, IW, C):
    return _op(OP_UPSAMPLE2, [B, OH, OW, IH, IW, C], (), [src, dst])


def op_act_embed_fwd(z, W, E, actions, e_used, mask, seed, B, hidden, dim, n_rows, w_ld, col0, p_drop=0.0):
    return _op(OP_ACT_EMBED_FWD, [B, hidden, dim, n_rows, w_ld, col0, 0], [p_drop],
               [z, W, E, actions, e_used, mask, seed])


def op_act_embed_bwd(dz16, W, E, actions, e_used, mask, dW, dE, B, hidden, dim, n_rows, w_ld, col0, dz_ld):
    return _op(OP_ACT_EMBED_BWD, [B, hidden, dim, n_rows, w_ld, col0, dz_ld], (),
               [dz16, W, E, actions, e_used, mask, dW, dE])


def op_bias_relu(z, bias, h32, h16, B, N, relu=True):
    return _op(OP_BIAS_RELU, [B, N, int(relu)], (), [z, bias, h32, h16])


def op_bias_relu_bwd(dh, h, dz16, db, B, N, accumulate=False):
    return _op(OP_BIAS_RELU_BWD, [B, N, int(accumulate)], (), [dh, h, dz16, db])


def op_head_fwd(h, W, bias, out, B, K, O):
    return _op(OP_HEAD_FWD, [B, K, O], (), [h, W, bias, out])


def op_head_bwd(dout, h, W, dW, db2, dz16, db1, B, K, O, accumulate=False, dh_scale=1.0):
    return _op(OP_HEAD_BWD, [B, K, O, int(accumulate)], [dh_scale], [dout, h, W, dW, db2, dz16, db1])


def op_dropout(buf, seed, site, p, advance=False):
    lo, hi = _lohi(buf.numel())
    return _op(OP_DROPOUT, [lo, hi, int(buf.dtype == torch.float16), site, int(advance)], [p], [buf, seed])


def op_mse_loss(pred, target, dz_mask, dout, loss, B, O, weights=(1.0, 1.0, 1.0), grad_scale=1.0, data_types=None):
    """data_types (int64 per row, 0 = cur-rel-to-prev, 1 = prev-rel-to-cur): one mean per data type, summed."""
    return _op(OP_MSE_LOSS, [B, O], [weights[0], weights[1], weights[2], grad_scale],
               [pred, target, dz_mask, dout, loss, data_types])


def op_conv_stem(x, wr, y, stats, B, IH, IW, G, cpg, stages=4):
    return _op(OP_CONV_STEM, [B, IH, IW, G, cpg, stages], (), [x, wr, y, stats])


def op_conv_stem2(x, wr, y, stats, B, IH, IW, G, cpg, x_lo=None, add=None, out_fp32=False, bias5=None, y_lo=None):
    """x_lo: residual plane of x (split mode: every row is multiplied twice against the same weights); add: fp16
    [B, OH, OW, 32] added in the epilogue; out_fp32: y is fp32; bias5: [5][5][32] fp32 border-class bias (exact-input stem)."""
    return _op(OP_CONV_STEM2, [B, IH, IW, G, cpg, int(out_fp32)], (), [x, wr, y, stats, x_lo, add, bias5, y_lo])


def op_wgrad_stem2(x, dy, dw, B, IH, IW, w_ld):
    return _op(OP_WGRAD_STEM2, [B, IH, IW, w_ld], (), [x, dy, dw])


def op_pack_w_stem2(w, wr, Cin, lo=False):
    return _op(OP_PACK_W_STEM2, [Cin, int(lo)], (), [w, wr])


def op_wgrad_stem(x, dy, dw, B, IH, IW, w_ld, rows_per_cta=32):
    return _op(OP_WGRAD_STEM, [B, IH, IW, w_ld, rows_per_cta], (), [x, dy, dw])


def op_pack_w_stem(w, wr, Cin):
    return _op(OP_PACK_W_STEM, [Cin], (), [w, wr])


def op_adam(p, g, m, v, n, step, lr, beta1, beta2, eps):
    lo, hi = _lohi(n)
    return _op(OP_ADAM, [lo, hi, step], [lr, beta1, beta2, eps], [p, g, m, v])


def op_avgpool2(src, out, B, H, W, C, Cpad, coff, pre_scale=1.0, out_lo=None, out32=None, ld32=0):
    return _op(OP_AVGPOOL2, [B, H, W, C, Cpad, coff, ld32], [pre_scale], [src, out, out_lo, out32])


USE_GRAPHS = os.environ.get("PNVO_GRAPHS", "1") != "0"


class Program:
    """A list of pnvo_op records packed into one ctypes array, replayed with a single C call.
    graph=True: after one eager run (which also sets kernel attributes and fills the tensor-map cache) the program
    is captured into a CUDA graph and later runs are one cudaGraphLaunch -- for programs whose buffers are fixed."""

    def __init__(self, ops, graph=False):
        self.n = len(ops)
        self.arr = (PnvoOp * max(1, self.n))(*ops)
        self.graph = bool(graph) and USE_GRAPHS
        self._runs = 0
        self._handle = None

    def run(self, device=None):
        if not self.n:
            return
        lib = load()
        if self.graph and self._runs >= 1:
            if self._handle is None:
                h = ctypes.c_void_p()
                check(lib.pnvo_graph_capture(self.arr, self.n, ctypes.byref(h)))
                self._handle = h
            check(lib.pnvo_graph_launch(self._handle, stream_ptr(device)))
        else:
            check(lib.pnvo_run_ops(self.arr, self.n, stream_ptr(device)))
        self._runs += 1

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None and _lib is not None:
            try:
                _lib.pnvo_graph_destroy(h)
            except Exception:
                pass


def patch_ops(ops, mapping):
    """Copies of `ops` with every pointer slot that equals a key of `mapping` replaced by its value: the same program
    over an alternate buffer (e.g. the second input staging buffer of a double-buffered pipeline)."""
    out = []
    for op in ops:
        c = PnvoOp()
        ctypes.memmove(ctypes.byref(c), ctypes.byref(op), ctypes.sizeof(PnvoOp))
        for k in range(10):
            v = c.p[k]
            if v is not None and v in mapping:
                c.p[k] = mapping[v]
        out.append(c)
    return out


def run_ops(ops, device=None):
    Program(list(ops)).run(device)


def conv_launch_info(op):
    vals = [ctypes.c_int32() for _ in range(5)]
    check(load().pnvo_conv_launch_info(ctypes.byref(op), *[ctypes.byref(v) for v in vals]))
    return dict(zip(("grid_x", "grid_y", "smem_bytes", "tmem_cols", "stages"), (v.value for v in vals)))
