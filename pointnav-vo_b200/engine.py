"""Builds and replays the op programs (libpnvo `pnvo_op` lists) of the GroupNorm-ResNet encoders.

One `EncoderPlan` per (architecture, batch size, device): it owns every activation / gradient /
workspace buffer (torch tensors used purely as device memory), the packed fp16 weights, and three
programs -- pack (weights -> kernel layouts), forward, backward -- that each run with a single C call.

Reference semantics: vo/models/vo_cnn.py:110-233 (ResNetEncoder + VisualOdometryCNNBase),
model_utils/visual_encoders/resnet.py:29-223 (BasicBlock / Bottleneck / ResNet),
rl/policies/resnet_policy.py:61-174 (RL ResNetEncoder).
Data layout: activations NHWC fp16 (channels padded to a multiple of 8, conv outputs to 32), raw conv
outputs fp16 (fp32 accumulate / GroupNorm statistics), weights [Cout][R][S][Cin] fp16.
"""
import math
import os

import torch

from . import lib as L

RESNET_LAYERS = {"resnet18": ("basic", [2, 2, 2, 2]), "resnet50": ("bottleneck", [3, 4, 6, 3]),
                 "resnet101": ("bottleneck", [3, 4, 23, 3])}


def _ru(x, m):
    return (x + m - 1) // m * m


def _pow2(x):
    p = 8
    while p < x:
        p *= 2
    return p


# exact-input stem: also issue the residual-weight product W'_lo * x (second launch of the stem kernel).  Without it the stem
# weights carry one fp16 rounding (PNVO_STEM_WLO=0; measured effect on the network output in tools/stem_wlo_probe.py).
STEM_WLO = os.environ.get("PNVO_STEM_WLO", "1") != "0"


class ConvLayer:
    """Geometry + packed-weight buffers of one convolution (or Linear seen as a 1x1 convolution)."""

    def __init__(self, key, Cin, Cout, R, S, stride, pad, IH, IW, need_dgrad=True, src_shape=None, cin_pad=None):
        self.key, self.Cin, self.Cout, self.R, self.S, self.stride, self.pad = key, Cin, Cout, R, S, stride, pad
        self.IH, self.IW = IH, IW
        self.OH = (IH + 2 * pad - R) // stride + 1
        self.OW = (IW + 2 * pad - S) // stride + 1
        if cin_pad is None:
            cin_pad = _ru(Cin, 8) if R * S == 1 else _pow2(Cin)
        self.cin_pad = cin_pad
        self.cout_pad = _ru(Cout, 32)
        self.K = R * S * self.cin_pad
        self.w_ld = _ru(self.K, 64)
        self.need_dgrad = need_dgrad
        # dgrad runs the same kernel with Cin' = cout_pad (must be a power of two unless 1x1)
        self.Kt = R * S * self.cout_pad
        self.wt_ld = _ru(self.Kt, 64)
        self.nt_total = _ru(self.cin_pad, 16)
        # how the OIHW source is viewed by the pack kernel (Linear layers: [Cout, C, H, W] of the flatten)
        self.src_shape = src_shape or (Cout, Cin, R, S)
        self.wp = self.wt = self.dwp = self.wp_lo = None
        self.up = None          # zero-upsampled dy (stride-2 data gradient), allocated on first use
        self.fast_s2 = True
        # 3x3 / stride 2 / pad 1: the data gradient runs as four parity-class convolutions of dy (1 + 2 + 2 + 4 taps for four
        # input pixels instead of 9 taps per pixel over a zero-upsampled dy); wt then holds the four class matrices
        # (32 input channels: N = 32 class GEMMs with 64-byte scattered rows measured slower than the upsampled raster route,
        # 114 vs 98 us for layer2.0 at B = 256)
        self.s2_classes = (R == 3 and S == 3 and stride == 2 and pad == 1 and need_dgrad and self.cin_pad >= 64
                           and os.environ.get("PNVO_S2_CLASSES", "1") != "0")
        if self.s2_classes:
            self.wt_ld = _ru(4 * self.cout_pad, 64)

    def alloc(self, dev, training, own_dwp=True, split=False):
        self.wp = torch.zeros(self.cout_pad, self.w_ld, dtype=torch.float16, device=dev)
        if split:  # residual plane w - fp16(w) of the split-fp16 representation
            self.wp_lo = torch.zeros(self.cout_pad, self.w_ld, dtype=torch.float16, device=dev)
        if self.need_dgrad and training:
            rows = 4 * self.nt_total if self.s2_classes else self.nt_total
            self.wt = torch.zeros(rows, self.wt_ld, dtype=torch.float16, device=dev)
        if training and own_dwp:  # inside a plan dwp is a view of the gradient arena (one zero-fill per step)
            self.dwp = torch.zeros(self.cout_pad, self.w_ld, dtype=torch.float32, device=dev)

    def flops(self, B):
        return 2.0 * B * self.OH * self.OW * self.Cout * self.Cin * self.R * self.S

    # ---- op builders ----
    def op_pack(self, w):
        return L.op_pack_w(w, self.wp, self.wt, self.Cout, self.Cin, self.R, self.S, self.cin_pad, self.w_ld,
                           self.cout_pad, self.wt_ld, 4 if self.s2_classes else 0)

    def op_fwd(self, x, y, B, stats=None, cpg=0, G=0, out_fp32=False, x_lo=None, y_lo=None, x_c=None):
        return L.op_conv(x, self.wp, y, B, self.IH, self.IW, self.cin_pad, self.OH, self.OW, self.R, self.S,
                         self.stride, self.pad, 1, self.w_ld, self.cout_pad, self.cout_pad, self.cout_pad, None, stats,
                         cpg, G, out_fp32, x_lo=x_lo, w_lo=self.wp_lo if x_lo is not None else None, y_lo=y_lo,
                         cin_real=self.Cin, x_c=x_c)

    def op_dgrad(self, dy, gx, B, add=None):
        # gx[b, h, w, c] = sum_{r,s,n} dy[b, (h + pad - r)/stride, (w + pad - s)/stride, n] * W[n, c, r, s]
        assert not self.s2_classes, "parity-class weights: use ops_dgrad"
        return L.op_conv(dy, self.wt, gx, B, self.OH, self.OW, self.cout_pad, self.IH, self.IW, self.R, self.S, 1,
                         self.R - 1 - self.pad, self.stride, self.wt_ld, self.nt_total, self.cin_pad, self.cin_pad, add,
                         None, 0, 0, False, pad_w=self.S - 1 - self.pad)

    def ops_dgrad(self, dy, gx, B, add=None, dev=None):
        """Data gradient as a list of ops.  Stride-2 convolutions go through a zero-upsampled copy of dy so that
        the stride-1 kernels (TMA im2col / shared-memory raster) do the work instead of the generic
        divisibility-testing producer: 3x3/s2 = 3x3/s1 conv of up2(dy) with the flipped weights; 1x1/s2 = compact
        1x1 conv of dy scattered to the even input positions."""
        if self.s2_classes:
            ops = []
            for ph in (0, 1):
                for pw in (0, 1):
                    wk = self.wt[(ph * 2 + pw) * self.nt_total:(ph * 2 + pw + 1) * self.nt_total]
                    # class (ph, pw): dense stride-1 conv of dy with (1 + ph) x (1 + pw) taps, zero beyond the high edge,
                    # computed on the whole dy grid and stored at the (2i + ph, 2j + pw) pixels of gx that exist
                    ops.append(L.op_conv(dy, wk, gx, B, self.OH, self.OW, self.cout_pad, self.OH, self.OW, 1 + ph, 1 + pw, 1,
                                         0, 1, self.wt_ld, self.nt_total, self.cin_pad, self.cin_pad, add, None, 0, 0,
                                         False, pad_w=0, o_mul=2, o_off=(ph, pw), o_hw=(self.IH, self.IW),
                                         pad_hi=(ph, pw)))
            return ops
        if self.stride == 1 or not self.fast_s2:
            return [self.op_dgrad(dy, gx, B, add=add)]
        assert self.stride == 2
        if self.R == 1:
            assert add is None and self.pad == 0
            if self.up is None:
                self.up = torch.empty(B, self.OH, self.OW, self.cin_pad, dtype=torch.float16, device=dy.device)
            compact = L.op_conv(dy, self.wt, self.up, B, self.OH, self.OW, self.cout_pad, self.OH, self.OW, 1, 1, 1, 0, 1,
                                self.wt_ld, self.nt_total, self.cin_pad, self.cin_pad, None, None, 0, 0, False, pad_w=0)
            return [compact, L.op_upsample2(self.up, gx, B, self.OH, self.OW, self.IH, self.IW, self.cin_pad)]
        if self.up is None:
            self.up = torch.empty(B, self.IH, self.IW, self.cout_pad, dtype=torch.float16, device=dy.device)
        conv = L.op_conv(self.up, self.wt, gx, B, self.IH, self.IW, self.cout_pad, self.IH, self.IW, self.R, self.S, 1,
                         self.R - 1 - self.pad, 1, self.wt_ld, self.nt_total, self.cin_pad, self.cin_pad, add, None, 0, 0,
                         False, pad_w=self.S - 1 - self.pad)
        return [L.op_upsample2(dy, self.up, B, self.OH, self.OW, self.IH, self.IW, self.cout_pad), conv]

    def op_wgrad(self, x, dy, B, x_row_pitch=0, x_c=None):
        return L.op_wgrad(x, dy, self.dwp, B, self.IH, self.IW, self.cin_pad, self.OH, self.OW, self.R, self.S,
                          self.stride, self.pad, self.w_ld, self.cout_pad, self.cout_pad, x_row_pitch=x_row_pitch,
                          cin_real=self.Cin, x_c=x_c)

    def op_unpack(self, grad):
        return L.op_unpack_dw(self.dwp, grad, self.Cout, self.Cin, self.R, self.S, self.cin_pad, self.w_ld)

    # descriptor-table entries of the batched pack / unpack launches
    def pack_desc(self, w):
        return L.PackDesc(w.data_ptr(), self.wp.data_ptr(), self.wt.data_ptr() if self.wt is not None else None,
                          self.Cout, self.Cin, self.R, self.S, self.cin_pad, self.w_ld, self.cout_pad, self.wt_ld,
                          4 if self.s2_classes else 0, self.Cin * self.R * self.S)

    def pack_desc_lo(self, w):
        return L.PackDesc(w.data_ptr(), self.wp_lo.data_ptr(), None, self.Cout, self.Cin, self.R, self.S, self.cin_pad,
                          self.w_ld, self.cout_pad, self.wt_ld, 2, self.Cin * self.R * self.S)

    def unpack_desc(self, grad):
        return L.UnpackDesc(self.dwp.data_ptr(), grad.data_ptr(), self.Cout, self.Cin, self.R, self.S, self.cin_pad,
                            self.w_ld, 0, self.Cin * self.R * self.S)


class LinearLayer(ConvLayer):
    """nn.Linear over an NCHW-flattened [C, H, W] feature map == 1x1 conv over the NHWC-flattened map
    (weight columns permuted by the pack kernel; vo_cnn.py:216-221, misc_utils.py:45-47)."""

    def __init__(self, key, C, H, W, c_pad, Cout, src_ld=0):
        super().__init__(key, H * W * c_pad, Cout, 1, 1, 1, 0, 1, 1, True, cin_pad=H * W * c_pad)
        self.fC, self.fH, self.fW, self.c_pad = C, H, W, c_pad
        self.src_ld = src_ld  # columns of the nn.Linear weight when it carries extra (action-embedding) inputs

    def op_pack(self, w):
        # source viewed as OIHW [Cout, C, H, W]; packed column (h*W + w)*c_pad + c; wt = plain transpose
        return L.op_pack_w(w, self.wp, self.wt, self.Cout, self.fC, self.fH, self.fW, self.c_pad, self.w_ld,
                           self.cout_pad, self.wt_ld, 1, src_ld=self.src_ld)

    def op_unpack(self, grad):
        return L.op_unpack_dw(self.dwp, grad, self.Cout, self.fC, self.fH, self.fW, self.c_pad, self.w_ld,
                              dst_ld=self.src_ld)

    def pack_desc(self, w):
        ld = self.src_ld or self.fC * self.fH * self.fW
        return L.PackDesc(w.data_ptr(), self.wp.data_ptr(), self.wt.data_ptr() if self.wt is not None else None,
                          self.Cout, self.fC, self.fH, self.fW, self.c_pad, self.w_ld, self.cout_pad, self.wt_ld, 1, ld)

    def pack_desc_lo(self, w):
        ld = self.src_ld or self.fC * self.fH * self.fW
        return L.PackDesc(w.data_ptr(), self.wp_lo.data_ptr(), None, self.Cout, self.fC, self.fH, self.fW, self.c_pad,
                          self.w_ld, self.cout_pad, self.wt_ld, 3, ld)

    def unpack_desc(self, grad):
        ld = self.src_ld or self.fC * self.fH * self.fW
        return L.UnpackDesc(self.dwp.data_ptr(), grad.data_ptr(), self.Cout, self.fC, self.fH, self.fW, self.c_pad,
                            self.w_ld, 0, ld)


class GNLayer:
    def __init__(self, key, C, G):
        self.key, self.C_real, self.G = key, C, G
        self.C = _ru(C, 32)
        self.cpg_real = C // G
        self.cpg = self.C // G


class EncoderPlan:
    """Programs for: input assembly -> ResNet(GroupNorm) -> compression -> [fc -> ReLU -> head]."""

    def __init__(self, *, params, buffers, B, H, W, in_channels, sources, backbone, baseplanes, ngroups,
                 compression_channels, prefix, head=None, training=True, avgpool_input=False, device="cuda",
                 world_size=1, raw_fp32=False, dropout_p=0.0, split=False, exact_stem=False, grad_bucket=None,
                 compact_input=False):
        """params / buffers: dict name -> CUDA fp32 tensor (reference state_dict names, stable storage).
        sources: list of (obs_key, n_channels, pre_scale) in the reference's concat order.
        head: None or dict(fc_w, fc_b, out_w, out_b, hidden, out_dim).
        split: precision mode of the FORWARD pass.  Every fp16 activation / weight tensor gets a residual plane
        (t - fp16(t), fp16) -- the raw conv outputs too -- and each conv accumulates x*w + x_lo*w + x*w_lo: operands carry
        ~22 mantissa bits, so outputs agree with the fp32 reference to ~1e-5 instead of ~5e-3 (3 MMAs per product; the
        stem / raster / implicit-GEMM kernels all have a split variant).  The backward pass of a split plan reads the
        value planes only (single-pass fp16 operands, fp32 accumulation), also of the raw conv outputs (GroupNorm backward)."""
        self._grad_bucket_arg = grad_bucket
        self._compact_input = bool(compact_input)   # avg-pooled one-channel input as a compact fp32 plane (no normalisation ops)
        self.split = bool(split)
        if self.split:
            raw_fp32 = False  # raw conv outputs are value + residual fp16 planes: same bytes, and backward reads one plane
        # exact_stem (split plans fed with raw uint8 rgb / fp16 depth pairs): the input tensor holds the EXACT raw values
        # (no residual plane) and the normalisation is folded into the stem weights + a border bias (csrc/stem_exact.cu);
        # _alloc clears the flag when the stem kernels cannot take the geometry
        self.exact_stem = bool(exact_stem) and self.split
        self._lo = {}
        self.P, self.Bf = params, buffers
        self.B, self.H, self.W = B, H, W
        self.dev = torch.device(device)
        self.training = training
        self.prefix = prefix
        self.sources = sources
        self.in_channels = in_channels
        self.avgpool_input = avgpool_input
        self.world_size = world_size
        self.raw_fp32 = raw_fp32
        self.head = head
        self.dropout_p = float(dropout_p) if head is not None and head.get("out_dim") else 0.0
        self.fuse_gn_bwd = True
        self.side_lane_wgrad = os.environ.get("PNVO_SIDE_LANE", "1") != "0"
        self.stem_version = 2
        self.batch_small_ops = True  # one launch for all weight packs / gradient unpacks / GN parameter gradients
        self.use_stem2 = False
        self.grads = {}
        self._build_layers(backbone, baseplanes, ngroups, compression_channels)
        self._alloc()
        self._build_programs()

    # ------------------------------------------------------------------------------------------
    def _build_layers(self, backbone, baseplanes, ngroups, comp_ch):
        kind, nblocks = RESNET_LAYERS[backbone]
        pfx = self.prefix + ".backbone"
        H, W = (self.H // 2, self.W // 2) if self.avgpool_input else (self.H, self.W)
        self.inH, self.inW = H, W
        self.cin_pad = _pow2(self.in_channels)
        ow1 = (W + 6 - 7) // 2 + 1
        if (not self.avgpool_input and (self.split or not self.raw_fp32) and self.in_channels <= 32 and baseplanes == 32
                and 128 <= ow1 <= 256):
            # the stem kernels (conv_stem2.cu) work on 64-byte pixels: an 8-channel input (rgb + depth) is padded to 32
            # zero channels rather than sent through the generic cp.async producer (measured 2.5 ms vs 0.5 ms at B=256)
            self.cin_pad = 32
        self.convs, self.gns = {}, {}
        c1 = ConvLayer(pfx + ".conv1.0.weight", self.in_channels, baseplanes, 7, 7, 2, 3, H, W, need_dgrad=False,
                       cin_pad=self.cin_pad)
        self.conv1, self.gn1 = c1, GNLayer(pfx + ".conv1.1", baseplanes, ngroups)
        self.PH, self.PW = (c1.OH + 2 - 3) // 2 + 1, (c1.OW + 2 - 3) // 2 + 1
        h, w, inpl = self.PH, self.PW, baseplanes
        self.blocks = []
        exp = 1 if kind == "basic" else 4
        for li, nb in enumerate(nblocks, start=1):
            planes = baseplanes * (2 ** (li - 1))
            for b in range(nb):
                stride = 2 if (b == 0 and li > 1) else 1
                p = f"{pfx}.layer{li}.{b}"
                blk = {"name": p, "convs": [], "gns": [], "down": None, "IH": h, "IW": w, "Cin": inpl}
                if kind == "basic":
                    ca = ConvLayer(p + ".convs.0.weight", inpl, planes, 3, 3, stride, 1, h, w)
                    cb = ConvLayer(p + ".convs.3.weight", planes, planes, 3, 3, 1, 1, ca.OH, ca.OW)
                    blk["convs"] = [ca, cb]
                    blk["gns"] = [GNLayer(p + ".convs.1", planes, ngroups), GNLayer(p + ".convs.4", planes, ngroups)]
                else:
                    ca = ConvLayer(p + ".convs.0.weight", inpl, planes, 1, 1, 1, 0, h, w)
                    cb = ConvLayer(p + ".convs.3.weight", planes, planes, 3, 3, stride, 1, h, w)
                    cc = ConvLayer(p + ".convs.6.weight", planes, planes * exp, 1, 1, 1, 0, cb.OH, cb.OW)
                    blk["convs"] = [ca, cb, cc]
                    blk["gns"] = [GNLayer(p + ".convs.1", planes, ngroups), GNLayer(p + ".convs.4", planes, ngroups),
                                  GNLayer(p + ".convs.7", planes * exp, ngroups)]
                outc = planes * exp
                last = blk["convs"][-1]
                if b == 0 and (stride != 1 or inpl != outc):
                    blk["down"] = (ConvLayer(p + ".downsample.0.weight", inpl, outc, 1, 1, stride, 0, h, w),
                                   GNLayer(p + ".downsample.1", outc, ngroups))
                blk["OH"], blk["OW"], blk["Cout"] = last.OH, last.OW, outc
                self.blocks.append(blk)
                h, w, inpl = last.OH, last.OW, outc
        self.fH, self.fW, self.final_channels = h, w, inpl
        self.comp = ConvLayer(self.prefix + ".compression.0.weight", inpl, comp_ch, 3, 3, 1, 1, h, w)
        self.gnc = GNLayer(self.prefix + ".compression.1", comp_ch, 1)
        self.comp_ch = comp_ch
        if self.head is not None:
            emb = self.head.get("embed")
            src_ld = comp_ch * h * w + emb["dim"] if emb else 0
            self.fc = LinearLayer(self.head["fc_w"], comp_ch, h, w, self.comp.cout_pad, self.head["hidden"], src_ld=src_ld)

    def all_convs(self):
        out = [self.conv1]
        for blk in self.blocks:
            out += blk["convs"]
            if blk["down"]:
                out.append(blk["down"][0])
        out.append(self.comp)
        if self.head is not None:
            out.append(self.fc)
        return out

    def all_gns(self):
        out = [self.gn1]
        for blk in self.blocks:
            out += blk["gns"]
            if blk["down"]:
                out.append(blk["down"][1])
        out.append(self.gnc)
        return out

    # ------------------------------------------------------------------------------------------
    def _act(self, B, H, W, C, dtype=torch.float16):
        t = torch.empty(B, H, W, C, dtype=dtype, device=self.dev)
        if self.split and dtype == torch.float16:
            self._lo[t.data_ptr()] = torch.zeros_like(t)
        return t

    def lo(self, t):
        """Residual plane of an activation tensor in split-fp16 mode (None otherwise)."""
        return self._lo.get(t.data_ptr()) if (self.split and t is not None) else None

    def _alloc(self):
        B, dev, tr = self.B, self.dev, self.training
        for c in self.all_convs():
            c.alloc(dev, tr, own_dwp=False, split=self.split)
        gns = self.all_gns()
        # one contiguous fp32 region for all GroupNorm partial sums -> a single ZERO op per forward
        tot = sum(B * g.G * 2 for g in gns)
        # fp64 accumulators: the order of the epilogues' atomics then no longer shows in the fp32 mean / rstd
        self.stats_all = torch.zeros(_ru(tot, 4), dtype=torch.float64, device=dev)
        off = 0
        for g in gns:
            g.stats = self.stats_all[off:off + B * g.G * 2]
            off += B * g.G * 2
        if tr:
            # one fp32 arena for everything the backward pass accumulates into: packed dW of every conv + the
            # GroupNorm per-(sample, channel) sums -> ONE zero-fill launch per step
            tot = sum(B * g.C * 2 for g in gns)
            n_dw = sum(c.cout_pad * c.w_ld for c in self.all_convs())
            self.bwd_arena = torch.zeros(_ru(tot, 4) + n_dw + 5 * 5 * 32, dtype=torch.float32, device=dev)
            self.stem_S = self.bwd_arena[_ru(tot, 4) + n_dw:]  # border-class sums of the stem's output gradient
            self.sums_all = self.bwd_arena[:_ru(tot, 4)]
            off = 0
            for g in gns:
                g.sums = self.sums_all[off:off + B * g.C * 2]
                off += B * g.C * 2
            off = _ru(tot, 4)
            for c in self.all_convs():
                n = c.cout_pad * c.w_ld
                c.dwp = self.bwd_arena[off:off + n].view(c.cout_pad, c.w_ld)
                off += n
        raw_dt = torch.float32 if self.raw_fp32 else torch.float16
        # stem kernel (conv_stem.cu): needs 64-byte pixels, 32 output channels and W-padded rows with a zero halo
        self.use_stem = (not self.avgpool_input and (self.split or not self.raw_fp32) and self.cin_pad == 32
                         and self.conv1.cout_pad == 32 and self.conv1.Cout == 32 and self.conv1.OW <= 256
                         and self.conv1.OW >= 128)
        if self.use_stem and self.split:  # only the pixels-as-N stem kernel has a split variant
            self.use_stem = bool(self.stem_version >= 2 and L.load().pnvo_conv_stem2_supported(self.inH, self.inW)
                                 and (not tr or L.load().pnvo_conv_stem_wgrad2_supported(self.inH, self.inW)))
        self.exact_stem = bool(self.exact_stem and self.use_stem and self.conv1.OH >= 4 and self.conv1.OW >= 4)
        if self.use_stem:
            self.x0_pitch = L.load().pnvo_stem_padded_width(self.inW)
            self.x0 = torch.zeros(B, self.inH, self.x0_pitch, self.cin_pad, dtype=torch.float16, device=dev)
            if self.split and not self.exact_stem:
                self._lo[self.x0.data_ptr()] = torch.zeros_like(self.x0)
            if self.exact_stem:
                self.xp = torch.zeros(6 * 32, dtype=torch.float32, device=dev)       # per-channel constants (stem_exact.cu)
                self.bias5 = torch.zeros(5 * 5 * 32, dtype=torch.float32, device=dev)
            self.x0_img = self.x0[:, :, 3:, :]  # view whose data_ptr is the first image pixel
            self.w_stem = torch.zeros(7 * 4 * 32, 64, dtype=torch.float16, device=dev)
            # pixels-as-N stem kernel (conv_stem2.cu): full-rate MMAs, resident weights
            self.use_stem2 = bool(self.stem_version >= 2 and L.load().pnvo_conv_stem2_supported(self.inH, self.inW))
            self.w_stem2 = torch.zeros(4 * 7 * 32, 64, dtype=torch.float16, device=dev) if self.use_stem2 else None
            if self.split:
                # residual plane of the stem weights + the fp16 tensor that carries the w_lo * x product between the two
                # launches of the split stem (value + residual weights do not fit in shared memory together)
                self.w_stem2_lo = torch.zeros(4 * 7 * 32, 64, dtype=torch.float16, device=dev)
                self.stem_corr = torch.empty(B, self.conv1.OH, self.conv1.OW, 32, dtype=torch.float16, device=dev)
        else:
            self.x0_pitch = 0
            self.x0 = self._act(B, self.inH, self.inW, self.cin_pad)
            self.x0_img = self.x0
        self.x0c = None
        if self.avgpool_input:
            self.x0.zero_()  # pad channels stay zero (avgpool writes only the real ones)
            # one-channel input (the depth-only policy) without input normalisation: the direct fp32 stem kernels read the
            # pooled image as a compact fp32 plane (4 instead of 2 x 16 bytes per pixel); x0 itself is then never written
            if self.in_channels == 1 and self._compact_input:
                self.x0c = torch.zeros(B, self.inH, self.inW, dtype=torch.float32, device=dev)
        self.in_stats = torch.zeros(2 * 32 + 2, dtype=torch.float64, device=dev)
        self.drop_seed = torch.tensor([torch.initial_seed() & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device=dev)
        self.in_scale = torch.ones(32, dtype=torch.float32, device=dev)
        self.in_shift = torch.zeros(32, dtype=torch.float32, device=dev)
        c1 = self.conv1
        self.raw1 = self._act(B, c1.OH, c1.OW, c1.cout_pad, raw_dt)
        self.pool = self._act(B, self.PH, self.PW, c1.cout_pad)
        self.argmax = torch.empty(B, self.PH, self.PW, c1.cout_pad, dtype=torch.uint8, device=dev) if tr else None
        for blk in self.blocks:
            blk["raw"] = [self._act(B, c.OH, c.OW, c.cout_pad, raw_dt) for c in blk["convs"]]
            blk["mid"] = [self._act(B, c.OH, c.OW, c.cout_pad) for c in blk["convs"][:-1]]
            blk["y"] = self._act(B, blk["OH"], blk["OW"], blk["convs"][-1].cout_pad)
            if blk["down"]:
                d = blk["down"][0]
                blk["raw_d"] = self._act(B, d.OH, d.OW, d.cout_pad, raw_dt)
                blk["res_d"] = self._act(B, d.OH, d.OW, d.cout_pad)
        self.raw_c = self._act(B, self.fH, self.fW, self.comp.cout_pad, raw_dt)
        self.feat = self._act(B, self.fH, self.fW, self.comp.cout_pad)
        if self.head is not None:
            hid, od = self.head["hidden"], self.head.get("out_dim")
            assert hid == self.fc.cout_pad, "hidden size must be a multiple of 32"
            self.z = torch.empty(B, self.fc.cout_pad, dtype=torch.float32, device=dev)
            self.h32 = torch.empty(B, hid, dtype=torch.float32, device=dev)
            self.h16 = torch.empty(B, hid, dtype=torch.float16, device=dev)
            # with an output layer the plan's result is [B, out_dim]; without, the hidden features [B, hidden]
            self.out = torch.empty(B, od, dtype=torch.float32, device=dev) if od else self.h32
            emb = self.head.get("embed")
            if emb:  # action-embedding inputs of the hidden layer (vo_cnn_act_embed.py:65-75)
                self.actions = torch.zeros(B, dtype=torch.int64, device=dev)
                self.e_used = torch.zeros(B, emb["dim"], dtype=torch.float32, device=dev)
                self.e_mask = torch.ones(B, emb["dim"], dtype=torch.float32, device=dev)
        if tr:
            # gradient buffers (w.r.t. post-activation tensors, fp16) and raw-gradient scratch
            self.g_feat = torch.empty_like(self.feat)
            self.dx_c = torch.empty_like(self.feat)
            for blk in self.blocks:
                blk["g_y"] = torch.empty_like(blk["y"])
                blk["dy_last"] = torch.empty_like(blk["y"])
                blk["dx"] = [torch.empty(r.shape, dtype=torch.float16, device=dev) for r in blk["raw"]]
                blk["g_mid"] = [torch.empty_like(m) for m in blk["mid"]]
                if blk["down"]:
                    blk["dx_d"] = torch.empty(blk["raw_d"].shape, dtype=torch.float16, device=dev)
            self.g_pool = torch.empty_like(self.pool)
            self.dy1 = torch.empty(self.raw1.shape, dtype=torch.float16, device=dev)
            self.dx1 = torch.empty(self.raw1.shape, dtype=torch.float16, device=dev)
            if self.head is not None:
                od = self.head.get("out_dim")
                self.dout = torch.zeros(B, od if od else self.head["hidden"], dtype=torch.float32, device=dev)
                self.dz16 = torch.zeros(B, self.fc.cout_pad, dtype=torch.float16, device=dev)
            else:
                self.g_feat_in = self.g_feat
            # flat gradient bucket: one fp32 buffer, per-parameter views (the DDP all-reduce payload)
            names = self.param_names()
            n = sum(self.P[k].numel() for k in names)
            grad_bucket = self._grad_bucket_arg
            if grad_bucket is not None:
                # caller-provided storage (parallel_utils.PeerBuckets: a peer-mapped region, so that the other ranks' fused
                # reduce-scatter + Adam kernels can read this rank's gradients in place)
                if grad_bucket.numel() < n or grad_bucket.dtype != torch.float32 or not grad_bucket.is_cuda:
                    raise L.PnvoError(f"grad_bucket must be an fp32 CUDA tensor of at least {n} elements")
                self.grad_flat = grad_bucket[:n]
                self.grad_flat.zero_()
            else:
                self.grad_flat = torch.zeros(n, dtype=torch.float32, device=dev)
            off = 0
            for k in names:
                m = self.P[k].numel()
                self.grads[k] = self.grad_flat[off:off + m].view(self.P[k].shape)
                off += m

    def param_names(self):
        names = []
        for c in self.all_convs():
            names.append(c.key)
        for g in self.all_gns():
            names += [g.key + ".weight", g.key + ".bias"]
        if self.head is not None:
            names += [self.head["fc_b"]]
            if self.head.get("embed"):
                names += [self.head["embed"]["table"]]
            if self.head.get("out_dim"):
                names += [self.head["out_w"], self.head["out_b"]]
        return names

    # ------------------------------------------------------------------------------------------
    def _gn_apply(self, g, x, y, HW, relu=True, res=None):
        B = self.B
        return L.op_gn_apply(x, g.stats, self.P[g.key + ".weight"], self.P[g.key + ".bias"], y, B, g.C, g.G, g.cpg, HW,
                             float(g.cpg_real * HW), relu, res, self.raw_fp32, 1e-5, g.C_real, y_lo=self.lo(y),
                             res_lo=self.lo(res), x_lo=self.lo(x))

    def _gn_bwd(self, reduce, g, gin, relu_ref, x, dx, dy_out, HW, g_scale=1.0, class_sums=None, ohw=(0, 0)):
        return L.op_gn_bwd(reduce, gin, relu_ref, x, g.stats, self.P[g.key + ".weight"], g.sums, dx, dy_out, self.B,
                           g.C, g.G, g.cpg, HW, float(g.cpg_real * HW), self.raw_fp32, 1e-5, g.C_real, g_scale,
                           class_sums=class_sums, ohw=ohw)

    def _gn_bwd_all(self, ops, g, gin, relu_ref, x, dx, dy_out, HW, g_scale=1.0, class_sums=None, ohw=(0, 0)):
        assert class_sums is None or not (self.fuse_gn_bwd and L.load().pnvo_gn_bwd_fused_supported(g.C, HW, int(self.raw_fp32)))
        if self.fuse_gn_bwd and L.load().pnvo_gn_bwd_fused_supported(g.C, HW, int(self.raw_fp32)):
            # one pass: a cluster per sample keeps g / x in registers between the reduction and the apply
            ops.append(self._gn_bwd("fused", g, gin, relu_ref, x, dx, dy_out, HW, g_scale))
            if not self.batch_small_ops:
                ops.append(L.op_gn_param_grad(g.sums, self.grads[g.key + ".weight"], self.grads[g.key + ".bias"], self.B,
                                              g.C, g.C_real))
            return
        ops.append(self._gn_bwd(True, g, gin, relu_ref, x, dx, dy_out, HW, g_scale))
        if not self.batch_small_ops:
            ops.append(L.op_gn_param_grad(g.sums, self.grads[g.key + ".weight"], self.grads[g.key + ".bias"], self.B, g.C,
                                          g.C_real))
        ops.append(self._gn_bwd(False, g, gin, relu_ref, x, dx, dy_out, HW, g_scale, class_sums=class_sums, ohw=ohw))

    def _build_programs(self):
        B = self.B
        # ---- pack ----
        if self.batch_small_ops:
            self._pack_tab = L.device_table([c.pack_desc(self.P[c.key]) for c in self.all_convs()], self.dev)
            pack = [L.op_multi(L.OP_PACK_W_MULTI, self._pack_tab, len(self.all_convs()))]
            if self.split:
                self._pack_tab_lo = L.device_table([c.pack_desc_lo(self.P[c.key]) for c in self.all_convs()], self.dev)
                pack.append(L.op_multi(L.OP_PACK_W_MULTI, self._pack_tab_lo, len(self.all_convs())))
        else:
            pack = [c.op_pack(self.P[c.key]) for c in self.all_convs()]
        if self.use_stem and not self.exact_stem:  # (the exact-input stem packs a_c * W inside the forward program)
            if not self.split:
                pack.append(L.op_pack_w_stem(self.P[self.conv1.key], self.w_stem, self.conv1.Cin))
            if self.use_stem2:
                pack.append(L.op_pack_w_stem2(self.P[self.conv1.key], self.w_stem2, self.conv1.Cin))
                if self.split:
                    pack.append(L.op_pack_w_stem2(self.P[self.conv1.key], self.w_stem2_lo, self.conv1.Cin, lo=True))
        self.pack_prog = L.Program(pack, graph=True)
        # ---- forward (after the input tensor x0 has been produced) ----
        ops = [L.op_zero(self.stats_all)]
        c1, g1 = self.conv1, self.gn1
        if self.exact_stem:
            # exact raw values in x0, normalisation folded into W' = a_c * W (packed here: a_c follows the running
            # statistics) and the border bias; W'_lo * x as an fp16 tensor first, then W' * x + that tensor + bias
            ops.append(L.op_stem_exact_pack(self.P[c1.key], self.xp, self.w_stem2, self.w_stem2_lo, self.bias5, c1.Cin,
                                            self.inH, self.inW))
            corr = None
            if STEM_WLO:
                corr = self.stem_corr
                ops.append(L.op_conv_stem2(self.x0, self.w_stem2_lo, corr, None, B, self.inH, self.inW, g1.G, g1.cpg))
            ops.append(L.op_conv_stem2(self.x0, self.w_stem2, self.raw1, g1.stats, B, self.inH, self.inW, g1.G, g1.cpg,
                                       add=corr, bias5=self.bias5, y_lo=self.lo(self.raw1)))
        elif self.use_stem2 and self.split:
            # raw1 = w * (x + x_lo) + (w_lo * x): the residual-weight product first, as an fp16 tensor (it is ~2^-11 of
            # the result), then the value weights against both input planes with that tensor added in the epilogue
            ops.append(L.op_conv_stem2(self.x0, self.w_stem2_lo, self.stem_corr, None, B, self.inH, self.inW, g1.G, g1.cpg))
            ops.append(L.op_conv_stem2(self.x0, self.w_stem2, self.raw1, g1.stats, B, self.inH, self.inW, g1.G, g1.cpg,
                                       x_lo=self.lo(self.x0), add=self.stem_corr, y_lo=self.lo(self.raw1)))
        elif self.use_stem2 and not self.raw_fp32:
            ops.append(L.op_conv_stem2(self.x0, self.w_stem2, self.raw1, g1.stats, B, self.inH, self.inW, g1.G, g1.cpg))
        elif self.use_stem and not self.raw_fp32:
            ops.append(L.op_conv_stem(self.x0, self.w_stem, self.raw1, g1.stats, B, self.inH, self.inW, g1.G, g1.cpg, 2))
        else:
            assert not self.use_stem, "raw_fp32 is not supported together with the stem kernel"
            ops.append(c1.op_fwd(self.x0, self.raw1, B, g1.stats, g1.cpg, g1.G, self.raw_fp32, x_lo=self.lo(self.x0),
                                 y_lo=self.lo(self.raw1), x_c=self.x0c))
        ops.append(L.op_gn_pool(self.raw1, g1.stats, self.P[g1.key + ".weight"], self.P[g1.key + ".bias"], self.pool,
                                self.argmax, B, g1.C, g1.G, g1.cpg, c1.OH, c1.OW, self.PH, self.PW,
                                float(g1.cpg_real * c1.OH * c1.OW), self.raw_fp32, 1e-5, g1.C_real,
                                y_lo=self.lo(self.pool), x_lo=self.lo(self.raw1)))
        x = self.pool
        for blk in self.blocks:
            convs, gns = blk["convs"], blk["gns"]
            res = x
            if blk["down"]:
                d, gd = blk["down"]
                ops.append(d.op_fwd(x, blk["raw_d"], B, gd.stats, gd.cpg, gd.G, self.raw_fp32, x_lo=self.lo(x),
                                    y_lo=self.lo(blk["raw_d"])))
                ops.append(self._gn_apply(gd, blk["raw_d"], blk["res_d"], d.OH * d.OW, relu=False))
                res = blk["res_d"]
            cur = x
            for k, (c, g) in enumerate(zip(convs, gns)):
                ops.append(c.op_fwd(cur, blk["raw"][k], B, g.stats, g.cpg, g.G, self.raw_fp32, x_lo=self.lo(cur),
                                    y_lo=self.lo(blk["raw"][k])))
                if k < len(convs) - 1:
                    ops.append(self._gn_apply(g, blk["raw"][k], blk["mid"][k], c.OH * c.OW, relu=True))
                    cur = blk["mid"][k]
                else:
                    ops.append(self._gn_apply(g, blk["raw"][k], blk["y"], c.OH * c.OW, relu=True, res=res))
            blk["x_in"] = x
            x = blk["y"]
        cc, gc = self.comp, self.gnc
        ops.append(cc.op_fwd(x, self.raw_c, B, gc.stats, gc.cpg, gc.G, self.raw_fp32, x_lo=self.lo(x),
                             y_lo=self.lo(self.raw_c)))
        ops.append(self._gn_apply(gc, self.raw_c, self.feat, self.fH * self.fW, relu=True))
        p_drop = self.dropout_p
        if p_drop > 0:  # nn.Dropout in front of visual_fc (vo_cnn.py:218)
            ops.append(L.op_dropout(self.feat, self.drop_seed, 0, p_drop))
            if self.split:  # same counter-based mask on the residual plane
                ops.append(L.op_dropout(self.lo(self.feat), self.drop_seed, 0, p_drop))
        if self.head is not None:
            hd = self.head
            feat_flat = self.feat  # [B, 1, 1, fH*fW*c_pad] as far as the 1x1 "conv" is concerned
            ops.append(self.fc.op_fwd(feat_flat, self.z, B, None, 0, 0, True, x_lo=self.lo(self.feat)))
            emb = hd.get("embed")
            if emb:
                E = self.P[emb["table"]]
                ops.append(L.op_act_embed_fwd(self.z, self.P[hd["fc_w"]], E, self.actions, self.e_used, self.e_mask,
                                              self.drop_seed, B, hd["hidden"], emb["dim"], E.shape[0], self.fc.src_ld,
                                              self.fc.src_ld - emb["dim"], p_drop))
            ops.append(L.op_bias_relu(self.z, self.P[hd["fc_b"]], self.h32, self.h16, B, hd["hidden"], True))
            if p_drop > 0:  # nn.Dropout in front of output_head (vo_cnn.py:224)
                ops.append(L.op_dropout(self.h32, self.drop_seed, 1, p_drop, advance=True))
            if hd.get("out_dim"):
                ops.append(L.op_head_fwd(self.h32, self.P[hd["out_w"]], self.P[hd["out_b"]], self.out, B, hd["hidden"],
                                         hd["out_dim"]))
        self.fwd_ops = ops
        self.fwd_prog = L.Program(ops, graph=True)
        if not self.training:
            return
        # ---- backward ----
        # Weight gradients only feed the gradient unpack at the end: in the captured graph they run on a side lane, under
        # the data-gradient / GroupNorm-backward chain that follows them (their inputs -- forward activations and the dx
        # buffers -- are written once per step)
        W = L.side if self.side_lane_wgrad else (lambda op: op)
        if os.environ.get("PNVO_DIAG_SKIP_WGRAD") == "1":   # timing diagnosis only: no weight gradients at all
            W = lambda op: None  # noqa: E731
        ops = [L.op_zero(self.bwd_arena)]
        if self.head is not None:
            hd = self.head
            if hd.get("out_dim"):
                ops.append(L.op_head_bwd(self.dout, self.h32, self.P[hd["out_w"]], self.grads[hd["out_w"]],
                                         self.grads[hd["out_b"]], self.dz16, self.grads[hd["fc_b"]], B, hd["hidden"],
                                         hd["out_dim"], False, 1.0 / (1.0 - self.dropout_p)))
            else:  # gradient arrives w.r.t. the hidden features
                ops.append(L.op_bias_relu_bwd(self.dout, self.h32, self.dz16, self.grads[hd["fc_b"]], B, hd["hidden"]))
            emb = hd.get("embed")
            if emb:
                E = self.P[emb["table"]]
                ops.append(L.op_zero(self.grads[emb["table"]]))
                ops.append(L.op_act_embed_bwd(self.dz16, self.P[hd["fc_w"]], E, self.actions, self.e_used, self.e_mask,
                                              self.grads[hd["fc_w"]], self.grads[emb["table"]], B, hd["hidden"],
                                              emb["dim"], E.shape[0], self.fc.src_ld, self.fc.src_ld - emb["dim"],
                                              self.fc.cout_pad))
            ops.append(W(self.fc.op_wgrad(self.feat, self.dz16, B)))
            ops.append(self.fc.op_dgrad(self.dz16, self.g_feat, B))
        HWf = self.fH * self.fW
        self._gn_bwd_all(ops, gc, self.g_feat, self.feat, self.raw_c, self.dx_c, None, HWf,
                         g_scale=1.0 / (1.0 - self.dropout_p))
        ops.append(W(cc.op_wgrad(x, self.dx_c, B)))
        ops.append(cc.op_dgrad(self.dx_c, self.blocks[-1]["g_y"], B))
        for bi in range(len(self.blocks) - 1, -1, -1):
            blk = self.blocks[bi]
            convs, gns = blk["convs"], blk["gns"]
            g_in = blk["g_y"]                                  # gradient w.r.t. the block output
            g_x = self.blocks[bi - 1]["g_y"] if bi > 0 else self.g_pool  # gradient w.r.t. the block input
            n = len(convs)
            # last GN (+ residual + ReLU): relu mask from the saved block output
            c, g = convs[-1], gns[-1]
            self._gn_bwd_all(ops, g, g_in, blk["y"], blk["raw"][-1], blk["dx"][-1], blk["dy_last"], c.OH * c.OW)
            add = blk["dy_last"]  # identity-branch gradient
            if blk["down"]:
                d, gd = blk["down"]
                self._gn_bwd_all(ops, gd, blk["dy_last"], None, blk["raw_d"], blk["dx_d"], None, d.OH * d.OW)
                ops.append(W(d.op_wgrad(blk["x_in"], blk["dx_d"], B)))
                ops += d.ops_dgrad(blk["dx_d"], g_x, B)
                add = g_x
            for k in range(n - 1, -1, -1):
                c, g = convs[k], gns[k]
                xin = blk["x_in"] if k == 0 else blk["mid"][k - 1]
                ops.append(W(c.op_wgrad(xin, blk["dx"][k], B)))
                if k > 0:
                    ops += c.ops_dgrad(blk["dx"][k], blk["g_mid"][k - 1], B)
                    cp, gp = convs[k - 1], gns[k - 1]
                    self._gn_bwd_all(ops, gp, blk["g_mid"][k - 1], blk["mid"][k - 1], blk["raw"][k - 1],
                                     blk["dx"][k - 1], None, cp.OH * cp.OW)
                else:
                    ops += c.ops_dgrad(blk["dx"][0], g_x, B, add=add)
        # stem: max-pool + ReLU routing, GN, conv1 weight gradient (no data gradient: the input is data)
        # (gathering the max-pool backward inside the two GroupNorm passes was measured slower -- 0.85 vs 0.62 ms --
        # than materialising dy1 once: the 4-window gather runs twice)
        ops.append(L.op_pool_bwd(self.g_pool, self.pool, self.argmax, self.dy1, B, g1.C, c1.OH, c1.OW, self.PH, self.PW))
        # exact-input stem: the border-class sums of dx1 (for the conv1 weight gradient) are accumulated by the apply pass of
        # the GroupNorm backward while it writes dx1 -- unless a sample fits the one-pass cluster kernel (small inputs)
        fuse_S = bool(self.exact_stem and c1.cout_pad == 32 and os.environ.get("PNVO_FUSED_DY_SUMS", "1") != "0"
                      and not (self.fuse_gn_bwd and L.load().pnvo_gn_bwd_fused_supported(g1.C, c1.OH * c1.OW, int(self.raw_fp32))))
        self._gn_bwd_all(ops, g1, self.dy1, None, self.raw1, self.dx1, None, c1.OH * c1.OW,
                         class_sums=self.stem_S if fuse_S else None, ohw=(c1.OH, c1.OW))
        if self.use_stem and self.stem_version >= 2 and L.load().pnvo_conv_stem_wgrad2_supported(self.inH, self.inW):
            ops.append(W(L.op_wgrad_stem2(self.x0, self.dx1, c1.dwp, B, self.inH, self.inW, c1.w_ld)))
        elif self.use_stem and 96 < c1.OW <= 176:
            ops.append(W(L.op_wgrad_stem(self.x0, self.dx1, c1.dwp, B, self.inH, self.inW, c1.w_ld, 48)))
        else:
            ops.append(W(c1.op_wgrad(self.x0_img, self.dx1, B, x_row_pitch=self.x0_pitch, x_c=self.x0c)))
        if self.exact_stem and not fuse_S:
            ops.append(W(L.op_stem_dy_sums(self.dx1, self.stem_S, B, c1.OH, c1.OW)))
        ops.append(L.op_join())
        if self.exact_stem:
            ops.append(L.op_stem_exact_unpack(c1.dwp, self.stem_S, self.xp, self.grads[c1.key], c1.w_ld, c1.Cin, self.inH,
                                              self.inW))
        if self.batch_small_ops:
            tab_convs = [c for c in self.all_convs() if not (self.exact_stem and c is c1)]
            self._unpack_tab = L.device_table([c.unpack_desc(self.grads[c.key]) for c in tab_convs], self.dev)
            ops.append(L.op_multi(L.OP_UNPACK_DW_MULTI, self._unpack_tab, len(tab_convs)))
            gns = self.all_gns()
            self._gnp_tab = L.device_table(
                [L.GnParamDesc(g.sums.data_ptr(), self.grads[g.key + ".weight"].data_ptr(),
                               self.grads[g.key + ".bias"].data_ptr(), g.C, g.C_real) for g in gns], self.dev)
            ops.append(L.op_multi(L.OP_GN_PARAM_GRAD_MULTI, self._gnp_tab, len(gns), B))
        else:
            for c in self.all_convs():
                if not (self.exact_stem and c is c1):
                    ops.append(c.op_unpack(self.grads[c.key]))
        ops = [o for o in ops if o is not None]
        self.bwd_ops = ops
        self.bwd_prog = L.Program(ops, graph=True)

    # ------------------------------------------------------------------------------------------
    # ------------------------------------------------------------------------------------------
    # Double-buffered input staging: programs_for(1) replays the forward / backward programs over a SECOND assembled
    # input tensor x0, so the preprocessing of batch i+1 (top-down, statistics, assembly: HBM-bound, 14 % of a step)
    # can run on a side stream while the tensor-core kernels of step i read the first one.
    def x0_for(self, parity):
        if parity == 0:
            return self.x0
        if getattr(self, "x0_alt", None) is None:
            self.x0_alt = torch.zeros_like(self.x0)
            if self.split and not self.exact_stem:
                self._lo[self.x0_alt.data_ptr()] = torch.zeros_like(self.x0)
        return self.x0_alt

    def xp_for(self, parity):
        """Per-channel constants of the exact-input stem that belong to staging buffer `parity`."""
        if parity == 0:
            return self.xp
        if getattr(self, "xp_alt", None) is None:
            self.xp_alt = torch.zeros_like(self.xp)
        return self.xp_alt

    def programs_for(self, parity):
        if parity == 0:
            return self.fwd_prog, (self.bwd_prog if self.training else None)
        if getattr(self, "_alt_progs", None) is None:
            alt = self.x0_for(1)
            off = self.x0_img.data_ptr() - self.x0.data_ptr()
            mapping = {self.x0.data_ptr(): alt.data_ptr(), self.x0.data_ptr() + off: alt.data_ptr() + off}
            if self.split and not self.exact_stem:
                mapping[self.lo(self.x0).data_ptr()] = self.lo(alt).data_ptr()
            if self.exact_stem:
                mapping[self.xp.data_ptr()] = self.xp_for(1).data_ptr()
            fwd = L.Program(L.patch_ops(self.fwd_ops, mapping), graph=True)
            bwd = L.Program(L.patch_ops(self.bwd_ops, mapping), graph=True) if self.training else None
            self._alt_progs = (fwd, bwd)
        return self._alt_progs

    def input_ops(self, obs, out32=None):
        """avg-pool input path of the RL encoder (resnet_policy.py:146-168): sources are pooled 2x2 and written
        into the channel slots of x0 (pad channels stay zero) -- or, with out32, into an fp32 [B, h, w, C] tensor that
        the RunningMeanAndVar ops then normalise into x0."""
        ops, coff = [], 0
        keep = []
        for k, n, scale in self.sources:
            t = obs[k]
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            assert t.shape[-1] == n
            if out32 is not None:
                ops.append(L.op_avgpool2(t, None, self.B, self.H, self.W, n, self.cin_pad, coff, scale, out32=out32,
                                         ld32=self.in_channels))
            else:
                if self.x0c is not None:
                    ops.append(L.op_avgpool2(t, None, self.B, self.H, self.W, n, self.cin_pad, coff, scale, out32=self.x0c,
                                             ld32=1))
                else:
                    ops.append(L.op_avgpool2(t, self.x0, self.B, self.H, self.W, n, self.cin_pad, coff, scale,
                                             out_lo=self.lo(self.x0)))
            coff += n
            keep.append(t)
        self._keepalive = keep
        return ops

    def conv_flops(self, backward=False):
        f = sum(c.flops(self.B) for c in self.all_convs())
        if self.head is not None and self.head.get("out_dim"):
            f += 2.0 * self.B * self.head["hidden"] * self.head["out_dim"]
        if backward:
            f = 3 * f - self.conv1.flops(self.B)
        return f
