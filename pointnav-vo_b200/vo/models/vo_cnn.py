"""Visual-odometry regression networks with the reference's nn.Module surface
(pointnav_vo/vo/models/vo_cnn.py:16-561): same constructor keywords, registry names, state_dict keys
and forward signature -- `model(observation_pairs: dict[str, NHWC fp32 tensor]) -> [B, output_dim]`.

The arithmetic runs in libpnvo (hand-written sm_100a kernels) through an `EncoderPlan` op program;
this file only owns parameters, the autograd boundary and the plan cache.  CUDA only: a CPU tensor or
a missing library raises (no fallback).
"""
import numpy as np
import torch
import torch.nn as nn

from ... import lib as L
from ...engine import EncoderPlan
from ...model_utils import resnet
from ...model_utils.running_mean_and_var import RunningMeanAndVar
from ...utils.baseline_registry import baseline_registry
from ..common.common_vars import (DEFAULT_DELTA_STATE_SIZE, DEPTH_PAIR_CHANNEL, RGB_PAIR_CHANNEL,
                                  TOP_DOWN_VIEW_PAIR_CHANNEL)

OBS_ORDER = ("rgb", "depth", "discretized_depth", "top_down_view")  # vo_cnn.py:114-166 append order



# timing diagnosis only (results become stale): bit 0 skips the top-down projection, bit 1 the batch statistics, bit 2 the
# input assembly of every batch after the first -- separates the cost of the overlapped input pipeline from the step's
import os as _os

_DIAG_SKIP = int(_os.environ.get("PNVO_DIAG_SKIP_INPUT", "0"))
_FUSED_STATS = _os.environ.get("PNVO_FUSED_STATS", "1") != "0"   # 0: separate raw_stats pass (A/B measurements)
_PEER_SUM = {}   # device index -> parallel_utils.PeerSmallSum, or False when peer memory is unavailable


def _allreduce_stats(t):
    """Packed RunningMeanAndVar batch statistics (sum, sum of squares per channel, fp64) summed over the ranks: ONE
    exchange instead of the reference's three all-reduces (running_mean_and_var.py:28-38).  On one NVLink node it is
    libpnvo's peer-memory kernel (parallel_utils.PeerSmallSum: a tiny CTA that co-resides with the persistent convolution
    kernels; an NCCL kernel waiting for the slowest rank blocks their placement, 0.21 ms per step measured); otherwise
    (gloo / other backends, PNVO_PEER_STATS=0, mapping failure) torch.distributed.all_reduce.
    PNVO_DIAG_LOCAL_STATS=1 skips the exchange (timing diagnosis only: the replicas' running statistics diverge)."""
    import os

    if os.environ.get("PNVO_DIAG_LOCAL_STATS") == "1":
        return
    key = t.device.index
    peer = _PEER_SUM.get(key)
    if peer is None:
        peer = False
        if (t.is_cuda and t.dtype == torch.float64 and os.environ.get("PNVO_PEER_STATS", "1") != "0"
                and torch.distributed.get_backend() == "nccl" and 2 <= torch.distributed.get_world_size() <= 8):
            from ...parallel_utils import PeerSmallSum

            try:
                peer = PeerSmallSum(t.device)
            except L.PnvoError as e:
                import warnings

                warnings.warn(f"peer-memory statistics exchange unavailable ({e}); using the NCCL all-reduce")
        _PEER_SUM[key] = peer
    if peer:
        peer.sum_(t)
    else:
        torch.distributed.all_reduce(t)


class Flatten(nn.Module):
    def forward(self, x):
        return x.contiguous().view(x.size(0), -1)


class ResNetEncoder(nn.Module):
    """vo_cnn.py:16-179.  Holds the backbone / compression parameters and the input bookkeeping."""

    def __init__(self, *, observation_space, observation_size, baseplanes=32, ngroups=32, spatial_size_w=128,
                 spatial_size_h=128, make_backbone=None, normalize_visual_inputs=False,
                 after_compression_flat_size=2048, rgb_pair_channel=RGB_PAIR_CHANNEL,
                 depth_pair_channel=DEPTH_PAIR_CHANNEL, discretized_depth_channels=0,
                 top_down_view_pair_channel=TOP_DOWN_VIEW_PAIR_CHANNEL):
        super().__init__()
        n = {"rgb": rgb_pair_channel, "depth": depth_pair_channel,
             "discretized_depth": discretized_depth_channels * 2, "top_down_view": top_down_view_pair_channel}
        self._sources = []
        for k in OBS_ORDER:
            if k in observation_space:
                spatial_size_w, spatial_size_h = observation_size
                self._sources.append((k, n[k], 1.0 / 255.0 if k == "rgb" else 1.0))
        self._n_input_rgb = n["rgb"] if "rgb" in observation_space else 0
        self._n_input_depth = n["depth"] if "depth" in observation_space else 0
        self._n_input_discretized_depth = n["discretized_depth"] if "discretized_depth" in observation_space else 0
        self._n_input_top_down_view = n["top_down_view"] if "top_down_view" in observation_space else 0
        input_channels = sum(s[1] for s in self._sources)
        assert input_channels > 0  # visual odometry must not be blind (vo_cnn.py:67-68)
        self.input_channels = input_channels
        self.spatial_size = (spatial_size_h, spatial_size_w)
        self.baseplanes, self.ngroups = baseplanes, ngroups
        if normalize_visual_inputs:
            self.running_mean_and_var = RunningMeanAndVar(input_channels)
        else:
            self.running_mean_and_var = nn.Sequential()
        self.backbone = make_backbone(input_channels, baseplanes, ngroups)
        final_w = int(np.ceil(spatial_size_w * self.backbone.final_spatial_compress))
        final_h = int(np.ceil(spatial_size_h * self.backbone.final_spatial_compress))
        num_compression_channels = int(round(after_compression_flat_size / (final_w * final_h)))
        self.compression = nn.Sequential(
            nn.Conv2d(self.backbone.final_channels, num_compression_channels, kernel_size=3, padding=1, bias=False),
            nn.GroupNorm(1, num_compression_channels), nn.ReLU(True))
        self.output_shape = (num_compression_channels, final_h, final_w)

    def layer_init(self):
        for layer in self.modules():
            if isinstance(layer, (nn.Conv2d, nn.Linear)):
                nn.init.kaiming_normal_(layer.weight, nn.init.calculate_gain("relu"))
                if layer.bias is not None:
                    nn.init.constant_(layer.bias, val=0)

    def forward(self, observation_pairs):
        raise RuntimeError("ResNetEncoder is executed as part of its owner's op program (libpnvo)")


class _VOFunction(torch.autograd.Function):
    """Autograd boundary: forward/backward are libpnvo op programs; parameters enter as inputs so that
    torch optimisers, .grad accumulation and DDP-style hooks see ordinary leaf gradients."""

    @staticmethod
    def forward(ctx, model, obs, training, need_grad, *params):
        plan = model._plan_for(obs, need_grad, training)
        model._run_forward(plan, obs, training)
        # the plan's activation buffers are shared by every forward of this shape: remember which forward filled them
        plan.generation = getattr(plan, "generation", 0) + 1
        ctx.model, ctx.plan, ctx.generation = model, plan, plan.generation
        return plan.out.clone()

    @staticmethod
    def backward(ctx, grad_out):
        model, plan = ctx.model, ctx.plan
        if not plan.training:
            raise RuntimeError("backward through a plan built without gradient buffers")
        if plan.generation != ctx.generation:
            raise RuntimeError("backward of a VO forward whose activations were overwritten by a later forward of the same "
                               "shape (the op program keeps ONE set of activation buffers per batch shape): call "
                               "backward before the next forward, or give the second forward its own module copy")
        plan.dout.copy_(grad_out)
        plan.bwd_prog.run(plan.dev)
        flat = plan.grad_flat.clone()  # detach from the plan's reusable bucket
        grads, off = [], 0
        by_name = {}
        for k in plan.param_names():
            m = plan.P[k].numel()
            by_name[k] = flat[off:off + m].view(plan.P[k].shape)
            off += m
        for k in model._param_order:
            grads.append(by_name.get(k))
        return (None, None, None, None, *grads)


class VisualOdometryCNNBase(nn.Module):
    """vo_cnn.py:182-233."""

    precision = "split"  # see set_precision
    exact_stem = True    # split plans on raw uint8 / fp16 inputs use the exact-input stem (csrc/stem_exact.cu)

    def __init__(self, *, observation_space, observation_size, hidden_size=512, resnet_baseplanes=32,
                 backbone="resnet18", normalize_visual_inputs=False, output_dim=DEFAULT_DELTA_STATE_SIZE,
                 dropout_p=0.2, after_compression_flat_size=2048, rgb_pair_channel=RGB_PAIR_CHANNEL,
                 depth_pair_channel=DEPTH_PAIR_CHANNEL, discretized_depth_channels=0,
                 top_down_view_pair_channel=TOP_DOWN_VIEW_PAIR_CHANNEL):
        super().__init__()
        self.visual_encoder = ResNetEncoder(
            observation_space=observation_space, observation_size=observation_size, baseplanes=resnet_baseplanes,
            ngroups=resnet_baseplanes // 2, make_backbone=resnet.make_backbone(backbone),
            normalize_visual_inputs=normalize_visual_inputs, after_compression_flat_size=after_compression_flat_size,
            rgb_pair_channel=rgb_pair_channel, depth_pair_channel=depth_pair_channel,
            discretized_depth_channels=discretized_depth_channels,
            top_down_view_pair_channel=top_down_view_pair_channel)
        self.visual_fc = nn.Sequential(Flatten(), nn.Dropout(dropout_p),
                                       nn.Linear(int(np.prod(self.visual_encoder.output_shape)), hidden_size),
                                       nn.ReLU(True))
        self.output_head = nn.Sequential(nn.Dropout(dropout_p), nn.Linear(hidden_size, output_dim))
        nn.init.orthogonal_(self.output_head[1].weight)
        nn.init.constant_(self.output_head[1].bias, 0)
        self._backbone_name = backbone
        self._hidden_size, self._output_dim, self._dropout_p = hidden_size, output_dim, dropout_p
        self._plans = {}
        self._packed_version = None
        self._ptr_sig = None
        self._param_order = [k for k, _ in self.named_parameters()]
        self._fc_keys = dict(fc_w="visual_fc.2.weight", fc_b="visual_fc.2.bias", out_w="output_head.1.weight",
                             out_b="output_head.1.bias")
        self.raw_fp32 = False  # store pre-GroupNorm conv outputs in fp32 instead of fp16

    # ---------------------------------------------------------------- runtime
    def _tensors(self):
        P = {k: p.data for k, p in self.named_parameters()}
        Bf = {k: b for k, b in self.named_buffers()}
        return P, Bf

    def _signature(self):
        return tuple(p.data_ptr() for p in self.parameters()) + tuple(b.data_ptr() for b in self.buffers())

    def set_precision(self, mode):
        """'split' (default): every activation and weight is a pair of fp16 planes (value + residual) and each forward
        convolution issues three tensor-core products into one fp32 accumulator (engine.EncoderPlan(split=True)):
        outputs agree with the fp32 reference to ~1e-5 (the north-star bound is 1e-3).  Training forwards use it too;
        the backward pass then runs single-pass fp16 operands on activations that match the reference's, so ReLU masks
        agree and gradients stay within ~1e-2 relative L2 per tensor.
        'fp16': single-pass fp16 operands / fp16 stored activations, ~1/3 of the forward convolution time, outputs
        within ~5e-3 of the fp32 reference (outside the north-star bound: a throughput mode, not the parity mode)."""
        if mode not in ("fp16", "split"):
            raise ValueError(mode)
        self.precision = mode
        return self

    # ---------------------------------------------------------------- raw observation pairs
    # forward() also accepts the step's inputs in the form the simulator / dataset stores them:
    #   {"rgb": uint8 [B,H,W,6], "depth": fp32 [B,H,W,2]}   (prev | cur on the channel axis)
    # The discretised-depth and top-down channels are then derived on the device (csrc/raw_input.cu,
    # pnvo_topdown_project_strided) instead of being shipped as 22 extra fp32 channels per pixel
    # (vo/dataset/regression_geo_invariance_iter_dataset.py:237-267, base_trainer_with_vo.py:212-269).
    def set_raw_input_config(self, top_down_generator=None, discretized_depth_end_vals=None):
        """Geometry of the derived channels; defaults = configs/vo/vo_pointnav.yaml (VO.GEOMETRY, 10 bins)."""
        self._raw_td_gen = top_down_generator
        self._raw_dd_end_vals = discretized_depth_end_vals

    def _is_raw(self, obs):
        enc = self.visual_encoder
        rgb = obs.get("rgb")
        if rgb is not None and rgb.dtype == torch.uint8:
            return True
        return any(k not in obs for k, _, _ in enc._sources)

    def _first_tensor(self, obs):
        for k in ("rgb", "depth"):
            if k in obs:
                return obs[k]
        return obs[self.visual_encoder._sources[0][0]]

    def _plan_for(self, obs, need_grad, training=False):
        enc = self.visual_encoder
        first = self._first_tensor(obs) if self._is_raw(obs) else obs[enc._sources[0][0]]
        if not first.is_cuda:
            raise L.PnvoError("VO model inputs must be CUDA tensors: the B200 path has no CPU fallback")
        L.load()
        sig = self._signature()
        if sig != self._ptr_sig:  # parameters moved (.to(), load with assign, ...): plans hold raw pointers
            self._plans.clear()
            self._ptr_sig = sig
            self._packed_version = None
        B, H, W = first.shape[0], first.shape[1], first.shape[2]
        if obs.get("pair_map") is not None:  # device-side inverse-pair augmentation: one network row per map entry
            B = obs["pair_map"].numel()
        drop = self._dropout_p if training else 0.0  # nn.Dropout is the identity in eval mode
        split = self.precision == "split"
        # exact-input stem (engine.EncoderPlan(exact_stem=True)): raw uint8 rgb / fp16 depth pairs are exactly representable
        # in fp16, so split plans fed with them fold the normalisation into the stem weights instead of carrying a residual
        # plane of the assembled input (fp32 depth, or the reference's four fp32 tensors, keep the residual plane)
        exact = bool(split and self.exact_stem and self._is_raw(obs)
                     and (obs.get("depth") is None or obs["depth"].dtype == torch.float16))
        key = (B, H, W, bool(need_grad), str(first.device), self.raw_fp32, drop, split, exact)
        plan = self._plans.get(key)
        if plan is None:
            P, Bf = self._tensors()
            for p in P.values():
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise L.PnvoError("parameters must be contiguous fp32 CUDA tensors")
            world = 1
            rmv = enc.running_mean_and_var
            if isinstance(rmv, RunningMeanAndVar) and rmv._distributed:
                world = torch.distributed.get_world_size()
            head = dict(self._fc_keys, hidden=self._hidden_size, out_dim=self._output_dim)
            if getattr(self, "_embed", None):
                head["embed"] = dict(self._embed)
            plan = EncoderPlan(params=P, buffers=Bf, B=B, H=H, W=W, in_channels=enc.input_channels,
                               sources=enc._sources, backbone=self._backbone_name, baseplanes=enc.baseplanes,
                               ngroups=enc.ngroups, compression_channels=enc.output_shape[0],
                               prefix="visual_encoder", head=head, training=bool(need_grad), device=first.device,
                               world_size=world, raw_fp32=self.raw_fp32, dropout_p=drop, split=split, exact_stem=exact,
                               grad_bucket=getattr(self, "_grad_bucket", None) if need_grad else None)
            self._plans[key] = plan
        return plan

    def _weights_version(self):
        return sum(p._version for p in self.parameters())

    def _run_forward_raw(self, plan, obs, training, parity=0, prepare_only=False):
        """parity selects the input staging buffer (engine.EncoderPlan.x0_for); prepare_only stops after the input
        pipeline (top-down, statistics, assembly) so that it can run ahead of the backbone on a side stream."""
        from ...utils import geometry_utils as gu

        enc = self.visual_encoder
        dev = plan.dev
        keys = [k for k, _, _ in enc._sources]
        use_rgb, use_depth = "rgb" in keys, "depth" in keys
        n_dd = enc._n_input_discretized_depth // 2
        use_td = "top_down_view" in keys
        H, W, B = plan.H, plan.W, plan.B
        rgb = depth = td = edges = None
        # pair_map (int32 [B], entry = 2 * source pair + swap flag): the raw tensors hold B_src <= B source pairs and the
        # kernels expand / swap them on the fly (regression_geo_invariance_iter_dataset.py:342-420 done on the device)
        pair_map = obs.get("pair_map")
        B_src = B
        if pair_map is not None:
            if pair_map.dtype != torch.int32 or not pair_map.is_cuda or pair_map.numel() != B:
                raise L.PnvoError("pair_map must be an int32 CUDA tensor with one entry per network row")
            B_src = self._first_tensor(obs).shape[0]
        if use_rgb:
            rgb = obs["rgb"]
            if rgb.dtype != torch.uint8 or not rgb.is_contiguous() or tuple(rgb.shape) != (B_src, H, W, 6):
                raise L.PnvoError("raw rgb pairs must be contiguous uint8 [B, H, W, 6]")
        if use_depth or n_dd or use_td:
            depth = obs["depth"]
            if (depth.dtype not in (torch.float32, torch.float16) or not depth.is_contiguous()
                    or tuple(depth.shape) != (B_src, H, W, 2)):
                raise L.PnvoError("raw depth pairs must be contiguous fp32 (or fp16, the dataset's type) [B, H, W, 2]")
        if n_dd:
            end_vals = getattr(self, "_raw_dd_end_vals", None) or gu.discretize_end_vals(n_dd)
            edges = gu._edges(end_vals, dev)
        if use_td:
            td = obs.get("top_down_view")
            if td is None:
                gen = getattr(self, "_raw_td_gen", None)
                if gen is None:  # configs/vo/vo_pointnav.yaml: VO.GEOMETRY (hfov passed verbatim)
                    gen = self._raw_td_gen = gu.NormalizedDepth2TopDownViewHabitatTorch(0.1, 10.0, H, W, 70)
                if getattr(plan, "td_pair", None) is None or plan.td_pair.shape[0] != B_src:
                    plan.td_pair = torch.empty(B_src, H, W, 2, dtype=torch.float32, device=dev)
                if _DIAG_SKIP & 1 and getattr(plan, "_td_done", False):
                    td = plan.td_pair          # timing diagnosis only (PNVO_DIAG_SKIP_INPUT bit 0): stale top-down maps
                else:
                    td = gu.gen_top_down_view_pairs(gen, depth, out=plan.td_pair)
                    plan._td_done = True
        C = enc.input_channels
        n_pix = B * H * W
        rmv = enc.running_mean_and_var
        have_rmv = isinstance(rmv, RunningMeanAndVar)
        ops = []
        full30 = bool(use_rgb and use_depth and n_dd == 10 and use_td)
        # exact-input stem: the stored tensor does not depend on the statistics, so ONE kernel writes it and accumulates them
        fused_stats = bool(plan.exact_stem and have_rmv and training and full30 and pair_map is None and _FUSED_STATS)

        def assemble(scale, shift, n_lo, stats=None):
            return L.op_raw_assemble(rgb, depth, td, edges, use_rgb, use_depth, n_dd, use_td, C, plan.cin_pad, n_pix,
                                     scale, shift, plan.x0_for(parity), row_w=plan.W if plan.x0_pitch else 0,
                                     out_pitch=plan.x0_pitch, out_lo=plan.lo(plan.x0_for(parity)),
                                     pair_map=pair_map, hw=H * W, n_lo=n_lo, exact=bool(plan.exact_stem), stats_f64=stats)

        skip_asm = bool(_DIAG_SKIP & 4 and parity in getattr(plan, "_asm_done", set()))
        if not hasattr(plan, "_asm_done"):
            plan._asm_done = set()
        plan._asm_done.add(parity)
        if have_rmv:
            if training and not (_DIAG_SKIP & 2 and getattr(plan, "_stats_done", False)):
                plan._stats_done = True
                ops.append(L.op_zero(plan.in_stats))
                if fused_stats:
                    ops.append(assemble(None, None, 2, stats=plan.in_stats))
                else:
                    ops.append(L.op_raw_stats(rgb, depth, td, edges, use_rgb, use_depth, n_dd, use_td, C, plan.cin_pad, n_pix,
                                              plan.in_stats, pair_map=pair_map, hw=H * W))
                if plan.world_size > 1:
                    L.run_ops(ops, dev)
                    ops = []
                    _allreduce_stats(plan.in_stats)
            ops.append(L.op_rmv_update(plan.in_stats, rmv._mean, rmv._var, rmv._count, plan.in_scale, plan.in_shift, C,
                                       training, True, plan.B * plan.world_size, plan.H * plan.W))
            scale, shift = plan.in_scale, plan.in_shift
        else:
            scale = shift = None
        n_lo = 0
        if plan.exact_stem:
            xp = plan.xp_for(parity)
            ops.append(L.op_stem_exact_prep(scale, shift, xp, use_rgb, use_depth, n_dd, use_td))
            scale, shift, n_lo = xp[:32], xp[32:64], (2 if use_td else 0)
        if not fused_stats and not skip_asm:
            ops.append(assemble(scale, shift, n_lo))
        L.run_ops(ops, dev)
        if not prepare_only:
            self._run_backbone(plan, parity)

    def _run_backbone(self, plan, parity=0):
        if getattr(self, "_embed", None):
            acts = self._cur_actions
            if not acts.is_cuda:
                raise L.PnvoError("actions must be a CUDA tensor")
            plan.actions.copy_(acts.reshape(-1))
        ver = self._weights_version()
        if ver != self._packed_version or plan is not getattr(self, "_packed_plan", None):
            plan.pack_prog.run(plan.dev)
            self._packed_version, self._packed_plan = ver, plan
        plan.programs_for(parity)[0].run(plan.dev)

    def _run_forward(self, plan, obs, training):
        if self._is_raw(obs):
            return self._run_forward_raw(plan, obs, training)
        enc = self.visual_encoder
        dev = plan.dev
        srcs, nch, pre, lut_prev, lut_cur = [], [], [], [], []
        for si, (k, n, scale) in enumerate(enc._sources):
            t = obs[k]
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            assert t.shape[-1] == n, f"{k}: expected {n} channels, got {t.shape[-1]}"
            srcs.append(t)
            nch.append(n)
            pre.append(scale)
            lut_prev += [(si, c) for c in range(n // 2)]
            lut_cur += [(si, c) for c in range(n // 2, n)]
        lut = lut_prev + lut_cur  # [prev of every source..., cur of every source...] (vo_cnn.py:169-174)
        C = enc.input_channels
        n_pix = plan.B * plan.H * plan.W
        rmv = enc.running_mean_and_var
        have_rmv = isinstance(rmv, RunningMeanAndVar)
        ops = []
        if have_rmv:
            if training:
                ops.append(L.op_zero(plan.in_stats))
                ops.append(L.op_input_stats(srcs, nch, pre, lut, C, plan.cin_pad, n_pix, plan.in_stats))
                L.run_ops(ops, dev)
                ops = []
                if plan.world_size > 1:
                    _allreduce_stats(plan.in_stats)
            ops.append(L.op_rmv_update(plan.in_stats, rmv._mean, rmv._var, rmv._count, plan.in_scale, plan.in_shift, C,
                                       training, True, plan.B * plan.world_size, plan.H * plan.W))
            scale, shift = plan.in_scale, plan.in_shift
        else:
            scale = shift = None
        ops.append(L.op_assemble(srcs, nch, pre, lut, C, plan.cin_pad, n_pix, scale, shift, plan.x0,
                                 row_w=plan.W if plan.x0_pitch else 0, out_pitch=plan.x0_pitch, out_lo=plan.lo(plan.x0)))
        L.run_ops(ops, dev)
        self._run_backbone(plan)

    def forward(self, observation_pairs):
        params = [p for _, p in self.named_parameters()]
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _VOFunction.apply(self, observation_pairs, self.training, need_grad, *params)


def _variant(name, required=(), forbidden=(), backbone="resnet18", widen=1, no_dd=False):
    """Builds one of the registered subclasses (vo_cnn.py:236-561); they differ only in their asserts
    and in the encoder width."""

    def __init__(self, *, observation_space, observation_size, hidden_size=512, resnet_baseplanes=32,
                 backbone="resnet18", normalize_visual_inputs=False, output_dim=DEFAULT_DELTA_STATE_SIZE,
                 dropout_p=0.2, discretized_depth_channels=0, top_down_view_pair_channel=TOP_DOWN_VIEW_PAIR_CHANNEL):
        assert backbone == self._required_backbone
        if self._no_dd:
            assert discretized_depth_channels == 0
        for k in self._required:
            assert k in observation_space
        for k in self._forbidden:
            assert k not in observation_space
        VisualOdometryCNNBase.__init__(
            self, observation_space=observation_space, observation_size=observation_size, hidden_size=hidden_size,
            resnet_baseplanes=resnet_baseplanes * self._widen, backbone=backbone,
            normalize_visual_inputs=normalize_visual_inputs, output_dim=output_dim, dropout_p=dropout_p,
            discretized_depth_channels=discretized_depth_channels,
            top_down_view_pair_channel=top_down_view_pair_channel)

    return type(name, (VisualOdometryCNNBase,), dict(
        __init__=__init__, _required=tuple(required), _forbidden=tuple(forbidden), _required_backbone=backbone,
        _widen=widen, _no_dd=no_dd, __doc__=f"registered VO variant {name} (vo_cnn.py:236-561)"))


DD, TD = "discretized_depth", "top_down_view"
VisualOdometryCNN = baseline_registry.register_vo_model(name="vo_cnn")(
    _variant("VisualOdometryCNN", forbidden=(DD, TD), no_dd=True))
VisualOdometryCNNRGB = baseline_registry.register_vo_model(name="vo_cnn_rgb")(
    _variant("VisualOdometryCNNRGB", forbidden=("depth", DD, TD), no_dd=True))
VisualOdometryCNNWider = baseline_registry.register_vo_model(name="vo_cnn_wider")(
    _variant("VisualOdometryCNNWider", forbidden=(DD, TD), widen=2, no_dd=True))
VisualOdometryCNNDeeper = baseline_registry.register_vo_model(name="vo_cnn_deeper")(
    _variant("VisualOdometryCNNDeeper", forbidden=(DD, TD), backbone="resnet101", no_dd=True))
VisualOdometryCNNDiscretizedDepth = baseline_registry.register_vo_model(name="vo_cnn_rgb_d_dd")(
    _variant("VisualOdometryCNNDiscretizedDepth", required=(DD,), forbidden=(TD,)))
VisualOdometryCNN_RGB_D_TopDownView = baseline_registry.register_vo_model(name="vo_cnn_rgb_d_top_down")(
    _variant("VisualOdometryCNN_RGB_D_TopDownView", required=("rgb", "depth", TD), forbidden=(DD,)))
VisualOdometryCNN_RGB_DD_TopDownView = baseline_registry.register_vo_model(name="vo_cnn_rgb_dd_top_down")(
    _variant("VisualOdometryCNN_RGB_DD_TopDownView", required=("rgb", DD, TD), forbidden=("depth",)))
VisualOdometryCNN_D_DD_TopDownView = baseline_registry.register_vo_model(name="vo_cnn_d_dd_top_down")(
    _variant("VisualOdometryCNN_D_DD_TopDownView", required=("depth", DD, TD), forbidden=("rgb",)))
VisualOdometryCNNDiscretizedDepthTopDownView = baseline_registry.register_vo_model(name="vo_cnn_rgb_d_dd_top_down")(
    _variant("VisualOdometryCNNDiscretizedDepthTopDownView", required=(DD, TD)))
LegacyVisualOdometryCNNDiscretizedDepthTopDownView = baseline_registry.register_vo_model(
    name="vo_cnn_discretize_depth_top_down")(
    type("LegacyVisualOdometryCNNDiscretizedDepthTopDownView", (VisualOdometryCNNDiscretizedDepthTopDownView,), {}))
