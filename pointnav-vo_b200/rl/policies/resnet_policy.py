"""PointGoal actor-critic with a GroupNorm-ResNet visual encoder
(pointnav_vo/rl/policies/resnet_policy.py:25-282): same constructor, state_dict keys and act() /
get_value() / evaluate_actions() signatures.

The visual path -- 2x2 average pool, ResNet, compression, visual_fc + ReLU -- runs as one libpnvo op
program (forward and backward); the goal / previous-action embeddings, the LSTM and the two linear heads
are a few hundred KFLOP per frame and stay in PyTorch (library LSTM).
"""
import numpy as np
import torch
import torch.nn as nn

from ... import lib as L
from ...engine import EncoderPlan
from ...model_utils import resnet
from ...model_utils.rnns.rnn_state_encoder import RNNStateEncoder
from ...model_utils.running_mean_and_var import RunningMeanAndVar
from ...utils.baseline_registry import baseline_registry
from ...vo.models.vo_cnn import Flatten
from .policy import Net, Policy

GOAL_POLAR_DIM = 2


class ResNetEncoder(nn.Module):
    """resnet_policy.py:61-174 (parameter container; executed through the owner's op program)."""

    def __init__(self, observation_space, baseplanes=32, ngroups=32, spatial_size_w=128, spatial_size_h=128,
                 make_backbone=None, normalize_visual_inputs=False, obs_transform=None, vis_types=("rgb", "depth")):
        super().__init__()
        if obs_transform is not None:
            raise NotImplementedError("obs_transform is not supported on the B200 path (configs use OBS_TRANSFORM none)")
        self.obs_transform = None
        spaces = observation_space.spaces
        self._sources = []
        self._n_input_rgb = self._n_input_depth = 0
        if "rgb" in spaces and "rgb" in vis_types:
            self._n_input_rgb = spaces["rgb"].shape[2]
            spatial_size_w, spatial_size_h = spaces["rgb"].shape[0] // 2, spaces["rgb"].shape[1] // 2
            self._sources.append(("rgb", self._n_input_rgb, 1.0 / 255.0))
        if "depth" in spaces and "depth" in vis_types:
            self._n_input_depth = spaces["depth"].shape[2]
            spatial_size_w, spatial_size_h = spaces["depth"].shape[0] // 2, spaces["depth"].shape[1] // 2
            self._sources.append(("depth", self._n_input_depth, 1.0))
        self.baseplanes, self.ngroups = baseplanes, ngroups
        if normalize_visual_inputs:  # applied to the 2x2-pooled input (resnet_policy.py:168-170)
            self.running_mean_and_var = RunningMeanAndVar(self._n_input_depth + self._n_input_rgb)
        else:
            self.running_mean_and_var = nn.Sequential()
        if not self.is_blind:
            input_channels = self._n_input_depth + self._n_input_rgb
            self.input_channels = input_channels
            self.backbone = make_backbone(input_channels, baseplanes, ngroups)
            # note: the reference reads shape[0] (H) into "w" and shape[1] (W) into "h" (:82-92); only the
            # product and the reported output_shape depend on it, and both are reproduced here
            final_w = int(np.ceil(spatial_size_w * self.backbone.final_spatial_compress))
            final_h = int(np.ceil(spatial_size_h * self.backbone.final_spatial_compress))
            num_compression_channels = int(round(2048 / (final_w * final_h)))
            self.compression = nn.Sequential(
                nn.Conv2d(self.backbone.final_channels, num_compression_channels, kernel_size=3, padding=1, bias=False),
                nn.GroupNorm(1, num_compression_channels), nn.ReLU(True))
            self.output_shape = (num_compression_channels, final_h, final_w)

    @property
    def is_blind(self):
        return self._n_input_rgb + self._n_input_depth == 0

    def layer_init(self):
        for layer in self.modules():
            if isinstance(layer, (nn.Conv2d, nn.Linear)):
                nn.init.kaiming_normal_(layer.weight, nn.init.calculate_gain("relu"))
                if layer.bias is not None:
                    nn.init.constant_(layer.bias, val=0)

    def forward(self, observations):
        raise RuntimeError("ResNetEncoder is executed as part of PointNavResNetNet's op program (libpnvo)")


class _VisualFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, obs, need_grad, *params):
        plan = net._plan_for(obs, need_grad)
        net._run_visual(plan, obs)
        plan.generation = getattr(plan, "generation", 0) + 1  # see _VOFunction: one set of activation buffers per shape
        ctx.net, ctx.plan, ctx.generation = net, plan, plan.generation
        return plan.h32.clone()

    @staticmethod
    def backward(ctx, grad_out):
        net, plan = ctx.net, ctx.plan
        if plan.generation != ctx.generation:
            raise RuntimeError("backward of a policy forward whose activations were overwritten by a later forward of the "
                               "same shape: call backward before the next forward of that shape")
        plan.dout.copy_(grad_out)
        plan.bwd_prog.run(plan.dev)
        flat = plan.grad_flat.clone()
        by_name, off = {}, 0
        for k in plan.param_names():
            m = plan.P[k].numel()
            by_name[k] = flat[off:off + m].view(plan.P[k].shape)
            off += m
        return (None, None, None, *[by_name.get(k) for k in net._visual_param_order])


class PointNavResNetNet(Net):
    """resnet_policy.py:177-282."""

    def __init__(self, observation_space, action_space, goal_sensor_uuid, hidden_size, num_recurrent_layers, rnn_type,
                 backbone, resnet_baseplanes, normalize_visual_inputs, obs_transform=None, vis_types=("rgb", "depth")):
        super().__init__()
        self.prev_action_embedding = nn.Embedding(action_space.n + 1, 32)
        self._n_prev_action = 32
        rnn_input_size = self._n_prev_action
        self.tgt_embeding = nn.Linear(GOAL_POLAR_DIM + 1, 32)
        rnn_input_size += 32
        self._hidden_size = hidden_size
        self._backbone_name = backbone
        self.visual_encoder = ResNetEncoder(observation_space, baseplanes=resnet_baseplanes,
                                            ngroups=resnet_baseplanes // 2, make_backbone=resnet.make_backbone(backbone),
                                            normalize_visual_inputs=normalize_visual_inputs, obs_transform=obs_transform,
                                            vis_types=vis_types)
        if not self.visual_encoder.is_blind:
            self.visual_fc = nn.Sequential(Flatten(),
                                           nn.Linear(int(np.prod(self.visual_encoder.output_shape)), hidden_size),
                                           nn.ReLU(True))
        self.state_encoder = RNNStateEncoder((0 if self.is_blind else self._hidden_size) + rnn_input_size,
                                             self._hidden_size, rnn_type=rnn_type, num_layers=num_recurrent_layers)
        self._plans, self._ptr_sig, self._packed_version, self._packed_plan = {}, None, None, None
        self.precision = "split"  # set_precision
        self._visual_param_order = [k for k, _ in self.named_parameters()
                                    if k.startswith("visual_encoder.") or k.startswith("visual_fc.")]
        self.train()

    @property
    def output_size(self):
        return self._hidden_size

    @property
    def is_blind(self):
        return self.visual_encoder.is_blind

    @property
    def num_recurrent_layers(self):
        return self.state_encoder.num_recurrent_layers

    # ------------------------------------------------------------------ libpnvo runtime
    def _visual_params(self):
        P = dict(self.named_parameters())
        return [P[k] for k in self._visual_param_order]

    def set_precision(self, mode):
        """'split' (default): value + residual fp16 planes and three tensor-core products per forward convolution
        (engine.EncoderPlan(split=True); within 1e-3 of the fp32 reference); 'fp16': single-pass operands (faster,
        ~5e-3)."""
        if mode not in ("fp16", "split"):
            raise ValueError(mode)
        self.precision = mode
        return self

    def _plan_for(self, obs, need_grad):
        enc = self.visual_encoder
        first = obs[enc._sources[0][0]]
        if not first.is_cuda:
            raise L.PnvoError("policy observations must be CUDA tensors: the B200 path has no CPU fallback")
        L.load()
        sig = tuple(p.data_ptr() for p in self._visual_params())
        if sig != self._ptr_sig:
            self._plans.clear()
            self._ptr_sig, self._packed_version = sig, None
        B, H, W = first.shape[0], first.shape[1], first.shape[2]
        split = self.precision == "split"
        key = (B, H, W, bool(need_grad), str(first.device), split)
        plan = self._plans.get(key)
        if plan is None:
            P = {k: p.data for k, p in self.named_parameters()}
            head = dict(fc_w="visual_fc.1.weight", fc_b="visual_fc.1.bias", hidden=self._hidden_size, out_dim=None)
            rmv = enc.running_mean_and_var
            world = 1
            if isinstance(rmv, RunningMeanAndVar) and rmv._distributed:
                world = torch.distributed.get_world_size()
            plan = EncoderPlan(params=P, buffers={}, B=B, H=H, W=W, in_channels=enc.input_channels, world_size=world,
                               sources=enc._sources, backbone=self._backbone_name, baseplanes=enc.baseplanes,
                               ngroups=enc.ngroups, compression_channels=enc.output_shape[0], prefix="visual_encoder",
                               head=head, training=bool(need_grad), avgpool_input=True, device=first.device,
                               split=split, compact_input=not isinstance(rmv, RunningMeanAndVar))
            self._plans[key] = plan
        return plan

    def _run_visual(self, plan, obs):
        rmv = self.visual_encoder.running_mean_and_var
        if isinstance(rmv, RunningMeanAndVar):
            self._run_normalised_input(plan, obs, rmv)
        else:
            L.run_ops(plan.input_ops(obs), plan.dev)
        ver = sum(p._version for p in self._visual_params())
        if ver != self._packed_version or plan is not self._packed_plan:
            plan.pack_prog.run(plan.dev)
            self._packed_version, self._packed_plan = ver, plan
        plan.fwd_prog.run(plan.dev)

    def _run_normalised_input(self, plan, obs, rmv):
        """avg_pool2d -> RunningMeanAndVar -> x0 (resnet_policy.py:168-170): pooled values in fp32, batch statistics and
        the Chan merge in training mode (one packed all-reduce in a process group), then normalise into the fp16 planes."""
        dev = plan.dev
        C, B, h, w = plan.in_channels, plan.B, plan.inH, plan.inW
        if getattr(plan, "pooled32", None) is None:
            plan.pooled32 = torch.empty(B, h, w, C, dtype=torch.float32, device=dev)
        ops = plan.input_ops(obs, out32=plan.pooled32)
        lut = [(0, c) for c in range(C)]
        n_pix = B * h * w
        if self.training:
            ops.append(L.op_zero(plan.in_stats))
            ops.append(L.op_input_stats([plan.pooled32], [C], [1.0], lut, C, plan.cin_pad, n_pix, plan.in_stats))
            if plan.world_size > 1:
                L.run_ops(ops, dev)
                ops = []
                torch.distributed.all_reduce(plan.in_stats)
        ops.append(L.op_rmv_update(plan.in_stats, rmv._mean, rmv._var, rmv._count, plan.in_scale, plan.in_shift, C,
                                   self.training, True, B * plan.world_size, h * w))
        ops.append(L.op_assemble([plan.pooled32], [C], [1.0], lut, C, plan.cin_pad, n_pix, plan.in_scale, plan.in_shift,
                                 plan.x0, out_lo=plan.lo(plan.x0)))
        L.run_ops(ops, dev)

    def visual_features(self, observations):
        """[N, hidden] = ReLU(visual_fc(Flatten(visual_encoder(obs)))) (resnet_policy.py:246-256)."""
        params = self._visual_params()
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _VisualFunction.apply(self, observations, need_grad, *params)

    def forward(self, observations, rnn_hidden_states, prev_actions, masks):
        x = []
        if not self.is_blind:
            x.append(self.visual_features(observations))
        if "pointgoal_with_gps_compass" in observations:
            g = observations["pointgoal_with_gps_compass"]
            g = torch.stack([g[:, 0], torch.cos(-g[:, 1]), torch.sin(-g[:, 1])], -1)
            x.append(self.tgt_embeding(g))
        prev_actions = self.prev_action_embedding(((prev_actions.float() + 1) * masks).long().squeeze(dim=-1))
        x.append(prev_actions)
        x = torch.cat(x, dim=1)
        return self.state_encoder(x, rnn_hidden_states, masks)


@baseline_registry.register_policy(name="resnet_rnn_policy")
class PointNavResNetPolicy(Policy):
    def __init__(self, *, observation_space, action_space, goal_sensor_uuid="pointgoal_with_gps_compass",
                 hidden_size=512, num_recurrent_layers=2, rnn_type="LSTM", resnet_baseplanes=32, backbone="resnet50",
                 normalize_visual_inputs=False, obs_transform=None, vis_types=("rgb", "depth"), **kwargs):
        super().__init__(
            PointNavResNetNet(observation_space=observation_space, action_space=action_space,
                              goal_sensor_uuid=goal_sensor_uuid, hidden_size=hidden_size,
                              num_recurrent_layers=num_recurrent_layers, rnn_type=rnn_type, backbone=backbone,
                              resnet_baseplanes=resnet_baseplanes, normalize_visual_inputs=normalize_visual_inputs,
                              obs_transform=obs_transform, vis_types=vis_types),
            action_space.n)
