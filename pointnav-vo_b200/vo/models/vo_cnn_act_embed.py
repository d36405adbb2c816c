"""Action-embedding VO variants (pointnav_vo/vo/models/vo_cnn_act_embed.py:17-113): same constructor keywords,
registry names, state_dict keys and `forward(observation_pairs, actions)` signature as the reference.  The encoder
and the visual columns of the hidden Linear run as the usual libpnvo op program; the 32 embedding columns are a
rank-32 fp32 update (csrc/act_embed.cu) sharing the hidden layer's Dropout semantics."""
import numpy as np
import torch
import torch.nn as nn

from ...model_utils import resnet
from ...utils.baseline_registry import baseline_registry
from ..common.common_vars import DEFAULT_DELTA_STATE_SIZE, EMBED_DIM, N_ACTS
from .vo_cnn import Flatten, ResNetEncoder, VisualOdometryCNNBase, _VOFunction


@baseline_registry.register_vo_model(name="vo_cnn_act_embed")
class VisualOdometryCNNActEmbed(VisualOdometryCNNBase):
    """Shares VisualOdometryCNNBase's runtime (plan cache, autograd boundary, raw-input path); only the module
    tree -- and therefore the state_dict keys -- differs (vo_cnn_act_embed.py:17-75)."""

    def __init__(self, *, observation_space, observation_size, hidden_size=512, resnet_baseplanes=32,
                 backbone="resnet18", normalize_visual_inputs=False, output_dim=DEFAULT_DELTA_STATE_SIZE,
                 dropout_p=0.2, discretized_depth_channels=0, after_compression_flat_size=2048, n_acts=N_ACTS):
        nn.Module.__init__(self)
        self.action_embedding = nn.Embedding(n_acts + 1, EMBED_DIM)
        self.visual_encoder = ResNetEncoder(
            observation_space=observation_space, observation_size=observation_size, baseplanes=resnet_baseplanes,
            ngroups=resnet_baseplanes // 2, make_backbone=resnet.make_backbone(backbone),
            normalize_visual_inputs=normalize_visual_inputs, discretized_depth_channels=discretized_depth_channels,
            after_compression_flat_size=after_compression_flat_size)
        self.flatten = Flatten()
        self.hidden_generator = nn.Sequential(
            nn.Dropout(dropout_p),
            nn.Linear(int(np.prod(self.visual_encoder.output_shape)) + EMBED_DIM, hidden_size), nn.ReLU(True))
        self.output_head = nn.Sequential(nn.Dropout(dropout_p), nn.Linear(hidden_size, output_dim))
        nn.init.orthogonal_(self.output_head[1].weight)
        nn.init.constant_(self.output_head[1].bias, 0)
        self._backbone_name = backbone
        self._hidden_size, self._output_dim, self._dropout_p = hidden_size, output_dim, dropout_p
        self._plans = {}
        self._packed_version = None
        self._ptr_sig = None
        self._param_order = [k for k, _ in self.named_parameters()]
        self._fc_keys = dict(fc_w="hidden_generator.1.weight", fc_b="hidden_generator.1.bias",
                             out_w="output_head.1.weight", out_b="output_head.1.bias")
        self._embed = dict(table="action_embedding.weight", dim=EMBED_DIM)
        self._cur_actions = None
        self.raw_fp32 = False

    def forward(self, observation_pairs, actions):
        """actions: LongTensor [B] (or [B, 1]) of action ids, as the reference's engine passes them."""
        self._cur_actions = actions
        params = [p for _, p in self.named_parameters()]
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _VOFunction.apply(self, observation_pairs, self.training, need_grad, *params)


@baseline_registry.register_vo_model(name="vo_cnn_wider_act_embed")
class VisualOdometryCNNWiderActEmbed(VisualOdometryCNNActEmbed):
    def __init__(self, *, observation_space, observation_size, hidden_size=512, resnet_baseplanes=32,
                 backbone="resnet18", normalize_visual_inputs=False, output_dim=DEFAULT_DELTA_STATE_SIZE,
                 dropout_p=0.2, n_acts=N_ACTS, discretized_depth_channels=0):
        assert backbone == "resnet18"
        assert discretized_depth_channels == 0
        assert "discretized_depth" not in observation_space
        assert "top_down_view" not in observation_space
        super().__init__(observation_space=observation_space, observation_size=observation_size,
                         hidden_size=hidden_size, resnet_baseplanes=2 * resnet_baseplanes, backbone=backbone,
                         normalize_visual_inputs=normalize_visual_inputs, output_dim=output_dim, dropout_p=dropout_p,
                         n_acts=n_acts)
