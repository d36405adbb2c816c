"""Action-embedding VO variants (pointnav_vo/vo/models/vo_cnn_act_embed.py:17-113): parameter layout and
registry names.  State-dict compatible with the reference; the fused forward (encoder + [feat | embed]
hidden layer) is not wired to libpnvo yet -- calling it raises (no silent PyTorch fallback).  None of the
reference's shipped configs uses these variants (configs/vo/vo_pointnav.yaml:66 selects
vo_cnn_rgb_d_dd_top_down)."""
import numpy as np
import torch.nn as nn

from ...model_utils import resnet
from ...utils.baseline_registry import baseline_registry
from ..common.common_vars import DEFAULT_DELTA_STATE_SIZE, EMBED_DIM, N_ACTS
from .vo_cnn import Flatten, ResNetEncoder


@baseline_registry.register_vo_model(name="vo_cnn_act_embed")
class VisualOdometryCNNActEmbed(nn.Module):
    def __init__(self, *, observation_space, observation_size, hidden_size=512, resnet_baseplanes=32,
                 backbone="resnet18", normalize_visual_inputs=False, output_dim=DEFAULT_DELTA_STATE_SIZE,
                 dropout_p=0.2, discretized_depth_channels=0, after_compression_flat_size=2048, n_acts=N_ACTS):
        super().__init__()
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

    def forward(self, observation_pairs, actions):
        raise NotImplementedError("vo_cnn_act_embed: the [features | action embedding] hidden layer is not implemented "
                                  "on the B200 path yet (DESIGN.md, 'not built')")


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
