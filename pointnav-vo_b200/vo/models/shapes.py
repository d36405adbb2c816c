"""state_dict key -> shape tables of the VO models (the weight-compatibility contract)."""
import torch


def vo_state_dict_shapes(observation_space, backbone="resnet18", act_embed=False, **kw):
    from . import vo_cnn

    if act_embed:
        from . import vo_cnn_act_embed
        cls = vo_cnn_act_embed.VisualOdometryCNNActEmbed
    else:
        cls = vo_cnn.VisualOdometryCNNBase
    with torch.device("meta"):
        m = cls(observation_space=observation_space, observation_size=(341, 192), hidden_size=512, backbone=backbone,
                normalize_visual_inputs=True, output_dim=3, dropout_p=0.0, **kw)
    return {k: tuple(v.shape) for k, v in m.state_dict().items()}
