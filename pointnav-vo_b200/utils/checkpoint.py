"""Checkpoint compatibility with the reference.

The modules of this package keep the reference's state_dict keys and shapes, so its published checkpoints load with
`load_state_dict`; what differs between the files is the envelope:

  VO, one network per file      {"model_state": sd, ...}                       (act_forward.pth)
  VO, several networks per file {"model_states": {act_id: sd}, "optim_states": ..., "epoch": ..., rng states}
                                (act_left_right_inv_joint.pth; written by
                                vo_cnn_regression_geo_invariance_engine.py:1425-1447, read by
                                rl/common/base_trainer_with_vo.py:84-99 with ACT_NAME2IDX)
  RL policy                     {"state_dict": {"actor_critic.<...>": t}, "config": ..., "extra_state": ...}
                                (rl_tune_vo.pth; rl/ppo/ppo_trainer.py:100-118, read by
                                rl/ddppo/algo/ddppo_trainer.py:136-160)

Weights stay in the reference's OIHW fp32 layout inside the modules; the kernel-native copies (fp16 [Cout][R][S][Cin],
the flipped data-gradient layout, the stem's row layout) are re-derived by the plan's pack program the next time the
module runs -- `load_state_dict` bumps the parameters' version counters, which the modules watch.
"""
import torch

from ..vo.common.common_vars import ACT_NAME2IDX


def _read(ckpt, map_location="cpu"):
    if isinstance(ckpt, (str, bytes)) or hasattr(ckpt, "__fspath__"):
        # the reference's files pickle their yacs config next to the tensors, hence weights_only=False
        return torch.load(ckpt, map_location=map_location, weights_only=False)
    return ckpt


def vo_state_dict_for(ckpt, act_name):
    """The state_dict of the VO network of action `act_name` ("forward" / "left" / "right" / "all") inside a VO
    checkpoint of either format (base_trainer_with_vo.py:92-99)."""
    if "model_state" in ckpt:
        return ckpt["model_state"]
    if "model_states" in ckpt:
        states = ckpt["model_states"]
        idx = ACT_NAME2IDX[act_name]
        if idx in states:
            return states[idx]
        if str(idx) in states:  # files that went through a JSON-ish round trip
            return states[str(idx)]
        raise KeyError(f"checkpoint holds VO networks for actions {sorted(states)}, not {idx} ({act_name})")
    raise ValueError("not a VO checkpoint: neither 'model_state' nor 'model_states'")


def load_vo_checkpoint(ckpt, vo_models, map_location="cpu", strict=True):
    """vo_models: {"forward" | "left" | "right" | "all": module} (the reference trainer's `self.vo_model`) and ONE
    checkpoint, or a dict name -> checkpoint/path as VO.REGRESS_MODEL.pretrained_ckpt gives (left and right usually
    name the same joint file).  Returns {name: result of load_state_dict}."""
    out = {}
    cache = {}
    for name, model in vo_models.items():
        src = ckpt[name] if isinstance(ckpt, dict) and name in ckpt and "model_state" not in ckpt else ckpt
        key = src if isinstance(src, (str, bytes)) else id(src)
        if key not in cache:
            cache[key] = _read(src, map_location)
        out[name] = model.load_state_dict(vo_state_dict_for(cache[key], name), strict=strict)
    return out


def save_vo_checkpoint(path, vo_models, optimizers=None, epoch=0, config=None):
    """Writes the joint format of vo_cnn_regression_geo_invariance_engine.py:1425-1447 (readable by the reference),
    including its four RNG states.  `optimizers`: torch optimisers or `FusedVOTrainStep`s -- the latter's state_dict()
    is in torch.optim.Adam's format (moments sliced from the flat buckets), so either side can resume the other's run.
    Files are read back with torch.load(weights_only=False), as the reference does: load trusted checkpoints only."""
    import random

    import numpy as np

    state = {"epoch": epoch, "config": config,
             "model_states": {ACT_NAME2IDX[k]: m.state_dict() for k, m in vo_models.items()},
             "optim_states": {ACT_NAME2IDX[k]: o.state_dict() for k, o in (optimizers or {}).items()},
             "rnd_state": random.getstate(),
             "np_rnd_state": np.random.get_state(),
             "torch_rnd_state": torch.get_rng_state(),
             "torch_cuda_rnd_state": torch.cuda.get_rng_state_all() if torch.cuda.is_available() else []}
    torch.save(state, path)
    return state


def policy_state_dict(ckpt, prefix="actor_critic."):
    """{"actor_critic.net...": t} -> {"net...": t} (ddppo_trainer.py:142-147)."""
    sd = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def load_policy_checkpoint(ckpt, actor_critic, encoder_only=False, map_location="cpu", strict=True):
    """RL.DDPPO.pretrained (whole actor-critic) or RL.DDPPO.pretrained_encoder (visual encoder only), as
    ddppo_trainer.py:136-160 does."""
    ckpt = _read(ckpt, map_location)
    if encoder_only:
        pfx = "actor_critic.net.visual_encoder."
        sd = {k[len(pfx):]: v for k, v in ckpt["state_dict"].items() if k.startswith(pfx)}
        return actor_critic.net.visual_encoder.load_state_dict(sd, strict=strict)
    return actor_critic.load_state_dict(policy_state_dict(ckpt), strict=strict)
