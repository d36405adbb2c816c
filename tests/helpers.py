"""Shared test helpers: the same seeded inputs / weights the golden generator used (numpy only)."""
import os

import numpy as np
import torch

from oracle import preproc_oracle as po
from pointnav_vo_b200.utils import synth

VO_CASES = {
    "r18_30ch": ("vo_cnn_rgb_d_dd_top_down", ["rgb", "depth", "discretized_depth", "top_down_view"], "resnet18",
                 dict(discretized_depth_channels=10)),
    "r18_8ch": ("vo_cnn", ["rgb", "depth"], "resnet18", {}),
    "r50_8ch": ("base", ["rgb", "depth"], "resnet50", {}),
    "r18_8ch_act_embed": ("vo_cnn_act_embed", ["rgb", "depth"], "resnet18", {}),
    "r18_wider": ("vo_cnn_wider", ["rgb", "depth"], "resnet18", {}),        # 64 base planes (vo_cnn.py:308-340)
    "r101_deeper": ("vo_cnn_deeper", ["rgb", "depth"], "resnet101", {}),    # vo_cnn.py:343-375
}
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def edge_depth_frames():
    """Must mirror tests/golden/make_golden.py:edge_depth_frames."""
    D = synth.depth_frames(24, seed=1)
    D[0] = 0
    D[1] = 0
    D[1, 100, 200] = 0.5
    D[2] = 1.0
    D[3, :90] = 0
    D[4, :, :170] = 0
    D[5] = np.float32(1e-6)
    D[6, :, 1:] = 0
    D[7, 1:, :] = 0
    return D


def edge_values():
    h = np.arange(0, 0x3C01, dtype=np.uint16).view(np.float16).astype(np.float32)
    th = np.array(po.discretize_end_vals(10), dtype=np.float32)
    lo = np.nextafter(th, np.float32(-1)).astype(np.float32)
    hi = np.nextafter(th, np.float32(2)).astype(np.float32)
    return np.clip(np.concatenate([h, th, lo, hi]), 0, 1).astype(np.float32)


def golden_topdown(pre, i):
    a, b = pre["td_ptr"][i], pre["td_ptr"][i + 1]
    out = np.zeros(192 * 341, dtype=np.float32)
    out[pre["td_idx"][a:b]] = pre["td_val"][a:b]
    return out.reshape(192, 341)


def vo_inputs_np(B, seed, observation_space):
    """Mirror of make_golden.vo_inputs (dd / top-down channels derived by the pinned oracle)."""
    rgb = synth.rgb_frames(2 * B, seed=seed).reshape(B, 2, synth.H, synth.W, 3)
    rgb = np.concatenate([rgb[:, 0], rgb[:, 1]], axis=-1).astype(np.float32)
    dep = synth.depth_frames(2 * B, seed=seed + 100).reshape(B, 2, synth.H, synth.W)
    obs = {"rgb": rgb, "depth": np.stack([dep[:, 0], dep[:, 1]], axis=-1)}
    if "discretized_depth" in observation_space:
        oh = po.discretize_depth_onehot(dep)
        obs["discretized_depth"] = np.concatenate([oh[:, 0], oh[:, 1]], axis=-1)
    if "top_down_view" in observation_space:
        orc = po.TopDownOracle()
        obs["top_down_view"] = np.stack(
            [np.stack([orc.gen_top_down_view(dep[b, j])[..., 0] for j in range(2)], -1) for b in range(B)])
    return {k: np.ascontiguousarray(v) for k, v in obs.items() if k in observation_space}


def vo_inputs(B, seed, observation_space, device="cpu"):
    return {k: torch.from_numpy(v).to(device) for k, v in vo_inputs_np(B, seed, observation_space).items()}


def vo_state_dict(case, seed=7, device="cpu"):
    """Reference-format state_dict filled by synth.fill_state_dict in the golden key order."""
    g = np.load(os.path.join(GOLDEN, f"vo_{case}.npz"))
    shapes = vo_state_shapes(case)
    keys = [str(k) for k in g["keys"]]
    proto = {k: np.empty(shapes[k], dtype=np.float32) for k in keys}
    sd = synth.fill_state_dict(proto, seed=seed)
    return {k: torch.from_numpy(v).to(device) for k, v in sd.items()}


def vo_state_shapes(case):
    from pointnav_vo_b200.vo.models.shapes import vo_state_dict_shapes
    name, space, backbone, kw = VO_CASES[case]
    if case == "r18_wider":
        kw = dict(kw, resnet_baseplanes=64)
    return vo_state_dict_shapes(space, backbone, act_embed="act_embed" in case, **kw)


class _Box:
    def __init__(self, shape):
        self.shape = tuple(shape)


class _Dict:
    def __init__(self, spaces):
        self.spaces = spaces


class _Discrete:
    def __init__(self, n):
        self.n = n


def policy_spaces(vis_types=("depth",)):
    """gym-like observation / action spaces of the shipped depth-only policy (ddppo_pointnav.yaml)."""
    sp = {"pointgoal_with_gps_compass": _Box((2,))}
    if "depth" in vis_types:
        sp["depth"] = _Box((192, 341, 1))
    if "rgb" in vis_types:
        sp["rgb"] = _Box((192, 341, 3))
    return _Dict(sp), _Discrete(4)


def policy_state_dict(seed=9, device="cpu", vis_types=("depth",), normalize=False):
    from pointnav_vo_b200.rl.policies.resnet_policy import PointNavResNetPolicy

    obs_space, act_space = policy_spaces(vis_types)
    pol = PointNavResNetPolicy(observation_space=obs_space, action_space=act_space, backbone="resnet18",
                               vis_types=list(vis_types), normalize_visual_inputs=normalize)
    proto = {k: np.empty(tuple(v.shape), dtype=np.float32) for k, v in pol.state_dict().items()}
    sd = synth.fill_state_dict(proto, seed=seed)
    pol.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return pol.to(device)
