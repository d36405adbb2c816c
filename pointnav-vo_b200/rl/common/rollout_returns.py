"""GAE / discounted returns on the device (pointnav_vo/rl/common/rollout_storage.py:102-120)."""
import torch

from ... import lib as _lib


def compute_returns(rewards, value_preds, masks, next_value, returns, use_gae, gamma, tau, mode="exact"):
    """In-place twin of RolloutStorage.compute_returns on its own buffers.
    rewards [T,N,1], value_preds / masks / returns [T+1,N,1], next_value [N,1]; all CUDA fp32 contiguous.
    mode "exact": sequential in t with the reference's rounding order; "scan": warp-scan over t."""
    T, N = rewards.shape[0], rewards.shape[1]
    for t in (rewards, value_preds, masks, returns):
        assert t.is_contiguous() and t.dtype == torch.float32
    nv = next_value.contiguous().float()
    lib = _lib.load()
    # gamma * tau is a python-double product rounded once when torch multiplies it into an fp32 tensor
    _lib.check(lib.pnvo_gae_scan(_lib.ptr(rewards), _lib.ptr(value_preds), _lib.ptr(masks), _lib.ptr(nv),
                                 _lib.ptr(returns), T, N, int(bool(use_gae)), float(gamma), float(gamma * tau),
                                 0 if mode == "exact" else 1, _lib.stream_ptr(rewards.device)))
    return returns
