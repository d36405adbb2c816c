#!/usr/bin/env python
"""One PPO update (K4: depth-only ResNet-18 policy, 128 envs x 128 steps, 2 minibatches of 8192 frames) for profiling:
`ncu --metrics gpu__time_duration.sum ... python tools/k4_update.py` gives the launch list of the update."""
import os
import sys
import time
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
H, W = 192, 341


def main(reps=1):
    from pointnav_vo_b200.rl.common.rollout_storage import RolloutStorage
    from pointnav_vo_b200.rl.policies.resnet_policy import PointNavResNetPolicy
    from pointnav_vo_b200.rl.ppo.ppo import PPO

    dev = torch.device("cuda", 0)
    T = N = 128
    box = lambda *sh: types.SimpleNamespace(shape=tuple(sh))  # noqa: E731
    obs_space = types.SimpleNamespace(spaces={"depth": box(H, W, 1), "pointgoal_with_gps_compass": box(2)})

    class ActionSpace:
        n = 4

    torch.manual_seed(0)
    pol = PointNavResNetPolicy(observation_space=obs_space, action_space=ActionSpace(), backbone="resnet18",
                               vis_types=["depth"]).to(dev)
    rs = RolloutStorage(T, N, obs_space, ActionSpace(), 512, num_recurrent_layers=pol.net.num_recurrent_layers)
    rs.to(dev)
    g = torch.Generator(device=dev).manual_seed(1)
    rs.observations["depth"].copy_(torch.rand(T + 1, N, H, W, 1, device=dev, generator=g))
    rs.observations["pointgoal_with_gps_compass"].copy_(torch.rand(T + 1, N, 2, device=dev, generator=g) * 4 - 2)
    rs.rewards.copy_(torch.randn(T, N, 1, device=dev, generator=g))
    rs.value_preds.copy_(torch.randn(T + 1, N, 1, device=dev, generator=g))
    rs.masks.copy_((torch.rand(T + 1, N, 1, device=dev, generator=g) < 0.98).float())
    rs.actions.copy_(torch.randint(0, 4, (T, N, 1), device=dev, generator=g))
    rs.prev_actions.copy_(torch.randint(0, 4, (T + 1, N, 1), device=dev, generator=g))
    rs.action_log_probs.copy_(-1.4 + 0.1 * torch.randn(T, N, 1, device=dev, generator=g))
    rs.step = T
    rs.compute_returns(torch.randn(N, 1, device=dev, generator=g), True, 0.99, 0.95)
    agent = PPO(pol, clip_param=0.2, ppo_epoch=1, num_mini_batch=2, value_loss_coef=0.5, entropy_coef=0.01, lr=2.5e-4,
                eps=1e-5, max_grad_norm=0.2, use_normalized_advantage=False)
    pol.train()
    agent.update(rs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        agent.update(rs)
    torch.cuda.synchronize()
    print("ppo_update ms", (time.perf_counter() - t0) * 1e3 / max(reps, 1))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
