"""Rollout buffers with the reference's interface (pointnav_vo/rl/common/rollout_storage.py:12-220), kept on the
device: `compute_returns` is ONE libpnvo launch (pnvo_gae_scan) instead of 128 sequential steps of ~6 tiny kernels,
and `recurrent_generator` gathers a minibatch with one index_select per buffer instead of a Python loop over
environments (rollout_storage.py:143-199).  Shapes, time-major flattening and the yielded tuple are the reference's."""
import torch

from .rollout_returns import compute_returns as _compute_returns


class RolloutStorage:
    def __init__(self, num_steps, num_envs, observation_space, action_space, recurrent_hidden_state_size,
                 num_recurrent_layers=1):
        T, N = num_steps, num_envs
        self.observations = {k: torch.zeros(T + 1, N, *sp.shape) for k, sp in observation_space.spaces.items()}
        self.recurrent_hidden_states = torch.zeros(T + 1, num_recurrent_layers, N, recurrent_hidden_state_size)
        self.rewards = torch.zeros(T, N, 1)
        self.value_preds = torch.zeros(T + 1, N, 1)
        self.returns = torch.zeros(T + 1, N, 1)
        self.action_log_probs = torch.zeros(T, N, 1)
        discrete = action_space.__class__.__name__ in ("ActionSpace", "Discrete", "_Discrete")
        action_shape = 1 if discrete else action_space.shape[0]
        dt = torch.long if discrete else torch.float32
        self.actions = torch.zeros(T, N, action_shape, dtype=dt)
        self.prev_actions = torch.zeros(T + 1, N, action_shape, dtype=dt)
        self.masks = torch.zeros(T + 1, N, 1)
        self.num_steps = T
        self.step = 0

    def _tensors(self):
        return ("recurrent_hidden_states", "rewards", "value_preds", "returns", "action_log_probs", "actions",
                "prev_actions", "masks")

    def to(self, device):
        self.observations = {k: v.to(device) for k, v in self.observations.items()}
        for name in self._tensors():
            setattr(self, name, getattr(self, name).to(device))

    def insert(self, observations, recurrent_hidden_states, actions, action_log_probs, value_preds, rewards, masks):
        t = self.step
        for k, v in observations.items():
            self.observations[k][t + 1].copy_(v)
        self.recurrent_hidden_states[t + 1].copy_(recurrent_hidden_states)
        self.actions[t].copy_(actions)
        self.prev_actions[t + 1].copy_(actions)
        self.action_log_probs[t].copy_(action_log_probs)
        self.value_preds[t].copy_(value_preds)
        self.rewards[t].copy_(rewards)
        self.masks[t + 1].copy_(masks)
        self.step = t + 1

    def after_update(self):
        t = self.step
        for v in self.observations.values():
            v[0].copy_(v[t])
        self.recurrent_hidden_states[0].copy_(self.recurrent_hidden_states[t])
        self.masks[0].copy_(self.masks[t])
        self.prev_actions[0].copy_(self.prev_actions[t])
        self.step = 0

    def compute_returns(self, next_value, use_gae, gamma, tau, mode="exact"):
        """rollout_storage.py:102-120 over the first `step` time steps; mode "exact" reproduces the reference's
        rounding order bit for bit, "scan" is the warp-scan variant."""
        t = self.step
        _compute_returns(self.rewards[:t], self.value_preds[:t + 1], self.masks[:t + 1], next_value, self.returns[:t + 1],
                         use_gae, gamma, tau, mode=mode)

    def recurrent_generator(self, advantages, num_mini_batch):
        N = self.rewards.size(1)
        assert N >= num_mini_batch, (f"Trainer requires the number of processes ({N}) to be greater than or equal to "
                                     f"the number of trainer mini batches ({num_mini_batch}).")
        per = N // num_mini_batch
        T = self.step
        fixed = getattr(self, "fixed_perms", None)  # tests: replay a recorded sequence of env permutations
        perm = fixed.pop(0) if fixed else torch.randperm(N, device=self.rewards.device)
        for start in range(0, N, per):
            ind = perm[start:start + per]
            n = ind.numel()

            def take(x):  # [T', N, ...] -> [T * n, ...], time-major like _flatten_helper
                y = x[:T].index_select(1, ind)
                return y.reshape(T * n, *y.shape[2:])

            obs = {k: take(v) for k, v in self.observations.items()}
            yield (obs, self.recurrent_hidden_states[0].index_select(1, ind), take(self.actions), take(self.prev_actions),
                   take(self.value_preds), take(self.returns), take(self.masks), take(self.action_log_probs),
                   take(advantages))

    @staticmethod
    def _flatten_helper(t, n, tensor):
        return tensor.view(t * n, *tensor.size()[2:])
