"""PPO / DD-PPO update around the B200 actor-critic (pointnav_vo/rl/ppo/ppo.py:14-160,
pointnav_vo/rl/ddppo/algo/ddppo.py:18-96).  Same constructor keywords, `update(rollouts)` contract and return
values.  The clipped-surrogate / clipped-value / entropy terms act on [T*N, 1] tensors (negligible); the policy's
visual encoder forward/backward -- where the time goes (8192 frames per minibatch at the shipped config) -- runs as a
libpnvo op program behind `evaluate_actions`.  Data parallelism: instead of borrowing DistributedDataParallel's reducer
(ddppo.py:68-90) the gradients are flattened into one bucket and summed with ONE all-reduce."""
import torch
import torch.nn as nn
import torch.optim as optim

EPS_PPO = 1e-5


def distributed_mean_and_var(values):
    """ddppo.py:18-42: mean, then mean squared deviation about the GLOBAL mean, over every rank."""
    world = torch.distributed.get_world_size()
    mean = values.mean()
    torch.distributed.all_reduce(mean)
    mean = mean / world
    sq = (values - mean).pow(2).mean()
    torch.distributed.all_reduce(sq)
    return mean, sq / world


class PPO(nn.Module):
    def __init__(self, actor_critic, clip_param, ppo_epoch, num_mini_batch, value_loss_coef, entropy_coef, lr=None,
                 eps=None, max_grad_norm=None, use_clipped_value_loss=True, use_normalized_advantage=True):
        super().__init__()
        self.actor_critic = actor_critic
        self.clip_param, self.ppo_epoch, self.num_mini_batch = clip_param, ppo_epoch, num_mini_batch
        self.value_loss_coef, self.entropy_coef = value_loss_coef, entropy_coef
        self.max_grad_norm = max_grad_norm
        self.use_clipped_value_loss = use_clipped_value_loss
        self.use_normalized_advantage = use_normalized_advantage
        self.optimizer = optim.Adam([p for p in actor_critic.parameters() if p.requires_grad], lr=lr, eps=eps)
        self.device = next(actor_critic.parameters()).device

    def forward(self, *x):
        raise NotImplementedError

    def get_advantages(self, rollouts):
        adv = rollouts.returns[:-1] - rollouts.value_preds[:-1]
        if not self.use_normalized_advantage:
            return adv
        return (adv - adv.mean()) / (adv.std() + EPS_PPO)

    def _losses(self, sample):
        (obs, hidden, actions, prev_actions, value_preds, returns, masks, old_log_probs, adv) = sample
        values, log_probs, entropy, _ = self.actor_critic.evaluate_actions(obs, hidden, prev_actions, masks, actions)
        ratio = torch.exp(log_probs - old_log_probs)
        clipped = torch.clamp(ratio, 1.0 - self.clip_param, 1.0 + self.clip_param)
        action_loss = -torch.min(ratio * adv, clipped * adv).mean()
        if self.use_clipped_value_loss:
            v_clip = value_preds + (values - value_preds).clamp(-self.clip_param, self.clip_param)
            value_loss = 0.5 * torch.max((values - returns).pow(2), (v_clip - returns).pow(2)).mean()
        else:
            value_loss = 0.5 * (returns - values).pow(2).mean()
        return value_loss, action_loss, entropy

    def update(self, rollouts):
        advantages = self.get_advantages(rollouts)
        sums = torch.zeros(3, device=advantages.device)
        for _ in range(self.ppo_epoch):
            for sample in rollouts.recurrent_generator(advantages, self.num_mini_batch):
                value_loss, action_loss, entropy = self._losses(sample)
                self.optimizer.zero_grad()
                total = value_loss * self.value_loss_coef + action_loss - entropy * self.entropy_coef
                self.before_backward(total)
                total.backward()
                self.after_backward(total)
                self.before_step()
                self.optimizer.step()
                self.after_step()
                sums += torch.stack((value_loss.detach(), action_loss.detach(), entropy.detach()))
        n = self.ppo_epoch * self.num_mini_batch
        v, a, e = (sums / n).tolist()  # one device->host read per update instead of three per minibatch
        return v, a, e

    def before_backward(self, loss):
        pass

    def after_backward(self, loss):
        pass

    def before_step(self):
        nn.utils.clip_grad_norm_(self.actor_critic.parameters(), self.max_grad_norm)

    def after_step(self):
        pass


class DecentralizedDistributedMixin:
    """ddppo.py:45-96.  `init_distributed` broadcasts rank 0's weights; gradients are averaged after backward with a
    single all-reduce over one flat bucket (NCCL over NVLink on a B200 box)."""

    def init_distributed(self, find_unused_params=True):
        self._world = torch.distributed.get_world_size()
        for p in self.actor_critic.parameters():
            torch.distributed.broadcast(p.data, 0)
        for b in self.actor_critic.buffers():
            torch.distributed.broadcast(b.data, 0)
        self._find_unused = find_unused_params

    def get_advantages(self, rollouts):
        adv = rollouts.returns[:-1] - rollouts.value_preds[:-1]
        if not self.use_normalized_advantage:
            return adv
        mean, var = distributed_mean_and_var(adv)
        return (adv - mean) / (var.sqrt() + EPS_PPO)

    def after_backward(self, loss):
        super().after_backward(loss)
        params = [p for p in self.actor_critic.parameters() if p.requires_grad]
        for p in params:  # parameters the loss did not reach take part with zeros (DDP's find_unused_parameters)
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        flat = torch.cat([p.grad.reshape(-1) for p in params])
        torch.distributed.all_reduce(flat)
        flat /= self._world
        off = 0
        for p in params:
            n = p.numel()
            p.grad.copy_(flat[off:off + n].view_as(p))
            off += n


class DDPPO(DecentralizedDistributedMixin, PPO):
    pass
