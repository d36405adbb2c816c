"""PPO / DD-PPO update around the B200 actor-critic (pointnav_vo/rl/ppo/ppo.py:14-160,
pointnav_vo/rl/ddppo/algo/ddppo.py:18-96).  Same constructor keywords, `update(rollouts)` contract and return
values.  The clipped-surrogate / clipped-value terms and their gradient are ONE libpnvo launch (pnvo_ppo_loss,
SURVEY 8f-3); the policy's visual encoder forward/backward -- where the time goes (8192 frames per minibatch at the shipped config) -- runs as a
libpnvo op program behind `evaluate_actions`.  Data parallelism: instead of borrowing DistributedDataParallel's reducer
(ddppo.py:68-90) the gradients are flattened into one bucket and summed with ONE all-reduce."""
import torch
import torch.nn as nn
import torch.optim as optim

from ... import lib as L

EPS_PPO = 1e-5


class _FusedPPOLoss(torch.autograd.Function):
    """value_loss * value_loss_coef + action_loss of ppo.py:101-126 with its gradient, one libpnvo launch
    (pnvo_ppo_loss) instead of ~25 elementwise kernels forward and as many backward.  Returns (partial total,
    value_loss, action_loss); only the first output is differentiable."""

    @staticmethod
    def forward(ctx, values, log_probs, value_preds, returns, old_log_probs, adv, clip, use_clipped, value_coef):
        f = [t.detach().reshape(-1).contiguous().float() for t in (values, log_probs, value_preds, returns, old_log_probs, adv)]
        n = f[0].numel()
        for t in f:
            assert t.numel() == n and t.is_cuda, "PPO loss terms must be CUDA tensors of one length (no CPU fallback)"
        losses = torch.empty(2, dtype=torch.float32, device=f[0].device)
        d_v, d_lp = torch.empty_like(f[0]), torch.empty_like(f[1])
        L.check(L.load().pnvo_ppo_loss(L.ptr(f[0]), L.ptr(f[1]), L.ptr(f[2]), L.ptr(f[3]), L.ptr(f[4]), L.ptr(f[5]), n,
                                       float(clip), int(bool(use_clipped)), float(value_coef), L.ptr(losses), L.ptr(d_v),
                                       L.ptr(d_lp), L.stream_ptr(f[0].device)))
        ctx.save_for_backward(d_v, d_lp)
        ctx.shapes = (values.shape, log_probs.shape)
        value_loss, action_loss = losses[0], losses[1]
        total = value_loss * value_coef + action_loss
        ctx.mark_non_differentiable(value_loss, action_loss)
        return total, value_loss, action_loss

    @staticmethod
    def backward(ctx, g_total, _gv, _ga):
        d_v, d_lp = ctx.saved_tensors
        return (g_total * d_v.view(ctx.shapes[0]), g_total * d_lp.view(ctx.shapes[1]), None, None, None, None, None, None,
                None)


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
        """(total loss, value_loss, action_loss, entropy) of one minibatch (ppo.py:86-133)."""
        (obs, hidden, actions, prev_actions, value_preds, returns, masks, old_log_probs, adv) = sample
        values, log_probs, entropy, _ = self.actor_critic.evaluate_actions(obs, hidden, prev_actions, masks, actions)
        partial, value_loss, action_loss = _FusedPPOLoss.apply(values, log_probs, value_preds, returns, old_log_probs, adv,
                                                               self.clip_param, self.use_clipped_value_loss,
                                                               self.value_loss_coef)
        return partial - entropy * self.entropy_coef, value_loss, action_loss, entropy

    def update(self, rollouts):
        advantages = self.get_advantages(rollouts)
        sums = torch.zeros(3, device=advantages.device)
        for _ in range(self.ppo_epoch):
            for sample in rollouts.recurrent_generator(advantages, self.num_mini_batch):
                total, value_loss, action_loss, entropy = self._losses(sample)
                self.optimizer.zero_grad()
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
