"""CPU ORACLE (test infrastructure, NOT product code): plain PyTorch fp32 restatement of the
reference's visual-odometry CNN and RL visual encoder, as pure functions of a reference-format
state_dict.  This is the "torch fp32 reference" the floating-point CUDA kernels are compared with.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product path never does.

Parity status: PINNED against the unmodified reference modules executed in the build container
(tests/golden/make_golden.py -> tests/golden/vo_*.npz, tests/test_oracle_golden.py): outputs, running
statistics after a training-mode forward, and parameter gradients.

Reference (paths relative to /root/reference/pointnav_vo):
  assemble_input       vo/models/vo_cnn.py:110-174   (NHWC->NCHW, rgb/255, prev|cur interleave)
  running_mean_var     model_utils/running_mean_and_var.py:22-63
  resnet_forward       model_utils/visual_encoders/resnet.py:29-55,95-120,153-223
  vo_forward           vo/models/vo_cnn.py:82-95,176-179,216-233 ; vo_cnn_act_embed.py:65-75
  rl_encoder_forward   rl/policies/resnet_policy.py:146-174
  vo_losses            vo/engine/vo_cnn_engine.py:135-198 ; vo_cnn_regression_geo_invariance_engine.py:367-449
  vo_total_loss        vo/engine/vo_cnn_regression_geo_invariance_engine.py:676-792 (_process_one_batch's loss)
  ppo_losses           rl/ppo/ppo.py:86-126 (clipped surrogate / clipped value / entropy)
"""
import torch
import torch.nn.functional as F

RESNET_LAYERS = {"resnet18": ("basic", [2, 2, 2, 2]), "resnet50": ("bottleneck", [3, 4, 6, 3]),
                 "resnet101": ("bottleneck", [3, 4, 23, 3])}

OBS_ORDER = ("rgb", "depth", "discretized_depth", "top_down_view")  # vo_cnn.py:114-166 append order

# Optional emulation of the CUDA path's storage precision (fp16 activations / weights, fp32 accumulate):
# with QUANT[0] = True every tensor the kernels keep in fp16 is rounded through fp16 here too (straight-
# through in autograd).  This separates "is the kernel logic right" (tight tolerance against the
# quantised oracle) from "what does fp16 storage cost" (looser tolerance against the fp32 oracle).
QUANT = [False]


def _q(x):
    if not QUANT[0]:
        return x
    return x + (x.half().float() - x).detach()


def _conv(x, w, stride=1, pad=0):
    return _q(F.conv2d(x, _q(w), None, stride, pad))


def assemble_input(obs, observation_space):
    """vo_cnn.py:110-174 -> [B, C, H, W] fp32, channels [prev_rgb, prev_d, prev_dd, prev_td, cur_...]."""
    pairs = []
    for k in OBS_ORDER:
        if k not in observation_space:
            continue
        x = obs[k].permute(0, 3, 1, 2)
        if k == "rgb":
            x = x / 255.0
        n = x.shape[1] // 2
        pairs.append((x[:, :n], x[:, n:]))
    prev = [p[0] for p in pairs]
    cur = [p[1] for p in pairs]
    return torch.cat(prev + cur, dim=1)


def running_mean_var(x, mean, var, count, training=False):
    """running_mean_and_var.py:22-63 (single process).  Returns (normalised x, mean, var, count)."""
    if training:
        new_mean = F.adaptive_avg_pool2d(x, 1).sum(0, keepdim=True)
        new_count = torch.full_like(count, x.size(0))
        new_mean = new_mean / new_count
        new_var = F.adaptive_avg_pool2d((x - new_mean).pow(2), 1).sum(0, keepdim=True)
        new_var = new_var / new_count
        m_a = var * count
        m_b = new_var * new_count
        M2 = m_a + m_b + (new_mean - mean).pow(2) * count * new_count / (count + new_count)
        var = M2 / (count + new_count)
        mean = (count * mean + new_count * new_mean) / (count + new_count)
        count = count + new_count
    stdev = torch.sqrt(torch.max(var, torch.full_like(var, 1e-2)))
    return (x - mean) / stdev, mean, var, count


def _gn(x, sd, key, groups):
    return F.group_norm(x, groups, sd[key + ".weight"], sd[key + ".bias"], 1e-5)


def _basic_block(x, sd, p, ngroups, stride, has_down):
    """resnet.py:29-55"""
    out = _conv(x, sd[p + ".convs.0.weight"], stride, 1)
    out = _q(F.relu(_gn(out, sd, p + ".convs.1", ngroups)))
    out = _conv(out, sd[p + ".convs.3.weight"], 1, 1)
    out = _gn(out, sd, p + ".convs.4", ngroups)
    res = x
    if has_down:
        res = _q(_gn(_conv(x, sd[p + ".downsample.0.weight"], stride, 0), sd, p + ".downsample.1", ngroups))
    return _q(F.relu(out + res))


def _bottleneck(x, sd, p, ngroups, stride, has_down):
    """resnet.py:58-120"""
    out = _conv(x, sd[p + ".convs.0.weight"])
    out = _q(F.relu(_gn(out, sd, p + ".convs.1", ngroups)))
    out = _conv(out, sd[p + ".convs.3.weight"], stride, 1)
    out = _q(F.relu(_gn(out, sd, p + ".convs.4", ngroups)))
    out = _conv(out, sd[p + ".convs.6.weight"])
    out = _gn(out, sd, p + ".convs.7", ngroups)
    res = x
    if has_down:
        res = _q(_gn(_conv(x, sd[p + ".downsample.0.weight"], stride, 0), sd, p + ".downsample.1", ngroups))
    return _q(F.relu(out + res))


def resnet_forward(x, sd, prefix, backbone, ngroups, taps=None):
    """resnet.py:214-223.  taps (optional dict) receives named intermediate activations (NCHW)."""
    kind, layers = RESNET_LAYERS[backbone]
    block = _basic_block if kind == "basic" else _bottleneck
    x = _conv(x, sd[prefix + ".conv1.0.weight"], 2, 3)
    if taps is not None:
        taps["conv1_raw"] = x
    x = _q(F.relu(_gn(x, sd, prefix + ".conv1.1", ngroups)))
    x = F.max_pool2d(x, 3, 2, 1)
    if taps is not None:
        taps["pool"] = x
    for li, nblocks in enumerate(layers, start=1):
        for b in range(nblocks):
            p = f"{prefix}.layer{li}.{b}"
            has_down = (p + ".downsample.0.weight") in sd
            stride = 2 if (b == 0 and li > 1) else 1
            x = block(x, sd, p, ngroups, stride, has_down)
        if taps is not None:
            taps[f"layer{li}"] = x
    return x


def encoder_tail(x, sd, prefix):
    """compression: conv3x3 -> GroupNorm(1, C) -> ReLU (vo_cnn.py:85-95)."""
    x = _conv(x, sd[prefix + ".compression.0.weight"], 1, 1)
    x = F.group_norm(x, 1, sd[prefix + ".compression.1.weight"], sd[prefix + ".compression.1.bias"], 1e-5)
    return _q(F.relu(x))


def vo_forward(obs, sd, observation_space, backbone="resnet18", ngroups=16, training=False,
               actions=None, taps=None):
    """Full VO model forward (dropout_p = 0 semantics; eval mode == reference eval mode).

    Returns (out [B, output_dim], new_running_stats or None).  sd: reference-named tensors
    (vo_cnn.py state_dict keys).  actions: LongTensor[B] for the act-embed variant."""
    pfx = "visual_encoder"
    x = assemble_input(obs, observation_space)
    stats = None
    if (pfx + ".running_mean_and_var._mean") in sd:
        x, m, v, c = running_mean_var(x, sd[pfx + ".running_mean_and_var._mean"],
                                      sd[pfx + ".running_mean_and_var._var"],
                                      sd[pfx + ".running_mean_and_var._count"], training)
        stats = (m, v, c)
    x = _q(x)
    if taps is not None:
        taps["input"] = x
    x = resnet_forward(x, sd, pfx + ".backbone", backbone, ngroups, taps)
    x = encoder_tail(x, sd, pfx)
    if taps is not None:
        taps["compression"] = x
    feat = x.contiguous().view(x.size(0), -1)  # Flatten over NCHW (misc_utils.py:45-47)
    if "action_embedding.weight" in sd:
        emb = F.embedding(actions, sd["action_embedding.weight"])
        feat = torch.cat((feat, emb), dim=1)
        h = F.relu(F.linear(feat, sd["hidden_generator.1.weight"], sd["hidden_generator.1.bias"]))
    else:
        h = F.relu(F.linear(feat, _q(sd["visual_fc.2.weight"]), sd["visual_fc.2.bias"]))
    out = F.linear(h, sd["output_head.1.weight"], sd["output_head.1.bias"])
    return out, stats


def rl_encoder_forward(obs, sd, prefix="net.visual_encoder", backbone="resnet18", ngroups=16,
                       use_rgb=False, use_depth=True, taps=None):
    """rl/policies/resnet_policy.py:146-174 (obs_transform=None, normalize_visual_inputs=False)."""
    inp = []
    if use_rgb:
        inp.append(obs["rgb"].permute(0, 3, 1, 2).contiguous() / 255.0)
    if use_depth:
        inp.append(obs["depth"].permute(0, 3, 1, 2).contiguous())
    x = torch.cat(inp, dim=1)
    x = _q(F.avg_pool2d(x, 2))
    if taps is not None:
        taps["input"] = x
    x = resnet_forward(x, sd, prefix + ".backbone", backbone, ngroups, taps)
    return encoder_tail(x, sd, prefix)


def vo_losses(pred, target, loss_weights=(1.0, 1.0, 1.0), dz_regress_masks=None):
    """vo_cnn_engine.py:135-198: per-delta mean((gt - pred)^2 * w); returns (loss_dx, loss_dz, loss_dyaw)."""
    out = []
    for i in range(3):
        diff = (target[:, i:i + 1] - pred[:, i:i + 1]) ** 2
        if i == 1 and dz_regress_masks is not None:
            diff = dz_regress_masks * diff
        out.append(torch.mean(diff * loss_weights[i]))
    return tuple(out)


def geo_invariance_inverse_loss(deltas, actions, move_forward=1):
    """vo_cnn_regression_geo_invariance_engine.py:367-449; deltas interleaved [a0, b0, a1, b1, ...]
    (a = cur_rel_to_prev, b = prev_rel_to_cur); actions has one entry per row of deltas."""
    a = deltas[0::2]
    b = deltas[1::2]
    act = actions[0::2]
    rot = (a[:, 2] + b[:, 2]) ** 2
    loss_rot = torch.mean(rot)
    yaw = b[:, 2]
    R = torch.stack((torch.cos(yaw), torch.sin(yaw), -1 * torch.sin(yaw), torch.cos(yaw)), dim=1).reshape(-1, 2, 2)
    pred = torch.matmul(R, a[:, :2].unsqueeze(-1)).squeeze(-1)
    pos = (b[:, :2] + pred) ** 2
    fwd = torch.nonzero(act == move_forward, as_tuple=True)[0]
    if fwd.numel() != 0:
        mask = torch.ones_like(pos)
        mask[fwd, 1] = 0.0
        pos = mask * pos
    return loss_rot + torch.mean(pos)


def vo_total_loss(pred, target, actions=None, data_types=None, loss_weights=(1.0, 1.0, 1.0), dz_regress_masks=None,
                  loss_inv_weight=0.0, turn_ids=(2, 3), move_forward=1):
    """Training loss of ONE VO model over its rows, as _process_one_batch composes it
    (vo_cnn_regression_geo_invariance_engine.py:676-792):
      * no geometric-invariance types (data_types is None): sum over dx/dz/dyaw of mean((gt - pred)^2 * w) (:676-699);
      * with data types: the same sum evaluated SEPARATELY on the cur-rel-to-prev rows and on the prev-rel-to-cur rows,
        and added (:700-750) -- each mean is over the rows of its own type;
      * inverse_joint_train: + loss_inv_weight * inversion loss over the rows whose action is TURN_LEFT / TURN_RIGHT,
        which must alternate [cur_rel_to_prev, prev_rel_to_cur, ...] (:781-792, :373-374)."""
    if data_types is None:
        loss = sum(vo_losses(pred, target, loss_weights, dz_regress_masks))
    else:
        data_types = data_types.reshape(-1)
        loss = pred.new_zeros(())
        for t in (0, 1):
            idx = torch.nonzero(data_types == t, as_tuple=True)[0]
            if idx.numel() == 0:
                continue
            m = dz_regress_masks[idx] if dz_regress_masks is not None else None
            loss = loss + sum(vo_losses(pred[idx], target[idx], loss_weights, m))
    if loss_inv_weight > 0:
        a = actions.reshape(-1)
        if data_types is None:
            valid = torch.arange(a.numel())
        else:
            valid = torch.nonzero((a == turn_ids[0]) | (a == turn_ids[1]), as_tuple=True)[0]
            vt = data_types[valid]
            assert bool((vt[0::2] == 0).all()) and bool((vt[1::2] == 1).all())
        if valid.numel():
            loss = loss + loss_inv_weight * geo_invariance_inverse_loss(pred[valid], a[valid], move_forward)
    return loss


def ppo_losses(values, action_log_probs, dist_entropy, value_preds, returns, old_action_log_probs, adv_targ,
               clip_param=0.2, use_clipped_value_loss=True):
    """rl/ppo/ppo.py:86-126 -> (value_loss, action_loss, dist_entropy.mean()); total = value_loss * value_loss_coef +
    action_loss - entropy * entropy_coef (:127-133)."""
    ratio = torch.exp(action_log_probs - old_action_log_probs)
    surr1 = ratio * adv_targ
    surr2 = torch.clamp(ratio, 1.0 - clip_param, 1.0 + clip_param) * adv_targ
    action_loss = -torch.min(surr1, surr2).mean()
    if use_clipped_value_loss:
        value_pred_clipped = value_preds + (values - value_preds).clamp(-clip_param, clip_param)
        value_losses = (values - returns).pow(2)
        value_losses_clipped = (value_pred_clipped - returns).pow(2)
        value_loss = 0.5 * torch.max(value_losses, value_losses_clipped).mean()
    else:
        value_loss = 0.5 * (returns - values).pow(2).mean()
    return value_loss, action_loss, dist_entropy.mean()
