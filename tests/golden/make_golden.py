#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by executing the UNMODIFIED reference
(/root/reference, imported through refshim.py) on seeded synthetic inputs.  Run in the build
container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

While generating, it also cross-checks the oracles (oracle/*.py) against the reference so a
mismatch is caught at fixture time; tests/test_oracle_golden.py repeats the check from the
committed fixtures alone.
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

import refshim  # noqa: E402

refshim.install()
import torch  # noqa: E402

from oracle import preproc_oracle as po  # noqa: E402
from oracle import vo_oracle as vo  # noqa: E402
from pointnav_vo_b200.utils import synth  # noqa: E402

torch.set_num_threads(8)


def ref_discretizer(n_channels=10):
    """Bind the reference's torch _discretize_depth_func onto a dummy self."""
    src = open(os.path.join(refshim.REF_ROOT, "pointnav_vo/rl/common/base_trainer_with_vo.py")).read()
    start = src.index("    def _discretize_depth_func")
    end = src.index("    def _compute_local_delta_states_from_vo")
    ns = {"torch": torch, "np": np}
    import textwrap

    exec(textwrap.dedent(src[start:end]), ns)
    cfg = types.SimpleNamespace(VO=types.SimpleNamespace(REGRESS_MODEL=types.SimpleNamespace(
        discretized_depth_channels=n_channels, discretize_depth="hard")))
    me = types.SimpleNamespace(config=cfg, _discretized_depth_end_vals=po.discretize_end_vals(n_channels))
    return lambda d: ns["_discretize_depth_func"](me, d)


def gen_preproc():
    from pointnav_vo.utils.geometry_utils import NormalizedDepth2TopDownViewHabitatTorch as RefTD
    from pointnav_vo.rl.common.rollout_storage import RolloutStorage

    ref_td = RefTD(min_depth=0.1, max_depth=10.0, vis_size_h=192, vis_size_w=341, hfov_rad=70)
    orc = po.TopDownOracle()
    disc = ref_discretizer()
    D = edge_depth_frames()
    td_sparse_idx, td_sparse_val, td_ptr = [], [], [0]
    dd_idx = np.empty(D.shape, dtype=np.uint8)
    for i in range(D.shape[0]):
        d = torch.from_numpy(D[i])
        ref = ref_td.gen_top_down_view(d[..., None]).numpy()[..., 0]
        mine = orc.gen_top_down_view(D[i])[..., 0]
        assert np.array_equal(ref, mine), f"top-down oracle != reference on frame {i}"
        nz = np.flatnonzero(ref)
        td_sparse_idx.append(nz.astype(np.int32))
        td_sparse_val.append(ref.reshape(-1)[nz])
        td_ptr.append(td_ptr[-1] + nz.size)
        oh = disc(d).numpy()
        assert np.array_equal(oh, po.discretize_depth_onehot(D[i])), f"discretize oracle != reference on frame {i}"
        dd_idx[i] = oh.argmax(-1).astype(np.uint8)
    # discretisation edge set: every fp16 value in [0,1] and each threshold +-1 ulp (fp32)
    e = edge_values()
    oh = disc(torch.from_numpy(e).reshape(1, -1)).numpy()[0]
    assert np.array_equal(oh, po.discretize_depth_onehot(e))
    out = dict(td_idx=np.concatenate(td_sparse_idx), td_val=np.concatenate(td_sparse_val),
               td_ptr=np.array(td_ptr, dtype=np.int64), dd_idx=dd_idx, edge_bins=oh.argmax(-1).astype(np.uint8))
    # GAE
    for name, (T, N, seed) in {"small": (16, 8, 3), "full": (128, 128, 4)}.items():
        for use_gae in (True, False):
            r, v, m, nv = synth.gae_inputs(T, N, seed)
            space = types.SimpleNamespace(spaces={})
            act_space = types.SimpleNamespace(shape=(1,))
            rs = RolloutStorage(T, N, space, act_space, 8)
            rs.rewards.copy_(torch.from_numpy(r))
            rs.value_preds.copy_(torch.from_numpy(v))
            rs.masks.copy_(torch.from_numpy(m))
            rs.step = T
            rs.compute_returns(torch.from_numpy(nv), use_gae, 0.99, 0.95)
            ref = rs.returns.numpy()
            mine, _ = po.gae_returns(r, v, m, nv, use_gae, 0.99, 0.95)
            assert np.array_equal(ref, mine), f"GAE oracle != reference ({name}, gae={use_gae})"
            out[f"gae_{name}_{int(use_gae)}"] = ref
    np.savez_compressed(os.path.join(HERE, "preproc.npz"), **out)
    print("preproc.npz: %d frames, top-down / discretise / GAE oracles == reference" % D.shape[0])


def edge_depth_frames():
    """24 seeded frames + hand-made edge cases (shared with the tests through synth + this recipe)."""
    D = synth.depth_frames(24, seed=1)
    D[0] = 0
    D[1] = 0
    D[1, 100, 200] = 0.5
    D[2] = 1.0
    D[3, :90] = 0
    D[4, :, :170] = 0
    D[5] = np.float32(1e-6)
    D[6, :, 1:] = 0  # single non-zero column
    D[7, 1:, :] = 0  # single non-zero row
    return D


def edge_values():
    h = np.arange(0, 0x3C01, dtype=np.uint16).view(np.float16).astype(np.float32)  # all fp16 in [0, 1]
    th = np.array(po.discretize_end_vals(10), dtype=np.float32)
    lo = np.nextafter(th, np.float32(-1)).astype(np.float32)
    hi = np.nextafter(th, np.float32(2)).astype(np.float32)
    e = np.concatenate([h, th, lo, hi])
    return np.clip(e, 0, 1).astype(np.float32)


def vo_inputs(B, seed, observation_space):
    """Synthetic frame-pair batch as the reference engine hands it to the model (NHWC fp32)."""
    rgb = synth.rgb_frames(2 * B, seed=seed).reshape(B, 2, synth.H, synth.W, 3)
    rgb = np.concatenate([rgb[:, 0], rgb[:, 1]], axis=-1).astype(np.float32)
    dep = synth.depth_frames(2 * B, seed=seed + 100).reshape(B, 2, synth.H, synth.W)
    obs = {"rgb": rgb, "depth": np.stack([dep[:, 0], dep[:, 1]], axis=-1)}
    if "discretized_depth" in observation_space:
        oh = po.discretize_depth_onehot(dep)  # [B,2,H,W,10]
        obs["discretized_depth"] = np.concatenate([oh[:, 0], oh[:, 1]], axis=-1)
    if "top_down_view" in observation_space:
        orc = po.TopDownOracle()
        td = np.stack([np.stack([orc.gen_top_down_view(dep[b, j])[..., 0] for j in range(2)], -1) for b in range(B)])
        obs["top_down_view"] = td
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in obs.items() if k in observation_space}


VO_CASES = {
    # name: (registry/class, observation_space, backbone, kwargs)
    "r18_30ch": ("vo_cnn_rgb_d_dd_top_down", ["rgb", "depth", "discretized_depth", "top_down_view"], "resnet18",
                 dict(discretized_depth_channels=10)),
    "r18_8ch": ("vo_cnn", ["rgb", "depth"], "resnet18", {}),
    "r50_8ch": ("base", ["rgb", "depth"], "resnet50", {}),
    "r18_8ch_act_embed": ("vo_cnn_act_embed", ["rgb", "depth"], "resnet18", {}),
    # the remaining registered configurations (vo_cnn.py:308-375): 64 base planes / ResNet-101
    "r18_wider": ("vo_cnn_wider", ["rgb", "depth"], "resnet18", {}),
    "r101_deeper": ("vo_cnn_deeper", ["rgb", "depth"], "resnet101", {}),
}


def build_ref_vo(case):
    from pointnav_vo.vo.models import vo_cnn, vo_cnn_act_embed  # noqa: F401
    from pointnav_vo.utils.baseline_registry import baseline_registry

    name, space, backbone, kw = VO_CASES[case]
    cls = vo_cnn.VisualOdometryCNNBase if name == "base" else baseline_registry.get_vo_model(name)
    m = cls(observation_space=space, observation_size=(synth.W, synth.H), hidden_size=512, backbone=backbone,
            normalize_visual_inputs=True, output_dim=3, dropout_p=0.0, **kw)
    sd = synth.fill_state_dict(m.state_dict(), seed=7)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return m, space, backbone


def gen_vo(case, B=2):
    m, space, backbone = build_ref_vo(case)
    ng = m.visual_encoder.backbone.conv1[1].num_groups  # resnet_baseplanes // 2: 16, or 32 for vo_cnn_wider
    obs = vo_inputs(B, seed=11, observation_space=space)
    actions = torch.tensor([1, 2, 3, 1][:B]) if "act_embed" in case else None
    args = (obs, actions) if actions is not None else (obs,)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    out = {"keys": np.array(list(sd.keys()))}
    # eval forward
    m.eval()
    with torch.no_grad():
        y = m(*args)
        taps = {}
        y_o, _ = vo.vo_forward(obs, sd, space, backbone, ngroups=ng, training=False, actions=actions, taps=taps)
    err = (y - y_o).abs().max().item()
    assert err <= 2e-6 * max(1.0, y.abs().max().item()), f"{case}: eval oracle != reference ({err})"
    out["eval_out"] = y.numpy()
    out["eval_compression_mean_abs"] = np.float32(taps["compression"].abs().mean().item())
    # training-mode forward + backward (dropout_p=0): outputs, running stats, grads
    m.train()
    target = torch.from_numpy(np.random.default_rng(5).normal(0, 0.1, size=(B, 3)).astype(np.float32))
    y = m(*args)
    loss = sum(vo.vo_losses(y, target))
    loss.backward()
    out["train_out"] = y.detach().numpy()
    out["train_loss"] = np.float32(loss.item())
    new_sd = m.state_dict()
    for k in ("_mean", "_var", "_count"):
        out["train" + k] = new_sd["visual_encoder.running_mean_and_var." + k].numpy()
    sd_o = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
    y_o, st = vo.vo_forward(obs, sd_o, space, backbone, ngroups=ng, training=True, actions=actions)
    loss_o = sum(vo.vo_losses(y_o, target))
    loss_o.backward()
    assert (y.detach() - y_o.detach()).abs().max().item() <= 2e-6
    assert torch.allclose(st[0], new_sd["visual_encoder.running_mean_and_var._mean"], rtol=1e-6, atol=1e-7)
    assert torch.allclose(st[1], new_sd["visual_encoder.running_mean_and_var._var"], rtol=1e-6, atol=1e-7)
    gnorm, gkeys = [], []
    for k, p in m.named_parameters():
        g = p.grad
        g_o = sd_o[k].grad
        rel = (g - g_o).norm().item() / max(g.norm().item(), 1e-12)
        assert rel <= 1e-4, f"{case}: grad oracle != reference for {k} ({rel})"
        gkeys.append(k)
        gnorm.append(g.norm().item())
        if g.numel() <= 4096 or k.endswith("conv1.0.weight"):
            out["grad/" + k] = g.numpy()
    out["grad_keys"] = np.array(gkeys)
    out["grad_norms"] = np.array(gnorm, dtype=np.float32)
    if actions is not None:
        out["actions"] = actions.numpy()
    out["target"] = target.numpy()
    out["ngroups"] = np.int64(ng)
    np.savez_compressed(os.path.join(HERE, f"vo_{case}.npz"), **out)
    print(f"vo_{case}.npz: eval/train outputs, running stats, {len(gkeys)} grads; oracle == reference")


def ref_engine_losses():
    """The reference's two loss functions (vo_cnn_engine.py:135-198, vo_cnn_regression_geo_invariance_engine.py:367-449),
    executed from their own source text on a dummy self: their modules import h5py / habitat through the dataset."""
    import textwrap

    out = {}
    for rel, name, end in (("pointnav_vo/vo/engine/vo_cnn_engine.py", "_compute_loss", "    def _compute_loss_weights"),
                           ("pointnav_vo/vo/engine/vo_cnn_regression_geo_invariance_engine.py",
                            "_compute_geo_invariance_inverse_loss", "    def _process_one_batch")):
        src = open(os.path.join(refshim.REF_ROOT, rel)).read()
        a = src.index("    def " + name)
        b = src.index(end)
        ns = {"torch": torch, "np": np, "EPSILON": 1e-8, "DEFAULT_LOSS_WEIGHTS": {"dx": 1.0, "dz": 1.0, "dyaw": 1.0},
              "CUR_REL_TO_PREV": 0, "PREV_REL_TO_CUR": 1, "MOVE_FORWARD": 1, "TURN_LEFT": 2, "TURN_RIGHT": 3}
        exec(textwrap.dedent(src[a:b]), ns)
        out[name] = ns[name]
    return out


def gen_losses():
    """Total training loss of one VO model as _process_one_batch composes it
    (vo_cnn_regression_geo_invariance_engine.py:676-792) from the reference's own _compute_loss /
    _compute_geo_invariance_inverse_loss, and its gradient w.r.t. the predictions: plain batches, geometric-invariance
    batches (one mean per data type) and inverse_joint_train batches with unpaired MOVE_FORWARD rows in between."""
    fn = ref_engine_losses()
    rng = np.random.default_rng(17)
    out = {}
    cases = {
        # actions, data types (None = no geometric-invariance types), loss_inv_weight
        "plain": ([1, 2, 3, 1, 2, 3, 3, 2, 1, 1], None, 0.0),
        "types_only": ([2, 3, 3, 2, 2, 3, 1, 3, 2], [0, 1, 0, 1, 0, 1, 0, 0, 1], 0.0),
        "joint": ([2, 3, 1, 3, 2, 1, 1, 2, 3, 3, 2], [0, 1, 0, 0, 1, 0, 0, 0, 1, 0, 1], 0.7),
    }
    w = {"dx": 1.0, "dz": 2.0, "dyaw": 0.5}
    for name, (acts, types, inv_w) in cases.items():
        B = len(acts)
        pred = torch.from_numpy(rng.normal(0, 0.2, (B, 3)).astype(np.float32)).requires_grad_(True)
        tgt = torch.from_numpy(rng.normal(0, 0.2, (B, 3)).astype(np.float32))
        actions = torch.tensor(acts).long().unsqueeze(1)
        dzm = torch.from_numpy((rng.random((B, 1)) > 0.3).astype(np.float32))
        groups = [torch.arange(B)] if types is None else [
            torch.nonzero(torch.tensor(types) == t, as_tuple=True)[0] for t in (0, 1)]
        loss = 0.0
        for idx in groups:                                       # :676-750
            for i, d in enumerate(("dx", "dz", "dyaw")):
                tg = [tgt[idx, k:k + 1] for k in range(3)]
                lw = {k: torch.full((idx.numel(), 1), v) for k, v in w.items()}
                loss = loss + fn["_compute_loss"](None, pred[idx, i:i + 1], tg, d_type=d, loss_weights=lw,
                                                  dz_regress_masks=dzm[idx])[0]
        if inv_w > 0:                                            # :781-792
            dt = torch.tensor(types).float().unsqueeze(1)
            valid = torch.nonzero((actions == 2) | (actions == 3), as_tuple=True)[0]
            l_inv, _, _ = fn["_compute_geo_invariance_inverse_loss"](None, pred[valid, :], actions[valid, :], dt[valid, :])
            loss = loss + inv_w * l_inv
        loss.backward()
        mine_pred = pred.detach().clone().requires_grad_(True)
        mine = vo.vo_total_loss(mine_pred, tgt, actions.reshape(-1), None if types is None else torch.tensor(types),
                                (w["dx"], w["dz"], w["dyaw"]), dzm, inv_w)
        mine.backward()
        assert abs(mine.item() - loss.item()) <= 1e-6 * abs(loss.item()), (name, mine.item(), loss.item())
        assert torch.allclose(mine_pred.grad, pred.grad, rtol=1e-5, atol=1e-8), name
        out.update({f"{name}/pred": pred.detach().numpy(), f"{name}/target": tgt.numpy(),
                    f"{name}/actions": np.array(acts, np.int64), f"{name}/dz_mask": dzm.numpy()[:, 0],
                    f"{name}/loss": np.float32(loss.item()), f"{name}/grad": pred.grad.numpy(),
                    f"{name}/inv_w": np.float32(inv_w)})
        if types is not None:
            out[f"{name}/data_types"] = np.array(types, np.int64)
    out["loss_weights"] = np.array([w["dx"], w["dz"], w["dyaw"]], np.float32)
    np.savez_compressed(os.path.join(HERE, "vo_losses.npz"), **out)
    print("vo_losses.npz: %d batches; vo_total_loss oracle == the reference's composition" % len(cases))


def _ref_policy(vis_types, normalize, backbone="resnet18"):
    from pointnav_vo.rl.policies.resnet_policy import PointNavResNetPolicy

    spaces = {"pointgoal_with_gps_compass": refshim.Box(-1e9, 1e9, (2,))}
    if "depth" in vis_types:
        spaces["depth"] = refshim.Box(0.0, 1.0, (synth.H, synth.W, 1))
    if "rgb" in vis_types:
        spaces["rgb"] = refshim.Box(0, 255, (synth.H, synth.W, 3))
    pol = PointNavResNetPolicy(observation_space=refshim.Dict(spaces), action_space=refshim.Discrete(4), hidden_size=512,
                               rnn_type="LSTM", num_recurrent_layers=2, backbone=backbone,
                               normalize_visual_inputs=normalize, obs_transform=None, vis_types=list(vis_types))
    sd = synth.fill_state_dict(pol.state_dict(), seed=9)
    pol.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return pol, sd


def gen_policy_rgbd(N=3):
    """The policy with rgb + depth and normalize_visual_inputs=True (resnet_policy.py:61-174): eval-mode act() and a
    training-mode evaluate_actions (running statistics update)."""
    pol, sd = _ref_policy(("rgb", "depth"), True)
    dep = synth.depth_frames(N, seed=31)[..., None]
    rgb = synth.rgb_frames(N, seed=32).astype(np.float32)
    rng = np.random.default_rng(33)
    goal = rng.uniform(-2, 2, size=(N, 2)).astype(np.float32)
    hid = rng.normal(0, 0.5, size=(4, N, 512)).astype(np.float32)
    prev_a = rng.integers(0, 4, size=(N, 1)).astype(np.int64)
    masks = np.array([[1.0], [0.0], [1.0]][:N], dtype=np.float32)
    obs = {"rgb": torch.from_numpy(rgb), "depth": torch.from_numpy(dep),
           "pointgoal_with_gps_compass": torch.from_numpy(goal)}
    pol.eval()
    with torch.no_grad():
        enc = pol.net.visual_encoder(obs)
        value, action, logp, new_hid = pol.act(obs, torch.from_numpy(hid), torch.from_numpy(prev_a),
                                               torch.from_numpy(masks), deterministic=True)
    pol.train()
    with torch.no_grad():
        enc_tr = pol.net.visual_encoder(obs)
    rm = pol.net.visual_encoder.running_mean_and_var
    np.savez_compressed(os.path.join(HERE, "policy_r18_rgbd_norm.npz"), keys=np.array(list(sd.keys())),
                        encoder_out=enc.numpy(), value=value.numpy(), action=action.numpy(), logp=logp.numpy(),
                        new_hidden=new_hid.numpy(), goal=goal, hidden=hid, prev_actions=prev_a, masks=masks,
                        train_encoder_out=enc_tr.numpy(), train_mean=rm._mean.numpy(), train_var=rm._var.numpy(),
                        train_count=rm._count.numpy())
    print("policy_r18_rgbd_norm.npz: rgb + depth policy with input normalisation")


class ActionSpace:
    """habitat's discrete action space as RolloutStorage recognises it (class name test, rollout_storage.py:42-52)."""

    def __init__(self, n):
        self.n = n


def gen_ppo(T=4, N=4):
    """One PPO.update of the UNMODIFIED reference (rl/ppo/ppo.py:62-146 on rl/common/rollout_storage.py) over a seeded
    rollout of the shipped depth-only policy: losses, gradient norms of the first minibatch and every parameter after
    the update (2 epochs x 2 minibatches; the env permutation of recurrent_generator is recorded)."""
    from pointnav_vo.rl.common.rollout_storage import RolloutStorage
    from pointnav_vo.rl.ppo.ppo import PPO

    pol, sd = _ref_policy(("depth",), False)
    pol.train()
    rng = np.random.default_rng(41)
    obs_space = refshim.Dict({"depth": refshim.Box(0.0, 1.0, (synth.H, synth.W, 1)),
                              "pointgoal_with_gps_compass": refshim.Box(-1e9, 1e9, (2,))})
    rs = RolloutStorage(T, N, obs_space, ActionSpace(4), 512, num_recurrent_layers=4)
    dep = synth.depth_frames((T + 1) * N, seed=42).reshape(T + 1, N, synth.H, synth.W, 1)
    goal = rng.uniform(-2, 2, size=(T + 1, N, 2)).astype(np.float32)
    data = dict(depth=dep, goal=goal,
                hidden0=rng.normal(0, 0.3, size=(4, N, 512)).astype(np.float32),
                actions=rng.integers(0, 4, size=(T, N, 1)).astype(np.int64),
                prev_actions=rng.integers(0, 4, size=(T + 1, N, 1)).astype(np.int64),
                masks=(rng.random((T + 1, N, 1)) > 0.2).astype(np.float32),
                rewards=rng.normal(0, 1, size=(T, N, 1)).astype(np.float32),
                value_preds=rng.normal(0, 1, size=(T + 1, N, 1)).astype(np.float32),
                action_log_probs=(-1.4 + 0.3 * rng.normal(0, 1, size=(T, N, 1))).astype(np.float32),
                next_value=rng.normal(0, 1, size=(N, 1)).astype(np.float32))
    rs.observations["depth"].copy_(torch.from_numpy(dep))
    rs.observations["pointgoal_with_gps_compass"].copy_(torch.from_numpy(goal))
    rs.recurrent_hidden_states[0].copy_(torch.from_numpy(data["hidden0"]))
    for k in ("actions", "prev_actions", "masks", "rewards", "value_preds", "action_log_probs"):
        getattr(rs, k).copy_(torch.from_numpy(data[k]))
    rs.step = T
    rs.compute_returns(torch.from_numpy(data["next_value"]), True, 0.99, 0.95)
    agent = PPO(actor_critic=pol, clip_param=0.2, ppo_epoch=2, num_mini_batch=2, value_loss_coef=0.5, entropy_coef=0.01,
                lr=2.5e-4, eps=1e-5, max_grad_norm=0.2, use_normalized_advantage=False)
    perms, real_randperm = [], torch.randperm

    def recording_randperm(n, *a, **k):
        p = real_randperm(n, *a, **k)
        perms.append(p.numpy().copy())
        return p

    first = {}
    real_before_step = agent.before_step

    def before_step():
        if not first:
            first.update({k: p.grad.norm().item() for k, p in pol.named_parameters() if p.grad is not None})
        real_before_step()

    agent.before_step = before_step
    # first minibatch: the tensors entering the loss (ppo.py:86-126) and the total loss handed to backward (:127-135)
    mb = {}
    real_eval, real_gen = pol.evaluate_actions, rs.recurrent_generator

    def eval_actions(*a, **k):
        r = real_eval(*a, **k)
        if "values" not in mb:
            mb.update(values=r[0].detach().numpy().copy(), log_probs=r[1].detach().numpy().copy(),
                      entropy=np.float32(r[2].item()))
        return r

    def generator(*a, **k):
        for sample in real_gen(*a, **k):
            if "value_preds" not in mb:
                mb.update(value_preds=sample[4].numpy().copy(), returns=sample[5].numpy().copy(),
                          old_log_probs=sample[7].numpy().copy(), adv=sample[8].numpy().copy())
            yield sample

    def before_backward(loss):
        mb.setdefault("total", np.float32(loss.item()))

    pol.evaluate_actions, rs.recurrent_generator, agent.before_backward = eval_actions, generator, before_backward
    torch.manual_seed(5)
    torch.randperm = recording_randperm
    try:
        vl, al, ent = agent.update(rs)
    finally:
        torch.randperm = real_randperm
    out = {("in/" + k): v for k, v in data.items() if k != "depth"}  # depth = synth.depth_frames((T + 1) * N, seed=42)
    t = {k: torch.from_numpy(np.asarray(v)) for k, v in mb.items()}
    vl_o, al_o, ent_o = vo.ppo_losses(t["values"], t["log_probs"], t["entropy"], t["value_preds"], t["returns"],
                                      t["old_log_probs"], t["adv"], 0.2, True)
    tot_o = (vl_o * 0.5 + al_o - ent_o * 0.01).item()
    assert abs(tot_o - float(mb["total"])) <= 1e-6 * abs(float(mb["total"])), (tot_o, mb["total"])
    out.update({("mb0/" + k): v for k, v in mb.items()})
    out["returns"] = rs.returns.numpy()
    out["perms"] = np.stack(perms)
    out["losses"] = np.array([vl, al, ent], np.float64)
    out["grad_keys"] = np.array(list(first.keys()))
    out["grad_norms"] = np.array(list(first.values()), np.float32)
    new_sd = pol.state_dict()
    out["keys"] = np.array(list(sd.keys()))
    dn, small = [], {}
    for k in sd:
        d = new_sd[k].numpy() - sd[k]
        dn.append(np.linalg.norm(d))
        if d.size <= 4096:
            small["delta/" + k] = d
    out["delta_norms"] = np.array(dn, np.float32)
    out.update(small)
    np.savez_compressed(os.path.join(HERE, "ppo_update.npz"), **out)
    print("ppo_update.npz: reference PPO.update losses %s, %d minibatch permutations" % (out["losses"], len(perms)))


def gen_policy(N=3):
    from pointnav_vo.rl.policies.resnet_policy import PointNavResNetPolicy

    obs_space = refshim.Dict({
        "depth": refshim.Box(0.0, 1.0, (synth.H, synth.W, 1)),
        "pointgoal_with_gps_compass": refshim.Box(-1e9, 1e9, (2,)),
    })
    pol = PointNavResNetPolicy(observation_space=obs_space, action_space=refshim.Discrete(4), hidden_size=512,
                               rnn_type="LSTM", num_recurrent_layers=2, backbone="resnet18",
                               normalize_visual_inputs=False, obs_transform=None, vis_types=["depth"])
    sd = synth.fill_state_dict(pol.state_dict(), seed=9)
    pol.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    pol.eval()
    dep = synth.depth_frames(N, seed=21)[..., None]
    rng = np.random.default_rng(22)
    goal = rng.uniform(-2, 2, size=(N, 2)).astype(np.float32)
    hid = rng.normal(0, 0.5, size=(4, N, 512)).astype(np.float32)
    prev_a = rng.integers(0, 4, size=(N, 1)).astype(np.int64)
    masks = np.array([[1.0], [0.0], [1.0]][:N], dtype=np.float32)
    obs = {"depth": torch.from_numpy(dep), "pointgoal_with_gps_compass": torch.from_numpy(goal)}
    with torch.no_grad():
        enc = pol.net.visual_encoder(obs)
        enc_o = vo.rl_encoder_forward(obs, {k: v for k, v in pol.state_dict().items()})
        assert (enc - enc_o).abs().max().item() <= 2e-6 * max(1.0, enc.abs().max().item())
        value, action, logp, new_hid = pol.act(obs, torch.from_numpy(hid), torch.from_numpy(prev_a),
                                               torch.from_numpy(masks), deterministic=True)
        feats, _ = pol.net(obs, torch.from_numpy(hid), torch.from_numpy(prev_a), torch.from_numpy(masks))
        logits = pol.action_distribution(feats).logits
    np.savez_compressed(os.path.join(HERE, "policy_r18_depth.npz"), keys=np.array(list(sd.keys())),
                        encoder_out=enc.numpy(), value=value.numpy(), action=action.numpy(), logp=logp.numpy(),
                        new_hidden=new_hid.numpy(), logits=logits.numpy(), goal=goal, hidden=hid, prev_actions=prev_a,
                        masks=masks, output_shape=np.array(pol.net.visual_encoder.output_shape))
    print("policy_r18_depth.npz: encoder / act() outputs; encoder oracle == reference")


if __name__ == "__main__":
    which = sys.argv[1:] or ["preproc", "vo", "policy", "losses", "policy_rgbd", "ppo"]
    if "preproc" in which:
        gen_preproc()
    if "vo" in which:
        for c in VO_CASES:
            gen_vo(c)
    for c in which:
        if c in VO_CASES:
            gen_vo(c)
    if "policy" in which:
        gen_policy()
    if "losses" in which:
        gen_losses()
    if "policy_rgbd" in which:
        gen_policy_rgbd()
    if "ppo" in which:
        gen_ppo()
