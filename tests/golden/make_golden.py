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
        y_o, _ = vo.vo_forward(obs, sd, space, backbone, training=False, actions=actions, taps=taps)
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
    y_o, st = vo.vo_forward(obs, sd_o, space, backbone, training=True, actions=actions)
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
    np.savez_compressed(os.path.join(HERE, f"vo_{case}.npz"), **out)
    print(f"vo_{case}.npz: eval/train outputs, running stats, {len(gkeys)} grads; oracle == reference")


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
    which = sys.argv[1:] or ["preproc", "vo", "policy"]
    if "preproc" in which:
        gen_preproc()
    if "vo" in which:
        for c in VO_CASES:
            gen_vo(c)
    if "policy" in which:
        gen_policy()
