"""CPU: the oracles (oracle/*.py) reproduce the committed golden fixtures, which were produced by the
unmodified reference (tests/golden/make_golden.py).  Bit-exact for index / integer work."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

from oracle import preproc_oracle as po
from oracle import vo_oracle as vo
from pointnav_vo_b200.utils import synth
from tests import helpers


@pytest.fixture(scope="module")
def pre(golden_dir):
    return np.load(os.path.join(golden_dir, "preproc.npz"))


def test_topdown_oracle_matches_reference_bit_exact(pre):
    D = helpers.edge_depth_frames()
    orc = po.TopDownOracle()
    for i in range(D.shape[0]):
        ref = helpers.golden_topdown(pre, i)
        assert np.array_equal(orc.gen_top_down_view(D[i])[..., 0], ref), i


def test_topdown_invariants(pre):
    # geometry_utils.py:640-641,651: coords in range; all-zero depth -> all-zero map (:519-525)
    orc = po.TopDownOracle()
    assert orc.gen_top_down_view(np.zeros((192, 341), np.float32)).sum() == 0
    D = helpers.edge_depth_frames()
    v = orc.gen_top_down_view(D[10])
    assert v.shape == (192, 341, 1) and v.max() == 1.0 and v.min() == 0.0


def test_discretize_oracle_matches_reference_bit_exact(pre):
    D = helpers.edge_depth_frames()
    assert np.array_equal(po.discretize_depth_index(D), pre["dd_idx"])
    e = helpers.edge_values()
    assert np.array_equal(po.discretize_depth_index(e), pre["edge_bins"])
    oh = po.discretize_depth_onehot(D[9])
    assert oh.sum() == D[9].size  # one-hot completeness, base_trainer_with_vo.py:162-163


@pytest.mark.parametrize("name,T,N,seed", [("small", 16, 8, 3), ("full", 128, 128, 4)])
@pytest.mark.parametrize("use_gae", [True, False])
def test_gae_oracle_matches_reference_bit_exact(pre, name, T, N, seed, use_gae):
    r, v, m, nv = synth.gae_inputs(T, N, seed)
    ret, _ = po.gae_returns(r, v, m, nv, use_gae, 0.99, 0.95)
    assert np.array_equal(ret, pre[f"gae_{name}_{int(use_gae)}"])


def test_goal_update_known_answers():
    # pure translation / pure rotation known answers for geometry_utils.py:115-144
    out = po.compute_goal_pos(np.array([0.0, 0.0, -2.0]), [0.0, -0.25, 0.0])
    assert np.allclose(out["cartesian"], [0, 0, -1.75]) and np.allclose(out["polar"], [1.75, 0.0])
    out = po.compute_goal_pos(np.array([0.0, 0.0, -1.0]), [0.0, 0.0, np.pi / 2])  # turn left 90deg: goal is to the right
    assert np.allclose(out["cartesian"], [1.0, 0.0, 0.0], atol=1e-12)
    assert np.allclose(out["polar"], [1.0, -np.pi / 2], atol=1e-6)


@pytest.mark.parametrize("case", ["r18_30ch", "r18_8ch", "r50_8ch", "r18_8ch_act_embed", "r18_wider", "r101_deeper"])
def test_vo_oracle_matches_reference(golden_dir, case):
    g = np.load(os.path.join(golden_dir, f"vo_{case}.npz"))
    space, backbone = helpers.VO_CASES[case][1], helpers.VO_CASES[case][2]
    ng = int(g["ngroups"]) if "ngroups" in g.files else 16
    sd = helpers.vo_state_dict(case)
    obs = helpers.vo_inputs(2, 11, space)
    actions = torch.from_numpy(g["actions"]) if "actions" in g.files else None
    with torch.no_grad():
        y, _ = vo.vo_forward(obs, sd, space, backbone, ngroups=ng, training=False, actions=actions)
    assert np.allclose(y.numpy(), g["eval_out"], rtol=1e-5, atol=2e-6)
    y, st = vo.vo_forward(obs, sd, space, backbone, ngroups=ng, training=True, actions=actions)
    assert np.allclose(y.detach().numpy(), g["train_out"], rtol=1e-5, atol=2e-6)
    assert np.allclose(st[0].numpy(), g["train_mean"], rtol=1e-6, atol=1e-7)
    assert np.allclose(st[1].numpy(), g["train_var"], rtol=1e-6, atol=1e-7)
    assert float(st[2]) == float(g["train_count"])


def test_geo_inverse_loss_is_zero_on_consistent_pairs():
    # inversion loss evaluated on ground truth must be ~0
    rng = np.random.default_rng(0)
    n = 16
    a = torch.from_numpy(rng.normal(0, 0.2, size=(n, 3)).astype(np.float32))
    yaw_b = -a[:, 2]
    c, s = torch.cos(yaw_b), torch.sin(yaw_b)
    pb = -torch.stack((c * a[:, 0] + s * a[:, 1], -s * a[:, 0] + c * a[:, 1]), 1)
    b = torch.cat((pb, yaw_b[:, None]), 1)
    deltas = torch.stack((a, b), 1).reshape(2 * n, 3)
    actions = torch.from_numpy(rng.integers(1, 4, size=2 * n))
    assert float(vo.geo_invariance_inverse_loss(deltas, actions)) < 1e-12


@pytest.mark.parametrize("name", ["plain", "types_only", "joint"])
def test_total_loss_oracle_matches_reference_composition(golden_dir, name):
    """vo_total_loss == the reference's _compute_loss / _compute_geo_invariance_inverse_loss composed as
    _process_one_batch does (vo_cnn_regression_geo_invariance_engine.py:676-792): value and gradient."""
    g = np.load(os.path.join(golden_dir, "vo_losses.npz"))
    pred = torch.from_numpy(g[f"{name}/pred"]).requires_grad_(True)
    types = torch.from_numpy(g[f"{name}/data_types"]) if f"{name}/data_types" in g.files else None
    loss = vo.vo_total_loss(pred, torch.from_numpy(g[f"{name}/target"]), torch.from_numpy(g[f"{name}/actions"]), types,
                            tuple(float(w) for w in g["loss_weights"]), torch.from_numpy(g[f"{name}/dz_mask"])[:, None],
                            float(g[f"{name}/inv_w"]))
    loss.backward()
    assert abs(loss.item() - float(g[f"{name}/loss"])) <= 1e-6 * abs(float(g[f"{name}/loss"]))
    assert np.allclose(pred.grad.numpy(), g[f"{name}/grad"], rtol=1e-5, atol=1e-8)


def test_total_loss_oracle_rejects_misaligned_pairs():
    # the reference asserts the [cur_rel_to_prev, prev_rel_to_cur] alternation of the TURN rows (:373-374)
    pred = torch.zeros(4, 3)
    with pytest.raises(AssertionError):
        vo.vo_total_loss(pred, pred, torch.tensor([2, 2, 3, 3]), torch.tensor([0, 0, 1, 1]), loss_inv_weight=1.0)


def test_ppo_loss_oracle_matches_reference_first_minibatch(golden_dir):
    """ppo_losses on the tensors the reference's PPO.update fed to its loss (first minibatch) reproduces the total loss
    it handed to backward (ppo.py:86-135)."""
    g = np.load(os.path.join(golden_dir, "ppo_update.npz"))
    t = {k[4:]: torch.from_numpy(np.asarray(g[k])) for k in g.files if k.startswith("mb0/")}
    vl, al, ent = vo.ppo_losses(t["values"], t["log_probs"], t["entropy"], t["value_preds"], t["returns"],
                                t["old_log_probs"], t["adv"], 0.2, True)
    total = (vl * 0.5 + al - ent * 0.01).item()
    assert abs(total - float(g["mb0/total"])) <= 1e-6 * abs(float(g["mb0/total"]))
