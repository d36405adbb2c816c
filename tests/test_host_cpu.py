"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/pnvo.h declares,
the host mirror has the reference's module surface (registry names, state_dict keys and shapes), the
kernel launch planner accepts every layer of the supported networks, and the data-parallel host logic
works over gloo with world_size 2."""
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="session")
def lib():
    import __graft_entry__ as g

    g.build()
    from pointnav_vo_b200 import lib as L

    return L


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "pnvo.h")).read()
    declared = set(re.findall(r"\b(pnvo_[a-z0-9_]+)\s*\(", header))
    declared -= {"pnvo_op", "pnvo_topdown_consts", "pnvo_opcode"}
    handle = lib.load()
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} declared in include/pnvo.h but not exported"
    assert set(lib.EXPORTS) <= declared
    assert handle.pnvo_abi_version() == 1


def test_product_path_fails_loudly_without_gpu(lib):
    from pointnav_vo_b200.utils import geometry_utils as gu

    with pytest.raises(lib.PnvoError):
        gu.discretize_depth(torch.zeros(4, 4))  # CPU tensor: no fallback


def test_registry_names():
    from pointnav_vo_b200.vo.models import vo_cnn, vo_cnn_act_embed  # noqa: F401
    from pointnav_vo_b200.rl.policies import resnet_policy  # noqa: F401
    from pointnav_vo_b200.utils.baseline_registry import baseline_registry as reg

    for n in ["vo_cnn", "vo_cnn_rgb", "vo_cnn_wider", "vo_cnn_deeper", "vo_cnn_rgb_d_dd", "vo_cnn_rgb_d_top_down",
              "vo_cnn_rgb_dd_top_down", "vo_cnn_d_dd_top_down", "vo_cnn_rgb_d_dd_top_down",
              "vo_cnn_discretize_depth_top_down", "vo_cnn_act_embed", "vo_cnn_wider_act_embed"]:
        assert reg.get_vo_model(n) is not None, n
    assert reg.get_policy("resnet_rnn_policy") is not None


@pytest.mark.parametrize("case", ["r18_30ch", "r18_8ch", "r50_8ch", "r18_8ch_act_embed"])
def test_state_dict_keys_and_shapes_match_reference(case, golden_dir):
    from tests import helpers

    g = np.load(os.path.join(golden_dir, f"vo_{case}.npz"))
    shapes = helpers.vo_state_shapes(case)
    assert [str(k) for k in g["keys"]] == list(shapes.keys())
    for k in g.files:
        if k.startswith("grad/"):
            assert tuple(g[k].shape) == shapes[k[5:]], k


def test_policy_state_dict_keys(golden_dir):
    from pointnav_vo_b200.rl.policies.resnet_policy import PointNavResNetPolicy
    from tests.helpers import policy_spaces

    obs_space, act_space = policy_spaces()
    pol = PointNavResNetPolicy(observation_space=obs_space, action_space=act_space, backbone="resnet18",
                               vis_types=["depth"])
    g = np.load(os.path.join(golden_dir, "policy_r18_depth.npz"))
    assert [str(k) for k in g["keys"]] == list(pol.state_dict().keys())
    assert tuple(g["output_shape"]) == tuple(pol.net.visual_encoder.output_shape)


def test_topdown_host_constants_match_oracle():
    from oracle import preproc_oracle as po
    from pointnav_vo_b200.utils.geometry_utils import NormalizedDepth2TopDownViewHabitatTorch as TD

    orc, mine = po.TopDownOracle(), TD(0.1, 10.0, 192, 341, 70)
    assert np.array_equal(orc.ray, mine._ray)
    c = mine._consts
    for a, b in [(orc.min_x, c.min_x), (orc.x_den, c.x_den), (orc.z_den, c.z_den), (orc.depth_scale, c.depth_scale),
                 (orc.depth_off, c.depth_off)]:
        assert np.float32(a) == np.float32(b)


def test_launch_planner_covers_every_layer(lib):
    """conv_plan (no GPU needed) accepts fprop / dgrad / wgrad of every layer of R18-30ch, R50-8ch and the policy."""
    from pointnav_vo_b200.engine import ConvLayer, LinearLayer, RESNET_LAYERS

    def layers(backbone, cin, H, W, comp, fc_hidden=512):
        kind, nb = RESNET_LAYERS[backbone]
        out = [ConvLayer("c1", cin, 32, 7, 7, 2, 3, H, W, need_dgrad=False, cin_pad=max(8, 1 << (cin - 1).bit_length()))]
        h, w = (out[0].OH + 1) // 2, (out[0].OW + 1) // 2
        inpl, exp = 32, (1 if kind == "basic" else 4)
        for li, n in enumerate(nb):
            planes = 32 * 2 ** li
            for b in range(n):
                s = 2 if (b == 0 and li > 0) else 1
                if kind == "basic":
                    cs = [ConvLayer("a", inpl, planes, 3, 3, s, 1, h, w)]
                    cs.append(ConvLayer("b", planes, planes, 3, 3, 1, 1, cs[0].OH, cs[0].OW))
                else:
                    cs = [ConvLayer("a", inpl, planes, 1, 1, 1, 0, h, w), ConvLayer("b", planes, planes, 3, 3, s, 1, h, w)]
                    cs.append(ConvLayer("c", planes, planes * 4, 1, 1, 1, 0, cs[1].OH, cs[1].OW))
                if b == 0 and (s != 1 or inpl != planes * exp):
                    cs.append(ConvLayer("d", inpl, planes * exp, 1, 1, s, 0, h, w))
                out += cs
                h, w, inpl = cs[-2 if len(cs) > 2 and cs[-1].key == "d" else -1].OH, cs[0 if kind == "basic" else 1].OW, planes * exp
                h = cs[0 if kind == "basic" else 1].OH
        out.append(ConvLayer("comp", inpl, comp, 3, 3, 1, 1, h, w))
        out.append(LinearLayer("fc", comp, h, w, out[-1].cout_pad, fc_hidden))
        return out

    n = 0
    for args in [("resnet18", 30, 192, 341, 31), ("resnet50", 8, 192, 341, 31), ("resnet18", 1, 96, 170, 114)]:
        for c in layers(*args):
            for B in (1, 256):
                info = lib.conv_launch_info(c.op_fwd(None, None, B, None, 2, 16))
                assert info["smem_bytes"] <= 200 * 1024 and info["tmem_cols"] <= 512
                if c.need_dgrad and c.s2_classes:   # 3x3 / stride 2: four parity-class convolutions of dy
                    c.wt = torch.zeros(4 * c.nt_total, c.wt_ld, dtype=torch.float16)
                    ops = c.ops_dgrad(None, None, B)
                    assert len(ops) == 4
                    for op in ops:
                        assert lib.conv_launch_info(op)["smem_bytes"] <= 200 * 1024
                elif c.need_dgrad:
                    assert lib.conv_launch_info(c.op_dgrad(None, None, B))["smem_bytes"] <= 200 * 1024
                assert lib.conv_launch_info(c.op_wgrad(None, None, B))["tmem_cols"] <= 512
                n += 1
    assert n > 100


def _ddp_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pointnav_vo_b200.parallel_utils import allreduce_flat_bucket, merge_input_stats

        # flat gradient bucket: SUM all-reduce; the 1/world factor lives in the loss-gradient kernel
        g = torch.full((1000,), float(rank + 1))
        allreduce_flat_bucket(g)
        ok = bool(torch.all(g == sum(range(1, world + 1))))
        # RunningMeanAndVar: packed (sum, sumsq) all-reduce == statistics of the concatenated batch
        torch.manual_seed(rank)
        x = torch.rand(4, 3, 8, 8, dtype=torch.float64) + rank
        packed = torch.stack((x.sum((0, 2, 3)), (x * x).sum((0, 2, 3))), 1).reshape(-1)
        mean, var = merge_input_stats(packed, n_local=4, pix=64)
        xs = [None] * world
        dist.all_gather_object(xs, x)
        full = torch.cat(xs)
        ok &= torch.allclose(mean, full.mean((0, 2, 3))) and torch.allclose(var, full.var((0, 2, 3), unbiased=False))
        # DD-PPO (ddppo.py:18-96): advantage statistics over every rank, rank-0 weights, averaged gradients.  The
        # mixin is host logic; a small torch module stands in for the CUDA actor-critic.
        import types

        from pointnav_vo_b200.rl.ppo.ppo import DDPPO, distributed_mean_and_var

        v = torch.arange(6, dtype=torch.float32) + 10 * rank
        m_all, var_all = distributed_mean_and_var(v)
        allv = torch.cat([torch.arange(6, dtype=torch.float32) + 10 * r for r in range(world)])
        ok &= torch.allclose(m_all, allv.mean()) and torch.allclose(var_all, allv.var(unbiased=False))
        torch.manual_seed(100 + rank)
        net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 1))
        unused = torch.nn.Parameter(torch.ones(2))  # a parameter the loss never reaches (find_unused_parameters)
        net.register_parameter("unused", unused)
        agent = DDPPO(net, clip_param=0.2, ppo_epoch=1, num_mini_batch=1, value_loss_coef=0.5, entropy_coef=0.01,
                      lr=1e-3, eps=1e-5, max_grad_norm=0.5)
        agent.init_distributed()
        w0 = [None] * world
        dist.all_gather_object(w0, net[0].weight.detach().clone())
        ok &= all(torch.equal(w0[0], w) for w in w0)  # every replica starts from rank 0's weights
        xin = torch.full((5, 4), float(rank + 1))
        loss = net(xin).sum()
        net.zero_grad()
        loss.backward()
        local = net[0].weight.grad.clone()
        agent.after_backward(loss)
        grads = [None] * world
        dist.all_gather_object(grads, local)
        ok &= torch.allclose(net[0].weight.grad, sum(grads) / world) and bool(torch.all(net.unused.grad == 0))
        rollouts = types.SimpleNamespace(returns=torch.arange(5.).view(5, 1, 1) + rank,
                                         value_preds=torch.zeros(5, 1, 1))
        adv = agent.get_advantages(rollouts)
        ok &= adv.shape == (4, 1, 1)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_data_parallel_host_logic_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_rollout_storage_generator_matches_per_env_gather():
    """a14 host side: the single-gather recurrent_generator yields exactly what the reference's per-environment
    Python loop stacks (rollout_storage.py:122-211): same shapes, time-major flattening, same permutation use."""
    import torch

    from pointnav_vo_b200.rl.common.rollout_storage import RolloutStorage
    from tests import helpers

    obs_space, act_space = helpers.policy_spaces()
    obs_space.spaces["depth"].shape = (6, 7, 1)  # small frames: this test is about indexing only
    T, N, mb = 5, 6, 3
    rs = RolloutStorage(T, N, obs_space, act_space, 8, num_recurrent_layers=4)
    g = torch.Generator().manual_seed(3)
    for t in range(T):
        rs.insert({k: torch.randn(N, *sp.shape, generator=g) for k, sp in obs_space.spaces.items()},
                  torch.randn(4, N, 8, generator=g), torch.randint(0, 4, (N, 1), generator=g),
                  torch.randn(N, 1, generator=g), torch.randn(N, 1, generator=g), torch.randn(N, 1, generator=g),
                  (torch.rand(N, 1, generator=g) > 0.2).float())
    assert rs.step == T and rs.actions.dtype == torch.long
    adv = torch.randn(T, N, 1, generator=g)
    torch.manual_seed(11)
    got = list(rs.recurrent_generator(adv, mb))
    torch.manual_seed(11)
    perm = torch.randperm(N)
    per = N // mb
    assert len(got) == mb
    for bi, sample in enumerate(got):
        ind = perm[bi * per:(bi + 1) * per]
        obs, hid, act, prev, vp, ret, msk, olp, ad = sample
        want = lambda x: torch.stack([x[:T, i] for i in ind], 1).reshape(T * per, *x.shape[2:])  # noqa: E731
        for k in obs:
            assert torch.equal(obs[k], want(rs.observations[k]))
        assert torch.equal(hid, torch.stack([rs.recurrent_hidden_states[0, :, i] for i in ind], 1))
        for a, b in ((act, rs.actions), (prev, rs.prev_actions), (vp, rs.value_preds), (ret, rs.returns),
                     (msk, rs.masks), (olp, rs.action_log_probs), (ad, adv)):
            assert torch.equal(a, want(b))
    rs.after_update()
    assert rs.step == 0 and torch.equal(rs.masks[0], rs.masks[T])


def test_descriptor_table_layouts_match_the_c_structs():
    """The batched pack / unpack / GN-parameter-gradient launches read device tables of C structs (csrc/elem.cuh,
    static_assert'ed there): the ctypes mirrors must have the same size and field offsets."""
    import ctypes

    from pointnav_vo_b200 import lib as L

    assert ctypes.sizeof(L.PackDesc) == 64 and L.PackDesc.Cout.offset == 24 and L.PackDesc.src_ld.offset == 60
    assert ctypes.sizeof(L.UnpackDesc) == 48 and L.UnpackDesc.Cout.offset == 16 and L.UnpackDesc.dst_ld.offset == 44
    assert ctypes.sizeof(L.GnParamDesc) == 32 and L.GnParamDesc.C.offset == 24
    assert ctypes.sizeof(L.PnvoOp) == 4 + 27 * 4 + 4 * 4 + 10 * 8


def test_reference_checkpoint_envelopes_load(tmp_path):
    """the three checkpoint envelopes of the reference (single-network VO file, joint VO file keyed by
    action id, DD-PPO policy file with the 'actor_critic.' prefix) load into the modules of this package."""
    from pointnav_vo_b200.rl.policies.resnet_policy import PointNavResNetPolicy
    from pointnav_vo_b200.utils import checkpoint as ck
    from pointnav_vo_b200.vo.common.common_vars import ACT_NAME2IDX
    from pointnav_vo_b200.vo.models import vo_cnn
    from tests.helpers import policy_spaces

    def make():
        return vo_cnn.baseline_registry.get_vo_model("vo_cnn")(
            observation_space=["rgb", "depth"], observation_size=(341, 192), hidden_size=512, backbone="resnet18",
            normalize_visual_inputs=True, output_dim=3, dropout_p=0.2)

    torch.manual_seed(0)
    src = {k: make() for k in ("forward", "left", "right")}
    for i, m in enumerate(src.values()):
        with torch.no_grad():
            for p in m.parameters():
                p.add_(0.01 * (i + 1))
            m.visual_encoder.running_mean_and_var._count.fill_(100.0 + i)
    single = str(tmp_path / "act_forward.pth")
    joint = str(tmp_path / "act_left_right_inv_joint.pth")
    torch.save({"model_state": src["forward"].state_dict(), "epoch": 3}, single)
    ck.save_vo_checkpoint(joint, {"left": src["left"], "right": src["right"]}, epoch=7)
    written = torch.load(joint, weights_only=False)
    assert set(written["model_states"]) == {ACT_NAME2IDX["left"], ACT_NAME2IDX["right"]}  # the reference's keys
    dst = {k: make() for k in ("forward", "left", "right")}
    # VO.REGRESS_MODEL.pretrained_ckpt: one path per action, left and right naming the same joint file
    res = ck.load_vo_checkpoint({"forward": single, "left": joint, "right": joint}, dst)
    assert set(res) == {"forward", "left", "right"}
    for k in dst:
        for (n1, t1), (n2, t2) in zip(src[k].state_dict().items(), dst[k].state_dict().items()):
            assert n1 == n2 and torch.equal(t1, t2), (k, n1)
    with pytest.raises(KeyError):
        ck.vo_state_dict_for(written, "forward")
    with pytest.raises(ValueError):
        ck.vo_state_dict_for({"state_dict": {}}, "forward")

    obs_space, act_space = policy_spaces()
    pol = PointNavResNetPolicy(observation_space=obs_space, action_space=act_space, backbone="resnet18", vis_types=["depth"])
    with torch.no_grad():
        for p in pol.parameters():
            p.add_(0.5)
    path = str(tmp_path / "rl_tune_vo.pth")
    torch.save({"state_dict": {"actor_critic." + k: v for k, v in pol.state_dict().items()}, "config": None}, path)
    pol2 = PointNavResNetPolicy(observation_space=obs_space, action_space=act_space, backbone="resnet18", vis_types=["depth"])
    ck.load_policy_checkpoint(path, pol2)
    for (n1, t1), (n2, t2) in zip(pol.state_dict().items(), pol2.state_dict().items()):
        assert n1 == n2 and torch.equal(t1, t2), n1
    pol3 = PointNavResNetPolicy(observation_space=obs_space, action_space=act_space, backbone="resnet18", vis_types=["depth"])
    ck.load_policy_checkpoint(path, pol3, encoder_only=True)
    for (n, t1), (_, t3) in zip(pol.net.visual_encoder.state_dict().items(), pol3.net.visual_encoder.state_dict().items()):
        assert torch.equal(t1, t3), n
    assert not torch.equal(pol.critic.fc.weight, pol3.critic.fc.weight)


def test_pair_map_follows_the_dataset_rules_and_inverse_targets_zero_the_inversion_loss():
    """8f-2: device-side inverse-pair augmentation, host part.  Row selection mirrors
    regression_geo_invariance_iter_dataset.py:290-366; the targets of the swapped rows make the reference's inversion loss
    (oracle, pinned to the reference) vanish."""
    from oracle import vo_oracle as vo
    from pointnav_vo_b200.vo.common.common_vars import MOVE_FORWARD, TURN_LEFT, TURN_RIGHT
    from pointnav_vo_b200.vo.dataset import geo_invariance as gi

    acts = [TURN_LEFT, TURN_RIGHT, MOVE_FORWARD, TURN_LEFT]
    # unified model, no augmentation: one row per pair
    pm = gi.make_pair_map(acts, act_type=-1, geo_invariance_types=())
    assert pm["pair_map"].tolist() == [0, 2, 4, 6] and pm["actions"].tolist() == acts
    # joint left/right training: every turn pair followed by its swapped twin with the opposite action; forward pairs
    # are kept (joint training) but never inverted
    pm = gi.make_pair_map(acts, act_type=[TURN_LEFT, TURN_RIGHT], geo_invariance_types=("inverse_joint_train",))
    assert pm["pair_map"].tolist() == [0, 1, 2, 3, 4, 6, 7]
    assert pm["actions"].tolist() == [TURN_LEFT, TURN_RIGHT, TURN_RIGHT, TURN_LEFT, MOVE_FORWARD, TURN_LEFT, TURN_RIGHT]
    assert pm["data_types"].tolist() == [0, 1, 0, 1, 0, 0, 1]
    # act-specific model with augmentation only: own action as is, the OTHER turn action only as its inverse
    pm = gi.make_pair_map(acts, act_type=TURN_LEFT, geo_invariance_types=("inverse_data_augment_only",))
    assert pm["pair_map"].tolist() == [0, 3, 6] and pm["actions"].tolist() == [TURN_LEFT, TURN_LEFT, TURN_LEFT]
    rng = np.random.default_rng(0)
    d = np.concatenate([rng.normal(0, 0.2, (16, 2)), rng.normal(0, 0.3, (16, 1))], 1).astype(np.float32)
    turn = [TURN_LEFT, TURN_RIGHT] * 8
    pm = gi.make_pair_map(turn, act_type=[TURN_LEFT, TURN_RIGHT], geo_invariance_types=("inverse_joint_train",))
    tg = gi.expand_targets(d, pm)
    assert tg.shape == (32, 3) and np.array_equal(tg[0::2], d)
    loss = vo.geo_invariance_inverse_loss(torch.from_numpy(tg).double(), torch.from_numpy(pm["actions"]))
    assert float(loss) <= 1e-12
    again = gi.inverse_delta_states(gi.inverse_delta_states(d))
    assert np.allclose(again, d, atol=1e-6)


@pytest.mark.parametrize("rnn_type", ["LSTM", "GRU"])
def test_rnn_state_encoder_packed_sequences_match_stepwise_masking(rnn_type):
    """RNNStateEncoder.seq_forward (every env cut at its own resets, one packed-sequence RNN call) == the reference's
    semantics (rnn_state_encoder.py:100-138: carried state multiplied by the step's mask), outputs, final state and
    input gradient; also with no reset at all and with resets at t = 0."""
    from pointnav_vo_b200.model_utils.rnns.rnn_state_encoder import RNNStateEncoder

    torch.manual_seed(0)
    enc = RNNStateEncoder(24, 16, num_layers=2, rnn_type=rnn_type)
    T, N = 13, 5
    x = torch.randn(T * N, 24, requires_grad=True)
    h = torch.randn(enc.num_recurrent_layers, N, 16)
    xs = x.view(T, N, -1)
    for p_reset in (0.25, 0.0, 1.0):
        masks = (torch.rand(T * N, 1) >= p_reset).float()
        y, hn = enc.seq_forward(x, h, masks)
        hs, outs, ms = enc._unpack_hidden(h), [], masks.view(T, N, 1)
        for t in range(T):
            o, hs = enc.rnn(xs[t:t + 1], enc._mask_hidden(hs, ms[t:t + 1]))
            outs.append(o)
        yr, hr = torch.cat(outs, 0).view(T * N, -1), enc._pack_hidden(hs)
        assert torch.allclose(y, yr, atol=1e-6) and torch.allclose(hn, hr, atol=1e-6)
        g1, = torch.autograd.grad(y.sum() + hn.sum(), x, retain_graph=True)
        g2, = torch.autograd.grad(yr.sum() + hr.sum(), x)
        assert torch.allclose(g1, g2, atol=1e-6)


def test_peer_slices_tile_the_bucket():
    """The per-rank slices of the fused reduce-scatter + Adam + all-gather kernel (parallel_utils.peer_slice mirrors
    csrc/peer_reduce.cu) are 16-byte aligned, disjoint and cover the padded bucket for every world size the kernel takes."""
    from pointnav_vo_b200.parallel_utils import peer_slice

    for n in (4, 8, 36, 3962308, 7235844, 1000):
        n_pad = (n + 3) // 4 * 4
        for world in range(2, 9):
            spans = [peer_slice(n_pad, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n_pad
            for (lo, hi), (lo2, _) in zip(spans, spans[1:] + [(n_pad, n_pad)]):
                assert lo % 4 == 0 and hi % 4 == 0 and lo <= hi == lo2
