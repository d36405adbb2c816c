"""Launched by torchrun (one rank per GPU): the fused peer-memory gradient exchange (csrc/peer_reduce.cu, reduce-scatter +
Adam + all-gather in one kernel over NVLink) against the NCCL all-reduce + adam_kernel path it replaces, on the same
model, data and steps.  Checks, after every step: (1) the replicas are bit-identical across ranks on the peer path;
(2) parameters of the two paths agree to fp32 summation-order noise; (3) the optimiser state round-trips through
state_dict() in torch.optim.Adam's format.  Prints PEER_ADAM_OK on rank 0."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests import helpers  # noqa: E402


def build(case, peer):
    from pointnav_vo_b200.vo.engine.train_step import FusedVOTrainStep
    from pointnav_vo_b200.vo.models import vo_cnn

    name, space, backbone, kw = helpers.VO_CASES[case]
    m = vo_cnn.baseline_registry.get_vo_model(name)(
        observation_space=space, observation_size=(341, 192), hidden_size=512, backbone=backbone,
        normalize_visual_inputs=True, output_dim=3, dropout_p=0.0, **kw)
    m.load_state_dict(helpers.vo_state_dict(case))
    m = m.cuda().train()
    t = FusedVOTrainStep(m, lr=2.5e-4, eps=1e-8)
    t.use_peer_exchange = bool(peer)   # (default: on when PNVO_PEER_ADAM != 0)
    return m, space, t


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # (0) the small fp64 exchange (RunningMeanAndVar statistics) against NCCL, repeatedly (slot parities, sequence flags)
    from pointnav_vo_b200.vo.models import vo_cnn as vc

    for it in range(6):
        t = torch.randn(66, dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(1000 * it + rank))
        a, b = t.clone(), t.clone()
        vc._allreduce_stats(a)
        dist.all_reduce(b)
        assert vc._PEER_SUM[torch.cuda.current_device()], "peer statistics exchange not active"
        assert torch.allclose(a, b, rtol=1e-14, atol=1e-14), (it, (a - b).abs().max().item())
        ref = a.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref, a), "statistics differ between the ranks"
    assert not vc._PEER_SUM[torch.cuda.current_device()].timed_out()
    case = "r18_8ch"
    m_peer, space, t_peer = build(case, True)
    m_nccl, _, t_nccl = build(case, False)
    B = 2
    for step in range(4):
        obs = helpers.vo_inputs(B, 100 + 7 * step + rank, space, "cuda")   # different data on every rank
        tgt = torch.randn(B, 3, device="cuda", generator=torch.Generator("cuda").manual_seed(step * 11 + rank)) * 0.1
        l1 = t_peer.step(obs, tgt)
        l2 = t_nccl.step(obs, tgt)
        torch.cuda.synchronize()
        assert t_peer._peer is not None, "peer path not active"
        assert t_nccl._peer is None
        assert not t_peer._peer.timed_out(), "peer flags timed out"
        flat = t_peer._flat
        # (1) replicas bit-identical
        ref = flat.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref, flat), f"rank {rank}: replica differs from rank 0 after step {step}"
        # (2) the two paths agree
        a, b = t_peer._flat, t_nccl._flat
        d = (a - b).abs().max().item()
        scale = b.abs().max().item()
        # Adam's first steps move a weight by ~lr * sign(g): a gradient within summation-order noise of zero may differ
        frac = ((a - b).abs() > 1e-6).float().mean().item()
        assert d <= 2.2 * 2.5e-4 * (step + 1) and frac <= (0.02 if step == 0 else 0.15), (step, d, frac, scale)
        assert abs(l1.item() - l2.item()) <= 1e-2 * abs(l2.item()) + 1e-6
    # running input statistics: identical on every rank (fixed summation order), equal between the two models
    for m in (m_peer, m_nccl):
        rmv = m.visual_encoder.running_mean_and_var
        for buf in (rmv._mean, rmv._var, rmv._count):
            ref = buf.clone()
            dist.broadcast(ref, 0)
            assert torch.equal(ref, buf), "running statistics differ between the ranks"
    assert torch.equal(m_peer.visual_encoder.running_mean_and_var._mean, m_nccl.visual_encoder.running_mean_and_var._mean)
    # (3) state_dict: full moments equal the NCCL path's (same tolerance class), and reload keeps the owned slice
    sd = t_peer.state_dict()
    sd2 = t_nccl.state_dict()
    assert sd["state"][0]["step"].item() == 4.0
    num = sum((sd["state"][i]["exp_avg"] - sd2["state"][i]["exp_avg"]).pow(2).sum().item() for i in sd["state"])
    den = sum(sd2["state"][i]["exp_avg"].pow(2).sum().item() for i in sd2["state"])
    assert num <= (5e-2 ** 2) * den, (num, den)
    m_before, v_before = t_peer._m.clone(), t_peer._v.clone()
    t_peer.load_state_dict(sd)
    assert torch.equal(m_before, t_peer._m) and torch.equal(v_before, t_peer._v)
    opt = torch.optim.Adam(m_peer.parameters(), lr=2.5e-4, eps=1e-8)
    opt.load_state_dict(sd)   # the format torch.optim.Adam reads
    dist.barrier()
    if rank == 0:
        print(f"PEER_ADAM_OK world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
