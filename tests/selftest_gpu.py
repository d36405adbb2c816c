#!/usr/bin/env python
"""GPU bring-up harness: runs every kernel check in its own subprocess (a hang in one kernel cannot
take the others down), prints error metrics instead of stopping at the first failure.

    python tests/selftest_gpu.py                 # all checks, each under a timeout
    python tests/selftest_gpu.py --check conv_fwd
"""
import argparse
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

CHECKS = {}


def check(fn):
    CHECKS[fn.__name__] = fn
    return fn


def report(name, got, ref, tol=None):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-12
    rms = ref.pow(2).mean().sqrt().item() + 1e-12
    bad = int((~torch.isfinite(got)).sum())
    msg = (f"  {name:34s} max|d|={err.max().item():.3e} mean|d|={err.mean().item():.3e} max|ref|={scale:.3e} "
           f"rms={rms:.3e} rel(max/rms)={err.max().item() / rms:.3e} nonfinite={bad}")
    ok = bad == 0 and (tol is None or err.max().item() <= tol * rms)
    print(msg + ("" if tol is None else ("  OK" if ok else "  **FAIL**")), flush=True)
    return ok


# -------------------------------------------------------------------------------------------------
@check
def preproc():
    from oracle import preproc_oracle as po
    from pointnav_vo_b200.utils import geometry_utils as gu
    from pointnav_vo_b200.rl.common.rollout_returns import compute_returns
    from pointnav_vo_b200.utils import synth
    from tests import helpers

    D = helpers.edge_depth_frames()
    d = torch.from_numpy(D).cuda()
    idx = gu.discretize_depth_index(d).cpu().numpy()
    print("  discretize index bit-exact:", np.array_equal(idx, po.discretize_depth_index(D)))
    oh = gu.discretize_depth(d[:4]).cpu().numpy()
    print("  discretize one-hot bit-exact:", np.array_equal(oh, po.discretize_depth_onehot(D[:4])))
    gen = gu.NormalizedDepth2TopDownViewHabitatTorch(0.1, 10.0, 192, 341, 70)
    out, cnt = gen.gen_top_down_view(d[..., None], return_counts=True)
    orc = po.TopDownOracle()
    okc = oko = True
    for i in range(D.shape[0]):
        c = orc.count_map(D[i])
        if not np.array_equal(c, cnt[i].cpu().numpy()):
            okc = False
            print("   frame", i, "count mismatch:", int((c != cnt[i].cpu().numpy()).sum()), "cells")
        if not np.array_equal(orc.gen_top_down_view(D[i])[..., 0], out[i, ..., 0].cpu().numpy()):
            oko = False
    print("  top-down counts bit-exact:", okc, " maps bit-exact:", oko)
    for use_gae in (True, False):
        r, v, m, nv = synth.gae_inputs(128, 128, 4)
        ref, _ = po.gae_returns(r, v, m, nv, use_gae, 0.99, 0.95)
        for mode in ("exact", "scan"):
            tr, tv, tm, tn = (torch.from_numpy(a.copy()).cuda() for a in (r, v, m, nv))
            ret = torch.zeros_like(tv)
            compute_returns(tr, tv, tm, tn, ret, use_gae, 0.99, 0.95, mode=mode)
            g = ret.cpu().numpy()
            print(f"  gae use_gae={use_gae} mode={mode}: bit-exact={np.array_equal(g, ref)} max|d|={np.abs(g - ref).max():.3e}")


def _conv_case(name, B, Cin, Cout, R, S, stride, pad, IH, IW, seed=0, stats_groups=16, test_bwd=True, cin_pad=None):
    from pointnav_vo_b200 import lib as L
    from pointnav_vo_b200.engine import ConvLayer

    torch.manual_seed(seed)
    dev = "cuda"
    c = ConvLayer("w", Cin, Cout, R, S, stride, pad, IH, IW, need_dgrad=test_bwd, cin_pad=cin_pad)
    c.alloc(dev, True)
    w = (torch.randn(Cout, Cin, R, S, device=dev) / (Cin * R * S) ** 0.5).contiguous()
    x = torch.randn(B, IH, IW, c.cin_pad, device=dev).half()
    x[..., Cin:] = 0
    G = stats_groups
    cpg = c.cout_pad // G
    stats = torch.zeros(B, G, 2, device=dev, dtype=torch.float64)
    y = torch.empty(B, c.OH, c.OW, c.cout_pad, dtype=torch.float16, device=dev)
    ops = [c.op_pack(w), c.op_fwd(x, y, B, stats, cpg, G)]
    info = L.conv_launch_info(ops[1])
    L.run_ops(ops)
    torch.cuda.synchronize()
    xr = x[..., :Cin].float().permute(0, 3, 1, 2)
    wr = w.half().float()
    ref = F.conv2d(xr, wr, None, stride, pad)
    ok = report(f"{name} fwd {info}", y[..., :Cout].permute(0, 3, 1, 2), ref, 3e-3)
    if c.cout_pad > Cout:
        print("   pad channels zero:", bool((y[..., Cout:] == 0).all()))
    if Cout % G == 0 and c.cout_pad == Cout:
        rs = ref.reshape(B, G, -1)
        sref = torch.stack((rs.sum(-1), rs.pow(2).sum(-1)), -1)
        ok &= report(f"{name} GN partial sums", stats, sref, 2e-3)
    if not test_bwd:
        # wgrad only
        dy = torch.randn(B, c.OH, c.OW, c.cout_pad, device=dev).half()
        dy[..., Cout:] = 0
        gw = torch.zeros(Cout, Cin, R, S, device=dev)
        L.run_ops([L.op_zero(c.dwp), c.op_wgrad(x, dy, B), c.op_unpack(gw)])
        torch.cuda.synchronize()
        wr2 = wr.clone().requires_grad_(True)
        F.conv2d(xr, wr2, None, stride, pad).backward(dy[..., :Cout].float().permute(0, 3, 1, 2))
        ok &= report(f"{name} wgrad", gw, wr2.grad, 3e-3)
        return ok
    dy = torch.randn(B, c.OH, c.OW, c.cout_pad, device=dev).half()
    dy[..., Cout:] = 0
    gx = torch.empty(B, IH, IW, c.cin_pad, dtype=torch.float16, device=dev)
    gw = torch.zeros(Cout, Cin, R, S, device=dev)
    L.run_ops([L.op_zero(c.dwp), c.op_dgrad(dy, gx, B), c.op_wgrad(x, dy, B), c.op_unpack(gw)])
    torch.cuda.synchronize()
    xr2 = xr.clone().requires_grad_(True)
    wr2 = wr.clone().requires_grad_(True)
    F.conv2d(xr2, wr2, None, stride, pad).backward(dy[..., :Cout].float().permute(0, 3, 1, 2))
    ok &= report(f"{name} dgrad", gx[..., :Cin].permute(0, 3, 1, 2), xr2.grad, 3e-3)
    ok &= report(f"{name} wgrad", gw, wr2.grad, 3e-3)
    # dgrad with accumulate input
    add = torch.randn_like(gx)
    gx2 = torch.empty_like(gx)
    L.run_ops([c.op_dgrad(dy, gx2, B, add=add)])
    torch.cuda.synchronize()
    ok &= report(f"{name} dgrad+add", gx2[..., :Cin].permute(0, 3, 1, 2),
                 xr2.grad + add[..., :Cin].float().permute(0, 3, 1, 2), 3e-3)
    return ok


@check
def conv_small():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    _conv_case("3x3s1 32->32 tiny", 1, 32, 32, 3, 3, 1, 1, 8, 16)
    _conv_case("1x1s1 64->64 tiny", 2, 64, 64, 1, 1, 1, 0, 8, 8)


@check
def conv_layers():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    _conv_case("layer1 3x3s1 32->32", 3, 32, 32, 3, 3, 1, 1, 48, 86)
    _conv_case("layer2.0 3x3s2 32->64", 3, 32, 64, 3, 3, 2, 1, 48, 86)
    _conv_case("layer2.0 down 1x1s2 32->64", 3, 32, 64, 1, 1, 2, 0, 48, 86)
    _conv_case("layer3 3x3s1 128->128", 3, 128, 128, 3, 3, 1, 1, 12, 22)
    _conv_case("layer4 3x3s1 256->256", 3, 256, 256, 3, 3, 1, 1, 6, 11)
    _conv_case("layer4.0 3x3s2 128->256", 3, 128, 256, 3, 3, 2, 1, 12, 22)
    _conv_case("compression 256->31", 3, 256, 31, 3, 3, 1, 1, 6, 11, stats_groups=1)


@check
def conv_stem_fc():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    _conv_case("conv1 7x7s2 30->32", 2, 30, 32, 7, 7, 2, 3, 192, 341, test_bwd=False)
    _conv_case("policy conv1 7x7s2 1->32", 2, 1, 32, 7, 7, 2, 3, 96, 170, test_bwd=False)
    _conv_case("r50 1x1 512->128", 2, 512, 128, 1, 1, 1, 0, 12, 22)
    _conv_case("r50 1x1 256->1024", 2, 256, 1024, 1, 1, 1, 0, 6, 11)
    _conv_case("fc as 1x1 2112->512", 64, 2112, 512, 1, 1, 1, 0, 1, 1, cin_pad=2112)


@check
def gn_ops():
    from pointnav_vo_b200 import lib as L

    torch.manual_seed(0)
    dev = "cuda"
    for (B, H, W, C, G, Cr) in [(3, 24, 43, 64, 16, 64), (2, 6, 11, 32, 1, 31), (2, 12, 22, 128, 16, 128)]:
        x = torch.randn(B, H, W, C, device=dev).half()
        x[..., Cr:] = 0
        res = torch.randn(B, H, W, C, device=dev).half()
        gamma = torch.rand(Cr, device=dev) + 0.5
        beta = torch.randn(Cr, device=dev) * 0.1
        xf = x[..., :Cr].float().permute(0, 3, 1, 2)
        xs = xf.reshape(B, G, -1)
        stats = torch.stack((xs.sum(-1), xs.pow(2).sum(-1)), -1).double().contiguous()
        y = torch.empty_like(x)
        cpg, cpg_r = C // G, Cr // G
        HW = H * W
        L.run_ops([L.op_gn_apply(x, stats, gamma, beta, y, B, C, G, cpg, HW, float(cpg_r * HW), True, res, False, 1e-5, Cr)])
        torch.cuda.synchronize()
        xr = xf.clone().requires_grad_(True)
        gr = gamma.clone().requires_grad_(True)
        br = beta.clone().requires_grad_(True)
        ref = F.relu(F.group_norm(xr, G, gr, br, 1e-5) + res[..., :Cr].float().permute(0, 3, 1, 2))
        report(f"gn_apply C={C} G={G}", y[..., :Cr].permute(0, 3, 1, 2), ref, 3e-3)
        g = torch.randn(B, H, W, C, device=dev).half()
        g[..., Cr:] = 0
        ref.backward(g[..., :Cr].float().permute(0, 3, 1, 2))
        sums = torch.zeros(B, C, 2, device=dev)
        dx = torch.empty_like(x)
        dyo = torch.empty_like(x)
        dg = torch.zeros(Cr, device=dev)
        db = torch.zeros(Cr, device=dev)
        yref16 = ref.detach().permute(0, 2, 3, 1).half()
        yfull = torch.zeros_like(x)
        yfull[..., :Cr] = yref16
        args = (g, yfull, x, stats, gamma, sums, dx, dyo, B, C, G, cpg, HW, float(cpg_r * HW), False, 1e-5, Cr)
        L.run_ops([L.op_gn_bwd(True, *args), L.op_gn_param_grad(sums, dg, db, B, C, Cr), L.op_gn_bwd(False, *args)])
        torch.cuda.synchronize()
        report(f"gn_bwd dx C={C} G={G}", dx[..., :Cr].permute(0, 3, 1, 2), xr.grad, 5e-3)
        report(f"gn_bwd dgamma C={C}", dg, gr.grad, 5e-3)
        report(f"gn_bwd dbeta C={C}", db, br.grad, 5e-3)
    # GN + ReLU + maxpool and its backward
    B, H, W, C, G = 2, 96, 171, 32, 16
    x = torch.randn(B, H, W, C, device=dev).half()
    gamma = torch.rand(C, device=dev) + 0.5
    beta = torch.randn(C, device=dev) * 0.1
    xf = x.float().permute(0, 3, 1, 2)
    xs = xf.reshape(B, G, -1)
    stats = torch.stack((xs.sum(-1), xs.pow(2).sum(-1)), -1).double().contiguous()
    PH, PW = 48, 86
    y = torch.empty(B, PH, PW, C, dtype=torch.float16, device=dev)
    am = torch.empty(B, PH, PW, C, dtype=torch.uint8, device=dev)
    L.run_ops([L.op_gn_pool(x, stats, gamma, beta, y, am, B, C, G, C // G, H, W, PH, PW, float(C // G * H * W))])
    torch.cuda.synchronize()
    a = F.relu(F.group_norm(xf, G, gamma, beta, 1e-5)).requires_grad_(True)
    ref = F.max_pool2d(a, 3, 2, 1)
    report("gn_pool", y.permute(0, 3, 1, 2), ref, 3e-3)
    gp = torch.randn(B, PH, PW, C, device=dev).half()
    ref.backward(gp.float().permute(0, 3, 1, 2))
    dy = torch.empty(B, H, W, C, dtype=torch.float16, device=dev)
    L.run_ops([L.op_pool_bwd(gp, y, am, dy, B, C, H, W, PH, PW)])
    torch.cuda.synchronize()
    mask = (a.detach() > 0).float()
    report("pool_bwd (x relu mask)", dy.permute(0, 3, 1, 2), a.grad * mask, 3e-3)


def _load_vo(case, dropout_p=0.0):
    from pointnav_vo_b200.vo.models import vo_cnn
    from tests import helpers

    name, space, backbone, kw = helpers.VO_CASES[case]
    cls = vo_cnn.VisualOdometryCNNBase if name == "base" else vo_cnn.baseline_registry.get_vo_model(name)
    m = cls(observation_space=space, observation_size=(341, 192), hidden_size=512, backbone=backbone,
            normalize_visual_inputs=True, output_dim=3, dropout_p=dropout_p, **kw)
    m.load_state_dict(helpers.vo_state_dict(case))
    return m.cuda(), space, backbone


@check
def vo_model():
    from oracle import vo_oracle as vo
    from tests import helpers

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    for case in ("r18_30ch", "r18_8ch", "r50_8ch"):
        g = np.load(os.path.join(ROOT, "tests", "golden", f"vo_{case}.npz"))
        m, space, backbone = _load_vo(case)
        obs = helpers.vo_inputs(2, 11, space, "cuda")
        m.eval()
        with torch.no_grad():
            y = m(obs)
        torch.cuda.synchronize()
        ref = torch.from_numpy(g["eval_out"])
        print(f" [{case}] eval out:\n   got {y.cpu().numpy().ravel()}\n   ref {ref.numpy().ravel()}")
        report(f"{case} eval", y, ref, 1e-3)
        # per-layer taps against the oracle run on the GPU in fp32
        sd = {k: v.cuda() for k, v in helpers.vo_state_dict(case).items()}
        taps = {}
        with torch.no_grad():
            vo.vo_forward(obs, sd, space, backbone, training=False, taps=taps)
        plan = list(m._plans.values())[0]
        C = m.visual_encoder.input_channels
        report(f"{case} tap input", plan.x0_img[:, :, :plan.W, :C].permute(0, 3, 1, 2), taps["input"], 2e-3)
        report(f"{case} tap conv1_raw", plan.raw1.permute(0, 3, 1, 2), taps["conv1_raw"], 3e-3)
        report(f"{case} tap pool", plan.pool.permute(0, 3, 1, 2), taps["pool"], 3e-3)
        li = 0
        for bi, blk in enumerate(plan.blocks):
            nxt = plan.blocks[bi + 1]["name"] if bi + 1 < len(plan.blocks) else None
            if nxt is None or nxt.split(".")[-2] != blk["name"].split(".")[-2]:
                li += 1
                report(f"{case} tap layer{li}", blk["y"].permute(0, 3, 1, 2), taps[f"layer{li}"], 5e-3)
        cc = m.visual_encoder.output_shape[0]
        report(f"{case} tap compression", plan.feat[..., :cc].permute(0, 3, 1, 2), taps["compression"], 5e-3)
        # training forward + backward
        m.train()
        target = torch.from_numpy(g["target"]).cuda()
        y = m(obs)
        loss = sum(vo.vo_losses(y, target))
        loss.backward()
        torch.cuda.synchronize()
        # gradient taps: oracle autograd (fp32, GPU) vs the plan's gradient buffers
        sdg = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
        taps = {}
        vo.QUANT[0] = os.environ.get("PNVO_QUANT_ORACLE", "1") == "1"  # emulate fp16 storage in the oracle
        yo, _ = vo.vo_forward(obs, sdg, space, backbone, training=True, taps=taps)
        vo.QUANT[0] = False
        for t in taps.values():
            if t.requires_grad:
                t.retain_grad()
        sum(vo.vo_losses(yo, target)).backward()
        plan = [p for p in m._plans.values() if p.training][0]
        report(f"{case} train out vs gpu oracle", y, yo.detach(), 1e-3)
        report(f"{case} gtap compression", plan.g_feat[..., :cc].permute(0, 3, 1, 2), taps["compression"].grad, 1e-2)
        li = 0
        for bi, blk in enumerate(plan.blocks):
            nxt = plan.blocks[bi + 1]["name"] if bi + 1 < len(plan.blocks) else None
            if nxt is None or nxt.split(".")[-2] != blk["name"].split(".")[-2]:
                li += 1
                report(f"{case} gtap layer{li}", blk["g_y"].permute(0, 3, 1, 2), taps[f"layer{li}"].grad, 1e-2)
        report(f"{case} gtap pool", plan.g_pool.permute(0, 3, 1, 2), taps["pool"].grad, 1e-2)
        report(f"{case} gtap conv1_raw", plan.dx1.permute(0, 3, 1, 2), taps["conv1_raw"].grad, 1e-2)
        Pm = dict(m.named_parameters())
        for k in ("output_head.1.weight", "visual_fc.2.bias", "visual_fc.2.weight", "visual_encoder.compression.0.weight",
                  "visual_encoder.compression.1.weight", "visual_encoder.backbone.conv1.0.weight"):
            report(f"{case} grad {k[-40:]}", Pm[k].grad, sdg[k].grad, 1e-2)
        report(f"{case} train out", y, torch.from_numpy(g["train_out"]), 1e-3)
        sdn = m.state_dict()
        report(f"{case} rmv mean", sdn["visual_encoder.running_mean_and_var._mean"], torch.from_numpy(g["train_mean"]), 1e-4)
        report(f"{case} rmv var", sdn["visual_encoder.running_mean_and_var._var"], torch.from_numpy(g["train_var"]), 1e-4)
        print("   count", float(sdn["visual_encoder.running_mean_and_var._count"]), float(g["train_count"]))
        gk = [str(k) for k in g["grad_keys"]]
        gn = g["grad_norms"]
        P = dict(m.named_parameters())
        worst = 0.0
        for k, n in zip(gk, gn):
            mine = P[k].grad.norm().item()
            rel = abs(mine - n) / max(n, 1e-12)
            worst = max(worst, rel)
            if rel > 2e-2:
                print(f"   grad-norm mismatch {k}: {mine:.4e} vs {n:.4e}")
        print(f"   worst grad-norm rel err over {len(gk)} params: {worst:.3e}")
        for k in g.files:
            if k.startswith("grad/"):
                kk = k[5:]
                if P[kk].grad.numel() > 64:
                    report(f"{case} {kk}", P[kk].grad, torch.from_numpy(g[k]), 2e-2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", default=None)
    ap.add_argument("--timeout", type=int, default=240)
    args = ap.parse_args()
    if args.check:
        t0 = time.time()
        CHECKS[args.check]()
        torch.cuda.synchronize()
        print(f"[{args.check}] done in {time.time() - t0:.1f}s", flush=True)
        return
    for name in CHECKS:
        print(f"===== {name} =====", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--check", name], timeout=args.timeout,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            print(r.stdout[-60000:], flush=True)
            print(f"[{name}] exit code {r.returncode}", flush=True)
        except subprocess.TimeoutExpired as e:
            out = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            print(out[-6000:])
            print(f"[{name}] TIMEOUT after {args.timeout}s", flush=True)



@check
def vo_blocks():
    """Backward logic, block by block: the oracle block is fed with the plan's OWN activations and upstream
    gradient, so only one block's rounding separates the two (no accumulated fp16 noise, few ReLU flips)."""
    from oracle import vo_oracle as vo
    from tests import helpers

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    for case in ("r18_8ch", "r50_8ch"):
        g = np.load(os.path.join(ROOT, "tests", "golden", f"vo_{case}.npz"))
        m, space, backbone = _load_vo(case)
        obs = helpers.vo_inputs(2, 11, space, "cuda")
        m.train()
        target = torch.from_numpy(g["target"]).cuda()
        y = m(obs)
        sum(vo.vo_losses(y, target)).backward()
        torch.cuda.synchronize()
        plan = [p for p in m._plans.values() if p.training][0]
        P = dict(m.named_parameters())
        ng = m.visual_encoder.ngroups
        kind = vo.RESNET_LAYERS[backbone][0]
        fn = vo._basic_block if kind == "basic" else vo._bottleneck

        def nchw(t, c=None):
            t = t.float().permute(0, 3, 1, 2)
            return t if c is None else t[:, :c]

        # head + compression
        sd = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
        x4 = nchw(plan.blocks[-1]["y"]).clone().requires_grad_(True)
        feat = vo.encoder_tail(x4, sd, "visual_encoder")
        h = F.relu(F.linear(feat.flatten(1), sd["visual_fc.2.weight"], sd["visual_fc.2.bias"]))
        out = F.linear(h, sd["output_head.1.weight"], sd["output_head.1.bias"])
        sum(vo.vo_losses(out, target)).backward()
        report(f"{case} head: out", y, out.detach(), 2e-3)
        report(f"{case} head: g layer4", nchw(plan.blocks[-1]["g_y"]), x4.grad, 1e-2)
        for k in ("output_head.1.weight", "output_head.1.bias", "visual_fc.2.weight", "visual_fc.2.bias",
                  "visual_encoder.compression.0.weight", "visual_encoder.compression.1.weight",
                  "visual_encoder.compression.1.bias"):
            report(f"{case} head: {k[-36:]}", P[k].grad, sd[k].grad, 1e-2)
        # residual blocks
        for bi in range(len(plan.blocks) - 1, -1, -1):
            blk = plan.blocks[bi]
            p = blk["name"].replace("visual_encoder.backbone.", "")
            sd = {k: v.detach().clone().requires_grad_(True) for k, v in P.items() if blk["name"] in k}
            xin = nchw(blk["x_in"]).clone().requires_grad_(True)
            stride = blk["convs"][0 if kind == "basic" else 1].stride
            yb = fn(xin, sd, blk["name"], ng, stride, blk["down"] is not None)
            report(f"{case} {p}: y", nchw(blk["y"]), yb.detach(), 5e-3)
            yb.backward(nchw(blk["g_y"]))
            gx = plan.blocks[bi - 1]["g_y"] if bi > 0 else plan.g_pool
            report(f"{case} {p}: g_x", nchw(gx), xin.grad, 1e-2)
            for k in sd:
                report(f"{case} {p}: {k.split(p)[-1]}", P[k].grad, sd[k].grad, 1e-2)
        # stem
        C = m.visual_encoder.input_channels
        sd = {k: v.detach().clone().requires_grad_(True) for k, v in P.items() if ".conv1." in k}
        x0 = nchw(plan.x0_img[:, :, :plan.W], C)
        pfx = "visual_encoder.backbone"
        a = F.conv2d(x0, sd[pfx + ".conv1.0.weight"], None, 2, 3)
        a = F.relu(F.group_norm(a, ng, sd[pfx + ".conv1.1.weight"], sd[pfx + ".conv1.1.bias"], 1e-5))
        pl = F.max_pool2d(a, 3, 2, 1)
        report(f"{case} stem: pool", nchw(plan.pool), pl.detach(), 5e-3)
        pl.backward(nchw(plan.g_pool))
        for k in sd:
            report(f"{case} stem: {k[-20:]}", P[k].grad, sd[k].grad, 1e-2)


if __name__ == "__main__":
    main()
