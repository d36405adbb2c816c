"""GPU parity tests (run on the B200 box with `-m gpu`).  Everything goes through the C ABI (libpnvo.so via
pointnav_vo_b200.lib); the checker is the oracle (oracle/*.py, pinned against the unmodified reference) and
the committed golden fixtures produced by the reference itself (tests/golden/make_golden.py).

Tolerances (also in DESIGN.md "Numerics"):
  * integer / index work (discretisation, top-down counts and maps, exact-mode GAE): bit-exact.
  * a single conv / GroupNorm op against torch fp32 on the same fp16-rounded inputs: max|d| <= 3e-3 * rms(ref)
    (one fp16 rounding of the stored result).
  * whole-network outputs against the fp32 reference, default mode: fp16 operands/activations with fp32 accumulation
    give max|d| <= 8e-3 * rms(ref) for ResNet-18 and 2.5e-2 for ResNet-50 (20 resp. 53 rounded layers).
  * whole-network outputs in split precision (set_precision("split"), no-grad forwards): <= 1e-3 (the north-star
    tolerance; measured 7e-6 .. 6e-5), both as max|d| / rms(ref) and element-wise relative to |ref|.
  * gradients: ReLU masks flip where a pre-activation is within rounding distance of zero, which perturbs
    weight gradients (random-sign sums) by ~sqrt(flip fraction); relative L2 error <= 0.15 per tensor.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import preproc_oracle as po  # noqa: E402
from oracle import vo_oracle as vo  # noqa: E402
from pointnav_vo_b200.utils import synth  # noqa: E402
from tests import helpers  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _setup():
    assert torch.cuda.is_available()
    from pointnav_vo_b200 import lib as L

    L.check(L.load().pnvo_check_device())
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def rel(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert torch.isfinite(got).all()
    return ((got - ref).abs().max() / (ref.pow(2).mean().sqrt() + 1e-12)).item()


def rel_l2(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return ((got - ref).norm() / (ref.norm() + 1e-12)).item()


# ------------------------------------------------------------------------------------------------ a7 / a8 / a13
def test_discretize_bit_exact(golden_dir):
    from pointnav_vo_b200.utils import geometry_utils as gu

    pre = np.load(os.path.join(golden_dir, "preproc.npz"))
    D = helpers.edge_depth_frames()
    d = torch.from_numpy(D).cuda()
    assert np.array_equal(gu.discretize_depth_index(d).cpu().numpy(), pre["dd_idx"])
    e = helpers.edge_values()
    assert np.array_equal(gu.discretize_depth_index(torch.from_numpy(e).cuda()).cpu().numpy(), pre["edge_bins"])
    oh = gu.discretize_depth(d[:6]).cpu().numpy()
    assert np.array_equal(oh, po.discretize_depth_onehot(D[:6]))
    # ragged / empty inputs
    assert gu.discretize_depth(torch.zeros(0, device="cuda")).shape == (0, 10)
    odd = torch.from_numpy(D[9].reshape(-1)[:12345]).cuda()
    assert np.array_equal(gu.discretize_depth_index(odd).cpu().numpy(), po.discretize_depth_index(D[9].reshape(-1)[:12345]))
    with pytest.raises(AssertionError):
        gu.discretize_depth(torch.full((8,), 1.5, device="cuda"))  # the reference asserts 0 <= d <= 1


def test_topdown_bit_exact(golden_dir):
    from pointnav_vo_b200.utils import geometry_utils as gu

    pre = np.load(os.path.join(golden_dir, "preproc.npz"))
    D = helpers.edge_depth_frames()
    gen = gu.NormalizedDepth2TopDownViewHabitatTorch(0.1, 10.0, 192, 341, 70)
    out, cnt = gen.gen_top_down_view(torch.from_numpy(D).cuda()[..., None], return_counts=True)
    orc = po.TopDownOracle()
    for i in range(D.shape[0]):
        assert np.array_equal(out[i, ..., 0].cpu().numpy(), helpers.golden_topdown(pre, i)), i
        assert np.array_equal(cnt[i].cpu().numpy(), orc.count_map(D[i])), i
    single = gen.gen_top_down_view(torch.from_numpy(D[10]).cuda()[..., None])  # reference signature [H, W, 1]
    assert single.shape == (192, 341, 1) and np.array_equal(single[..., 0].cpu().numpy(), helpers.golden_topdown(pre, 10))


@pytest.mark.parametrize("name,T,N,seed", [("small", 16, 8, 3), ("full", 128, 128, 4)])
@pytest.mark.parametrize("use_gae", [True, False])
def test_gae(golden_dir, name, T, N, seed, use_gae):
    from pointnav_vo_b200.rl.common.rollout_returns import compute_returns

    pre = np.load(os.path.join(golden_dir, "preproc.npz"))
    r, v, m, nv = synth.gae_inputs(T, N, seed)
    ref = pre[f"gae_{name}_{int(use_gae)}"]
    for mode in ("exact", "scan"):
        tr, tv, tm, tn = (torch.from_numpy(a.copy()).cuda() for a in (r, v, m, nv))
        ret = torch.zeros_like(tv)
        compute_returns(tr, tv, tm, tn, ret, use_gae, 0.99, 0.95, mode=mode)
        got = ret.cpu().numpy()
        if mode == "exact":
            assert np.array_equal(got, ref)
            if use_gae:
                assert np.array_equal(tv[T].cpu().numpy(), nv)  # value_preds[T] <- next_value, as the reference
        else:
            assert np.allclose(got, ref, rtol=2e-5, atol=2e-5)
    # size-independent property: scaling rewards and values by 2 scales the returns by exactly 2
    tr, tv, tm, tn = (torch.from_numpy(a.copy()).cuda() for a in (2 * r, 2 * v, m, 2 * nv))
    ret = torch.zeros_like(tv)
    compute_returns(tr, tv, tm, tn, ret, use_gae, 0.99, 0.95)
    assert np.array_equal(ret.cpu().numpy(), 2 * ref)


def test_goal_update_matches_oracle():
    from pointnav_vo_b200.utils.geometry_utils import compute_goal_pos_batched

    rng = np.random.default_rng(0)
    g = rng.uniform(-3, 3, size=(33, 3))
    d = rng.normal(0, 0.2, size=(33, 3)).astype(np.float32)
    out = compute_goal_pos_batched(torch.from_numpy(g.copy()).cuda(), torch.from_numpy(d).cuda())
    for i in range(33):
        ref = po.compute_goal_pos(g[i], d[i])
        assert np.allclose(out["cartesian"][i].cpu().numpy(), ref["cartesian"], atol=1e-12)
        assert np.allclose(out["polar"][i].cpu().numpy(), ref["polar"], atol=1e-6)


# ------------------------------------------------------------------------------------------------ conv / GN ops
CONV_CASES = [
    ("conv1 7x7s2 30->32", 2, 30, 32, 7, 7, 2, 3, 192, 341, 16, False, 32),
    ("layer1 3x3s1 32->32", 3, 32, 32, 3, 3, 1, 1, 48, 86, 16, True, None),
    ("layer2 3x3s1 64->64", 5, 64, 64, 3, 3, 1, 1, 24, 43, 16, True, None),
    ("raster many units 32->32", 160, 32, 32, 3, 3, 1, 1, 48, 86, 16, True, None),
    ("layer2.0 3x3s2 32->64", 3, 32, 64, 3, 3, 2, 1, 48, 86, 16, True, None),
    ("layer2.0 down 1x1s2 32->64", 3, 32, 64, 1, 1, 2, 0, 48, 86, 16, True, None),
    ("layer3.0 3x3s2 64->128 (odd width)", 3, 64, 128, 3, 3, 2, 1, 24, 43, 16, True, None),
    ("3x3s2 64->128 (odd height and width)", 2, 64, 128, 3, 3, 2, 1, 23, 43, 16, True, None),
    ("layer3.0 down 1x1s2 64->128 (odd width)", 3, 64, 128, 1, 1, 2, 0, 24, 43, 16, True, None),
    ("layer3 3x3s1 128->128", 3, 128, 128, 3, 3, 1, 1, 12, 22, 16, True, None),
    ("raster128 many units 128->128", 300, 128, 128, 3, 3, 1, 1, 12, 22, 16, True, None),
    ("layer4 3x3s1 256->256", 3, 256, 256, 3, 3, 1, 1, 6, 11, 16, True, None),
    ("compression 256->31", 3, 256, 31, 3, 3, 1, 1, 6, 11, 1, True, None),
    ("policy conv1 1->32", 2, 1, 32, 7, 7, 2, 3, 96, 170, 16, False, None),
    ("policy conv1 1->32 (odd sizes, partial tiles)", 5, 1, 32, 7, 7, 2, 3, 37, 53, 16, False, None),
    ("r50 1x1 256->1024", 2, 256, 1024, 1, 1, 1, 0, 6, 11, 16, True, None),
    ("fc 2112->512", 64, 2112, 512, 1, 1, 1, 0, 1, 1, 16, True, 2112),
    ("tiny 1 pixel tile", 1, 32, 32, 3, 3, 1, 1, 1, 1, 16, True, None),
]


@pytest.mark.parametrize("force_generic", [0, 1, 2])
@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_fprop_dgrad_wgrad(case, force_generic):
    from pointnav_vo_b200 import lib as L
    from pointnav_vo_b200.engine import ConvLayer

    name, B, Cin, Cout, R, S, stride, pad, IH, IW, G, bwd, cin_pad = case
    torch.manual_seed(0)
    dev = "cuda"
    c = ConvLayer("w", Cin, Cout, R, S, stride, pad, IH, IW, need_dgrad=bwd, cin_pad=cin_pad)
    c.alloc(dev, True)
    w = (torch.randn(Cout, Cin, R, S, device=dev) / (Cin * R * S) ** 0.5).contiguous()
    x = torch.randn(B, IH, IW, c.cin_pad, device=dev).half()
    x[..., Cin:] = 0
    stats = torch.zeros(B, G, 2, device=dev, dtype=torch.float64)
    y = torch.empty(B, c.OH, c.OW, c.cout_pad, dtype=torch.float16, device=dev)
    fwd = c.op_fwd(x, y, B, stats, c.cout_pad // G, G)
    fwd.i[19] = force_generic
    L.run_ops([c.op_pack(w), fwd])
    xr = x[..., :Cin].float().permute(0, 3, 1, 2)
    wr = w.half().float()
    ref = F.conv2d(xr, wr, None, stride, pad)
    assert rel(y[..., :Cout].permute(0, 3, 1, 2), ref) <= 3e-3
    assert bool((y[..., Cout:] == 0).all())
    if c.cout_pad == Cout:
        rs = ref.reshape(B, G, -1)
        assert rel(stats, torch.stack((rs.sum(-1), rs.pow(2).sum(-1)), -1)) <= 1e-4
    dy = torch.randn(B, c.OH, c.OW, c.cout_pad, device=dev).half()
    dy[..., Cout:] = 0
    gw = torch.zeros(Cout, Cin, R, S, device=dev)
    xr2, wr2 = xr.clone().requires_grad_(True), wr.clone().requires_grad_(True)
    F.conv2d(xr2, wr2, None, stride, pad).backward(dy[..., :Cout].float().permute(0, 3, 1, 2))
    wg = c.op_wgrad(x, dy, B)
    wg.i[19] = force_generic
    L.run_ops([L.op_zero(c.dwp), wg, c.op_unpack(gw)])
    assert rel(gw, wr2.grad) <= 3e-3
    if bwd:
        add = torch.randn(B, IH, IW, c.cin_pad, device=dev).half()
        if c.s2_classes:
            # 3x3 / stride 2: four parity-class convolutions of dy written straight onto their lattice of gx (each producer)
            gx = torch.full_like(add, float("nan"))
            ops = c.ops_dgrad(dy, gx, B, add=add)
            assert len(ops) == 4
            for o in ops:
                o.i[19] = force_generic
            L.run_ops(ops)
            assert rel(gx[..., :Cin].permute(0, 3, 1, 2), xr2.grad + add[..., :Cin].float().permute(0, 3, 1, 2)) <= 3e-3
            return
        gx = torch.empty_like(add)
        dg = c.op_dgrad(dy, gx, B, add=add)
        dg.i[19] = force_generic
        L.run_ops([dg])
        assert rel(gx[..., :Cin].permute(0, 3, 1, 2), xr2.grad + add[..., :Cin].float().permute(0, 3, 1, 2)) <= 3e-3
        if stride == 2 and force_generic == 0:
            # the engine's route: zero-upsampled dy through the stride-1 kernels (3x3), compact 1x1 + scatter (1x1)
            gx2 = torch.full_like(add, float("nan"))
            a2 = add if R == 3 else None
            L.run_ops(c.ops_dgrad(dy, gx2, B, add=a2))
            want = xr2.grad + (add[..., :Cin].float().permute(0, 3, 1, 2) if a2 is not None else 0)
            assert rel(gx2[..., :Cin].permute(0, 3, 1, 2), want) <= 3e-3


@pytest.mark.parametrize("force_generic", [0, 1, 2])
@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: c[0])
def test_conv_split_fp16_operands(case, force_generic):
    """Split-fp16 mode of the conv op: fp32 inputs / weights held as value + residual fp16 planes, three MMAs per
    product, fp32 output.  Checked against an fp64 convolution of the fp32 tensors: max|d| <= 1e-4 * rms (measured
    3e-5: what is left is the tensor core's fp32 accumulation, which truncates rather than rounds; the single-pass
    fp16 kernel gives ~1e-3 on the same data).  force_generic 0 = the kernel the engine uses (split variants of the
    resident-weight raster for 32 channels, the streamed-weight raster for 64 / 128 channels, TMA im2col otherwise),
    1 = cp.async im2col producer, 2 = TMA im2col producer."""
    from pointnav_vo_b200 import lib as L
    from pointnav_vo_b200.engine import ConvLayer

    name, B, Cin, Cout, R, S, stride, pad, IH, IW, G, bwd, cin_pad = case
    torch.manual_seed(1)
    dev = "cuda"
    c = ConvLayer("w", Cin, Cout, R, S, stride, pad, IH, IW, need_dgrad=False, cin_pad=cin_pad)
    c.alloc(dev, False, split=True)
    w = (torch.randn(Cout, Cin, R, S, device=dev) / (Cin * R * S) ** 0.5).contiguous()
    x32 = torch.randn(B, IH, IW, c.cin_pad, device=dev)
    x32[..., Cin:] = 0
    x_hi = x32.half()
    x_lo = (x32 - x_hi.float()).half()
    stats = torch.zeros(B, G, 2, device=dev, dtype=torch.float64)
    y = torch.empty(B, c.OH, c.OW, c.cout_pad, dtype=torch.float32, device=dev)
    tab = L.device_table([c.pack_desc(w), c.pack_desc_lo(w)], dev)
    fwd = c.op_fwd(x_hi, y, B, stats, c.cout_pad // G, G, True, x_lo=x_lo)
    fwd.i[19] = force_generic
    L.run_ops([L.op_multi(L.OP_PACK_W_MULTI, tab, 2), fwd])
    ref = F.conv2d(x32[..., :Cin].double().permute(0, 3, 1, 2), w.double(), None, stride, pad)
    err = rel(y[..., :Cout].permute(0, 3, 1, 2).double(), ref)
    print(name, "split conv max|d|/rms", err)
    assert err <= 1e-4
    assert bool((y[..., Cout:] == 0).all())
    if c.cout_pad == Cout:
        rs = ref.reshape(B, G, -1)
        assert rel(stats, torch.stack((rs.sum(-1), rs.pow(2).sum(-1)), -1)) <= 1e-4  # truncating accumulation: ~1e-5 low
    # the same output as value + residual fp16 planes (what the plans store: backward reads the value plane alone)
    yh = torch.full((B, c.OH, c.OW, c.cout_pad), float("nan"), dtype=torch.float16, device=dev)
    yl = torch.full_like(yh, float("nan"))
    fwd2 = c.op_fwd(x_hi, yh, B, None, 0, 0, False, x_lo=x_lo, y_lo=yl)
    fwd2.i[19] = force_generic
    L.run_ops([fwd2])
    assert torch.equal(yh, y.half())
    assert rel(yh.float() + yl.float(), y) <= 1e-6


@pytest.mark.parametrize("version", [1, 2])
@pytest.mark.parametrize("B,IH,IW,Cin", [(2, 192, 341, 30), (3, 16, 341, 30), (150, 10, 341, 8)])
def test_stem_conv_kernels(version, B, IH, IW, Cin):
    """The two stem kernels (row raster with N = 32; pixels-as-N with masked filter-row windows) against torch fp32
    on fp16-rounded operands: outputs <= 3e-3 * rms (one fp16 rounding), GroupNorm partial sums <= 1e-4."""
    from pointnav_vo_b200 import lib as L

    dev = "cuda"
    torch.manual_seed(0)
    OH, OW = (IH - 1) // 2 + 1, (IW - 1) // 2 + 1
    w = (torch.randn(32, Cin, 7, 7, device=dev) / (Cin * 49) ** 0.5).contiguous()
    Wp = L.load().pnvo_stem_padded_width(IW)
    xp = torch.zeros(B, IH, Wp, 32, device=dev, dtype=torch.float16)
    x = torch.randn(B, IH, IW, 32, device=dev).half()
    x[..., Cin:] = 0
    xp[:, :, 3:3 + IW] = x
    y = torch.full((B, OH, OW, 32), float("nan"), dtype=torch.float16, device=dev)
    stats = torch.zeros(B, 16, 2, device=dev, dtype=torch.float64)
    wr = torch.zeros(4 * 7 * 32, 64, dtype=torch.float16, device=dev)
    if version == 1:
        if not 128 <= OW <= 256:
            pytest.skip("row-raster stem kernel needs 128 <= OW <= 256")
        ops = [L.op_pack_w_stem(w, wr, Cin), L.op_conv_stem(xp, wr, y, stats, B, IH, IW, 16, 2, 2)]
    else:
        assert L.load().pnvo_conv_stem2_supported(IH, IW) == 1
        ops = [L.op_pack_w_stem2(w, wr, Cin), L.op_conv_stem2(xp, wr, y, stats, B, IH, IW, 16, 2)]
    L.run_ops(ops)
    ref = F.conv2d(x[..., :Cin].float().permute(0, 3, 1, 2), w.half().float(), None, 2, 3)
    assert rel(y.permute(0, 3, 1, 2), ref) <= 3e-3
    rs = ref.reshape(B, 16, -1)
    assert rel(stats, torch.stack((rs.sum(-1), rs.pow(2).sum(-1)), -1)) <= 1e-4
    if version == 2:
        # split-fp16 stem (the engine's default precision): fp32 input / weights as value + residual planes; launch 1
        # = w_lo * x into an fp16 tensor, launch 2 = w * (x, x_lo) + that tensor -> fp32 raw output + statistics.
        # Against an fp64 convolution of the fp32 tensors: <= 1e-4 * rms.
        x32 = torch.randn(B, IH, IW, 32, device=dev)
        x32[..., Cin:] = 0
        xh = x32.half()
        xl = (x32 - xh.float()).half()
        xp_hi, xp_lo = torch.zeros_like(xp), torch.zeros_like(xp)
        xp_hi[:, :, 3:3 + IW] = xh
        xp_lo[:, :, 3:3 + IW] = xl
        wr_lo = torch.zeros_like(wr)
        corr = torch.full((B, OH, OW, 32), float("nan"), dtype=torch.float16, device=dev)
        y32 = torch.full((B, OH, OW, 32), float("nan"), dtype=torch.float32, device=dev)
        stats.zero_()
        L.run_ops([L.op_pack_w_stem2(w, wr, Cin), L.op_pack_w_stem2(w, wr_lo, Cin, lo=True),
                   L.op_conv_stem2(xp_hi, wr_lo, corr, None, B, IH, IW, 16, 2),
                   L.op_conv_stem2(xp_hi, wr, y32, stats, B, IH, IW, 16, 2, x_lo=xp_lo, add=corr, out_fp32=True)])
        ref64 = F.conv2d(x32[..., :Cin].double().permute(0, 3, 1, 2), w.double(), None, 2, 3)
        err = rel(y32.permute(0, 3, 1, 2).double(), ref64)
        print("split stem max|d|/rms", err)
        assert err <= 1e-4
        yh, yl = torch.empty_like(y), torch.empty_like(y)
        L.run_ops([L.op_conv_stem2(xp_hi, wr, yh, None, B, IH, IW, 16, 2, x_lo=xp_lo, add=corr, y_lo=yl)])
        assert torch.equal(yh, y32.half()) and rel(yh.float() + yl.float(), y32) <= 1e-6
        rs = ref64.reshape(B, 16, -1)
        assert rel(stats, torch.stack((rs.sum(-1), rs.pow(2).sum(-1)), -1)) <= 1e-4
    # weight gradient on the same staged rows
    dy = torch.randn(B, OH, OW, 32, device=dev).half()
    w_ld = 1600
    dw = torch.zeros(32, w_ld, device=dev)
    if version == 1:
        if not 96 < OW <= 176:
            return
        L.run_ops([L.op_wgrad_stem(xp, dy, dw, B, IH, IW, w_ld, 8)])
    else:
        assert L.load().pnvo_conv_stem_wgrad2_supported(IH, IW) == 1
        L.run_ops([L.op_wgrad_stem2(xp, dy, dw, B, IH, IW, w_ld)])
    wz = torch.zeros(32, Cin, 7, 7, device=dev, requires_grad=True)
    F.conv2d(x[..., :Cin].float().permute(0, 3, 1, 2), wz, None, 2, 3).backward(dy.float().permute(0, 3, 1, 2))
    got = dw[:, :49 * 32].reshape(32, 7, 7, 32)[..., :Cin].permute(0, 3, 1, 2)
    assert rel(got, wz.grad) <= 3e-3


@pytest.mark.parametrize("B,H,W,C,G,Cr", [(3, 24, 43, 64, 16, 64), (2, 6, 11, 32, 1, 31), (2, 12, 22, 128, 16, 128),
                                         (2, 3, 6, 128, 1, 114), (3, 48, 86, 32, 16, 32), (2, 6, 11, 256, 16, 256)])
def test_groupnorm_forward_backward(B, H, W, C, G, Cr):
    from pointnav_vo_b200 import lib as L

    torch.manual_seed(0)
    dev = "cuda"
    x = torch.randn(B, H, W, C, device=dev).half()
    x[..., Cr:] = 0
    res = torch.randn(B, H, W, C, device=dev).half()
    res[..., Cr:] = 0
    gamma, beta = torch.rand(Cr, device=dev) + 0.5, torch.randn(Cr, device=dev) * 0.1
    xf = x[..., :Cr].float().permute(0, 3, 1, 2)
    xs = xf.reshape(B, G, -1)
    stats = torch.stack((xs.sum(-1), xs.pow(2).sum(-1)), -1).double().contiguous()
    y = torch.empty_like(x)
    cpg, cpg_r, HW = C // G, Cr // G, H * W
    L.run_ops([L.op_gn_apply(x, stats, gamma, beta, y, B, C, G, cpg, HW, float(cpg_r * HW), True, res, False, 1e-5, Cr)])
    xr, gr, br = xf.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.relu(F.group_norm(xr, G, gr, br, 1e-5) + res[..., :Cr].float().permute(0, 3, 1, 2))
    assert rel(y[..., :Cr].permute(0, 3, 1, 2), ref) <= 3e-3
    assert bool((y[..., Cr:] == 0).all()) or Cr == C
    g = torch.randn(B, H, W, C, device=dev).half()
    g[..., Cr:] = 0
    ref.backward(g[..., :Cr].float().permute(0, 3, 1, 2))
    sums = torch.zeros(B, C, 2, device=dev)
    dx, dyo = torch.empty_like(x), torch.empty_like(x)
    dg, db = torch.zeros(Cr, device=dev), torch.zeros(Cr, device=dev)
    yfull = torch.zeros_like(x)
    yfull[..., :Cr] = ref.detach().permute(0, 2, 3, 1).half()
    args = (g, yfull, x, stats, gamma, sums, dx, dyo, B, C, G, cpg, HW, float(cpg_r * HW), False, 1e-5, Cr)
    L.run_ops([L.op_gn_bwd(True, *args), L.op_gn_param_grad(sums, dg, db, B, C, Cr), L.op_gn_bwd(False, *args)])
    assert rel(dx[..., :Cr].permute(0, 3, 1, 2), xr.grad) <= 4e-3
    assert rel(dg, gr.grad) <= 1e-3 and rel(db, br.grad) <= 1e-3
    mask = (ref.detach() > 0).float()
    assert rel(dyo[..., :Cr].permute(0, 3, 1, 2), g[..., :Cr].float().permute(0, 3, 1, 2) * mask) <= 1e-6
    # one-pass variant (cluster per sample, DSMEM reduction): same results, bit-identical from run to run
    assert L.load().pnvo_gn_bwd_fused_supported(C, HW, 0) == 1
    outs = []
    for _ in range(2):
        sums2 = torch.full((B, C, 2), float("nan"), device=dev)  # the fused kernel needs no pre-zeroed buffer
        dx2, dyo2 = torch.empty_like(x), torch.empty_like(x)
        dg2, db2 = torch.zeros(Cr, device=dev), torch.zeros(Cr, device=dev)
        args2 = (g, yfull, x, stats, gamma, sums2, dx2, dyo2, B, C, G, cpg, HW, float(cpg_r * HW), False, 1e-5, Cr)
        L.run_ops([L.op_gn_bwd("fused", *args2), L.op_gn_param_grad(sums2, dg2, db2, B, C, Cr)])
        outs.append((dx2, dyo2, dg2, db2))
    assert rel(dx2[..., :Cr].permute(0, 3, 1, 2), xr.grad) <= 4e-3
    assert rel(dg2, gr.grad) <= 1e-3 and rel(db2, br.grad) <= 1e-3
    assert torch.equal(dyo2, dyo)
    for u, v in zip(*outs):
        assert torch.equal(u, v)
    # gradient scale (the compression GroupNorm behind inverted dropout), no identity-branch output
    args3 = (g, yfull, x, stats, gamma, sums, dx, None, B, C, G, cpg, HW, float(cpg_r * HW), False, 1e-5, Cr, 1.25)
    sums.zero_()
    L.run_ops([L.op_gn_bwd(True, *args3), L.op_gn_bwd(False, *args3)])
    dx3 = torch.empty_like(x)
    args4 = (g, yfull, x, stats, gamma, sums2, dx3, None, B, C, G, cpg, HW, float(cpg_r * HW), False, 1e-5, Cr, 1.25)
    L.run_ops([L.op_gn_bwd("fused", *args4)])
    assert rel(dx3, dx) <= 4e-3  # two fp16-rounded evaluations of the same expression: one ulp of the largest values
    # no ReLU (downsample branch GroupNorm)
    args5 = (g, None, x, stats, gamma, sums, dx, None, B, C, G, cpg, HW, float(cpg_r * HW), False, 1e-5, Cr)
    sums.zero_()
    L.run_ops([L.op_gn_bwd(True, *args5), L.op_gn_bwd(False, *args5)])
    args6 = (g, None, x, stats, gamma, sums2, dx3, None, B, C, G, cpg, HW, float(cpg_r * HW), False, 1e-5, Cr)
    L.run_ops([L.op_gn_bwd("fused", *args6)])
    assert rel(dx3, dx) <= 4e-3
    # fp32 raw conv outputs (the split-precision plans): both variants read x as fp32
    x32 = x.float().contiguous()
    assert L.load().pnvo_gn_bwd_fused_supported(C, HW, 1) == 1
    dx7, dyo7, dx8, dyo8 = (torch.empty_like(x) for _ in range(4))
    sums7 = torch.zeros(B, C, 2, device=dev)
    args7 = (g, yfull, x32, stats, gamma, sums7, dx7, dyo7, B, C, G, cpg, HW, float(cpg_r * HW), True, 1e-5, Cr)
    args8 = (g, yfull, x32, stats, gamma, sums2, dx8, dyo8, B, C, G, cpg, HW, float(cpg_r * HW), True, 1e-5, Cr)
    L.run_ops([L.op_gn_bwd(True, *args7), L.op_gn_bwd(False, *args7), L.op_gn_bwd("fused", *args8)])
    assert rel(dx7[..., :Cr].permute(0, 3, 1, 2), xr.grad) <= 4e-3
    assert rel(dx8[..., :Cr].permute(0, 3, 1, 2), xr.grad) <= 4e-3
    assert torch.equal(dyo8, dyo) and torch.equal(dyo7, dyo)
    assert rel(sums2, sums7) <= 1e-4


def test_groupnorm_maxpool_forward_backward():
    from pointnav_vo_b200 import lib as L

    torch.manual_seed(1)
    dev = "cuda"
    B, H, W, C, G, PH, PW = 2, 96, 171, 32, 16, 48, 86
    x = torch.randn(B, H, W, C, device=dev).half()
    gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
    xf = x.float().permute(0, 3, 1, 2)
    xs = xf.reshape(B, G, -1)
    stats = torch.stack((xs.sum(-1), xs.pow(2).sum(-1)), -1).double().contiguous()
    y = torch.empty(B, PH, PW, C, dtype=torch.float16, device=dev)
    am = torch.empty(B, PH, PW, C, dtype=torch.uint8, device=dev)
    L.run_ops([L.op_gn_pool(x, stats, gamma, beta, y, am, B, C, G, C // G, H, W, PH, PW, float(C // G * H * W))])
    a = F.relu(F.group_norm(xf, G, gamma, beta, 1e-5)).requires_grad_(True)
    ref = F.max_pool2d(a, 3, 2, 1)
    assert rel(y.permute(0, 3, 1, 2), ref) <= 3e-3
    gp = torch.randn(B, PH, PW, C, device=dev).half()
    ref.backward(gp.float().permute(0, 3, 1, 2))
    dy = torch.empty(B, H, W, C, dtype=torch.float16, device=dev)
    L.run_ops([L.op_pool_bwd(gp, y, am, dy, B, C, H, W, PH, PW)])
    want = a.grad * (a.detach() > 0).float()
    assert rel_l2(dy.permute(0, 3, 1, 2), want) <= 2e-3  # fp16 rounding of sums of up to 4 gradients


# ------------------------------------------------------------------------------------------------ whole networks
def _load_vo(case):
    from pointnav_vo_b200.vo.models import vo_cnn, vo_cnn_act_embed  # noqa: F401  (registers the act-embed variants)

    name, space, backbone, kw = helpers.VO_CASES[case]
    cls = vo_cnn.VisualOdometryCNNBase if name == "base" else vo_cnn.baseline_registry.get_vo_model(name)
    m = cls(observation_space=space, observation_size=(341, 192), hidden_size=512, backbone=backbone,
            normalize_visual_inputs=True, output_dim=3, dropout_p=0.0, **kw)
    m.load_state_dict(helpers.vo_state_dict(case))
    return m.cuda(), space, backbone


# Default precision ("split": value + residual fp16 operand planes in every forward convolution, single-pass fp16
# backward) -- the mode bench.py times.  Forward: BASELINE.json north_star "within 1e-3 rel fp32" against the reference's
# own fp32 outputs.  Gradients: relative L2 per parameter tensor against the reference's fp32 gradients.
SPLIT_TOL = 1e-3
GRAD_TOL_SPLIT = 2e-2
# "fp16" throughput mode (single-pass operands, fp16 stored activations): measured 3e-3 .. 6e-3 run to run
FWD_TOL = {"r18_30ch": 8e-3, "r18_8ch": 8e-3, "r50_8ch": 2.5e-2, "r18_8ch_act_embed": 8e-3}
GRAD_TOL = {"r18_30ch": 0.15, "r18_8ch": 0.15, "r50_8ch": 0.35, "r18_8ch_act_embed": 0.15}  # relative L2 per tensor (ReLU-flip noise, 53 layers)


@pytest.mark.parametrize("case,precision", [(c, p) for c in ("r18_30ch", "r18_8ch", "r50_8ch", "r18_8ch_act_embed")
                                            for p in ("split", "fp16")] +
                         [("r18_wider", "split"), ("r101_deeper", "split")])  # vo_cnn_wider / vo_cnn_deeper (vo_cnn.py:308-375)
def test_vo_model_against_reference_golden(case, precision, golden_dir):
    g = np.load(os.path.join(golden_dir, f"vo_{case}.npz"))
    m, space, backbone = _load_vo(case)
    assert m.precision == "split"  # the default
    m.set_precision(precision)
    fwd_tol = SPLIT_TOL if precision == "split" else FWD_TOL[case]
    grad_tol = GRAD_TOL_SPLIT if precision == "split" else GRAD_TOL[case]
    obs = helpers.vo_inputs(2, 11, space, "cuda")
    if "actions" in g.files:  # act-embed variants: forward(observation_pairs, actions) (vo_cnn_act_embed.py:65)
        acts = torch.from_numpy(g["actions"]).cuda()
        net = m
        m_call = lambda o: net(o, acts)  # noqa: E731
    else:
        m_call = m
    m.eval()
    with torch.no_grad():
        y = m_call(obs)
    assert y.shape == (2, 3)
    print(case, precision, "eval max|d|/rms", rel(y, torch.from_numpy(g["eval_out"])))
    assert rel(y, torch.from_numpy(g["eval_out"])) <= fwd_tol
    # training-mode forward: running statistics are updated exactly like the reference's buffers
    m.train()
    target = torch.from_numpy(g["target"]).cuda()
    y = m_call(obs)
    loss = sum(vo.vo_losses(y, target))
    loss.backward()
    print(case, precision, "train max|d|/rms", rel(y, torch.from_numpy(g["train_out"])))
    assert rel(y, torch.from_numpy(g["train_out"])) <= fwd_tol
    assert abs(loss.item() - float(g["train_loss"])) <= (1e-3 if precision == "split" else 2e-2) * float(g["train_loss"])
    sd = m.state_dict()
    assert rel(sd["visual_encoder.running_mean_and_var._mean"], torch.from_numpy(g["train_mean"])) <= 1e-5
    assert rel(sd["visual_encoder.running_mean_and_var._var"], torch.from_numpy(g["train_var"])) <= 1e-5
    assert float(sd["visual_encoder.running_mean_and_var._count"]) == float(g["train_count"])
    P = dict(m.named_parameters())
    norms = dict(zip([str(k) for k in g["grad_keys"]], g["grad_norms"]))
    assert set(norms) == set(P)
    for k, n in norms.items():
        assert P[k].grad is not None and torch.isfinite(P[k].grad).all(), k
        assert abs(P[k].grad.norm().item() - n) <= grad_tol * n + 1e-7, (k, P[k].grad.norm().item(), n)
    worst = 0.0
    for k in g.files:
        if k.startswith("grad/") and g[k].size > 64:
            e = rel_l2(P[k[5:]].grad, torch.from_numpy(g[k]))
            worst = max(worst, e)
            assert e <= grad_tol, (k, e)
    print(case, precision, "worst gradient rel-L2", worst)


@pytest.mark.parametrize("case", ["r18_30ch", "r18_8ch", "r50_8ch", "r18_8ch_act_embed"])
def test_vo_split_precision_forward_within_north_star_tolerance(case, golden_dir):
    """set_precision("split"): value + residual fp16 planes, three tensor-core products per convolution.  The eval-mode
    outputs must agree with the reference's own fp32 outputs (golden, generated by the unmodified reference) within the
    north-star 1e-3 -- both max|d| / rms(ref) and element by element relative to |ref|."""
    g = np.load(os.path.join(golden_dir, f"vo_{case}.npz"))
    m, space, backbone = _load_vo(case)
    m.set_precision("split").eval()
    obs = helpers.vo_inputs(2, 11, space, "cuda")
    want = torch.from_numpy(g["eval_out"])
    with torch.no_grad():
        if "actions" in g.files:
            y = m(obs, torch.from_numpy(g["actions"]).cuda())
        else:
            y = m(obs)
    d = (y.float().cpu() - want).abs()
    print(case, "split: max|d|/rms", rel(y, want), "max elementwise rel", (d / want.abs().clamp_min(1e-6)).max().item())
    assert rel(y, want) <= SPLIT_TOL
    assert (d <= SPLIT_TOL * want.abs() + 1e-6).all(), (y, want)
    # the fast mode on the same module still works (separate plan) and is the less accurate of the two
    m.set_precision("fp16")
    with torch.no_grad():
        y16 = m(obs, torch.from_numpy(g["actions"]).cuda()) if "actions" in g.files else m(obs)
    assert rel(y16, want) <= FWD_TOL[case]
    assert rel(y, want) < rel(y16, want)


def test_vo_split_precision_raw_pairs():
    """The raw uint8 / fp32-depth input path in split mode equals the dict path in split mode (same derived channels),
    and a no-grad training-mode forward (the RL loop's "rnd" mode, dropout active) runs."""
    m, space, _ = _load_vo("r18_30ch")
    m.set_precision("split").eval()
    obs = helpers.vo_inputs(2, 21, space, "cuda")
    raw = {"rgb": obs["rgb"].to(torch.uint8).contiguous(), "depth": obs["depth"].contiguous()}
    with torch.no_grad():
        y_raw = m(raw)
        y_dict = m(obs)
    assert rel(y_raw, y_dict) <= 1e-4  # rgb / 255 as a division vs a multiplication: one fp32 ulp on the input
    m.train()
    with torch.no_grad():
        y = m(raw)
    assert torch.isfinite(y).all()


@pytest.mark.parametrize("case", ["r18_30ch", "r18_8ch"])
def test_raw_input_pipeline_matches_reference_inputs(case):
    """uint8 rgb + fp32 depth pairs (dd / top-down derived on the device) against the same model fed with the
    reference's four fp32 tensors (dd / top-down from the pinned numpy oracle): the assembled fp16 input must
    be identical up to the rgb/255 division-vs-multiplication ulp, the running statistics equal to 1e-6."""
    model, space, backbone = _load_vo(case)
    import copy

    B = 3
    obs = helpers.vo_inputs(B, 21, space, "cuda")
    raw = {"rgb": obs["rgb"].to(torch.uint8).contiguous(), "depth": obs["depth"].contiguous()}
    assert torch.equal(raw["rgb"].float(), obs["rgb"])
    m2 = copy.deepcopy(model)
    for train in (False, True):
        model.train(train)
        m2.train(train)
        with torch.no_grad():
            y1 = model(obs)
            x1 = model._plan_for(obs, False, train).x0.clone()
            y2 = m2(raw)
            x2 = m2._plan_for(raw, False, train).x0.clone()
        d = (x1.float() - x2.float()).abs()
        assert d.max().item() <= 2e-3 and (d > 0).float().mean().item() < 0.05  # isolated fp16 ulps only
        if train:
            r1, r2 = model.visual_encoder.running_mean_and_var, m2.visual_encoder.running_mean_and_var
            assert rel(r2._mean, r1._mean) <= 1e-5 and rel(r2._var, r1._var) <= 1e-5
            assert float(r1._count) == float(r2._count)
        else:
            # same noise floor as two runs of one path: fp32-atomic GroupNorm statistics flip isolated fp16 roundings
            assert rel(y2, y1) <= 6e-3


@pytest.mark.parametrize("case", ["r18_30ch", "r18_8ch"])
def test_inverse_pair_augmentation_on_device(case):
    """8f-2: `pair_map` rows (2 * source pair + swap flag) expanded by the input kernels == the reference dataset's way
    (a second, channel-swapped copy of the frames built on the host, regression_geo_invariance_iter_dataset.py:342-366):
    identical assembled input in eval mode, equal running statistics and outputs in training mode."""
    import copy

    from pointnav_vo_b200.vo.dataset import geo_invariance as gi

    model, space, _ = _load_vo(case)
    obs = helpers.vo_inputs(3, 23, space, "cuda")
    rgb, dep = obs["rgb"].to(torch.uint8).contiguous(), obs["depth"].contiguous()
    pm = gi.make_pair_map([2, 3, 2], act_type=[2, 3], geo_invariance_types=("inverse_joint_train",))
    pm["pair_map"] = np.concatenate([pm["pair_map"], np.array([4, 1], np.int32)])  # ragged tail: repeats, odd row count
    pmap = torch.from_numpy(pm["pair_map"]).cuda()
    src, sw = (pmap >> 1).long(), (pmap & 1).bool()

    def expand(t, half):
        e = t[src].clone()
        e[sw] = torch.cat([e[sw][..., half:], e[sw][..., :half]], -1)
        return e.contiguous()

    host = {"rgb": expand(rgb, 3), "depth": expand(dep, 1)}
    devm = {"rgb": rgb, "depth": dep, "pair_map": pmap}
    m2 = copy.deepcopy(model)
    R = pmap.numel()
    for train in (False, True):
        model.train(train)
        m2.train(train)
        with torch.no_grad():
            y1 = model(host)
            x1 = model._plan_for(host, False, train).x0.clone()
            y2 = m2(devm)
            x2 = m2._plan_for(devm, False, train).x0.clone()
        assert y1.shape == y2.shape == (R, 3)
        if not train:
            assert torch.equal(x1, x2)
            assert torch.equal(y1, y2)
        else:
            r1, r2 = model.visual_encoder.running_mean_and_var, m2.visual_encoder.running_mean_and_var
            assert rel(r2._mean, r1._mean) <= 1e-6 and rel(r2._var, r1._var) <= 1e-6
            assert float(r1._count) == float(r2._count)
            d = (x1.float() - x2.float()).abs()
            assert d.max().item() <= 2e-3 and (d > 0).float().mean().item() < 0.01  # isolated fp16 ulps from the fp32 statistics
            assert rel(y2, y1) <= 8e-3
    # one fused optimisation step with the inversion loss on an interleaved batch: same loss either way
    from pointnav_vo_b200.vo.engine.train_step import FusedVOTrainStep

    pm = gi.make_pair_map([2, 3, 2], act_type=[2, 3], geo_invariance_types=("inverse_joint_train",))
    pmap = torch.from_numpy(pm["pair_map"]).cuda()
    src, sw = (pmap >> 1).long(), (pmap & 1).bool()
    host = {"rgb": expand(rgb, 3), "depth": expand(dep, 1)}
    devm = {"rgb": rgb, "depth": dep, "pair_map": pmap}
    rng = np.random.default_rng(5)
    tg = torch.from_numpy(gi.expand_targets(rng.normal(0, 0.2, (3, 3)).astype(np.float32), pm)).cuda()
    acts = torch.from_numpy(pm["actions"]).cuda()
    losses = []
    for m, o in ((model, host), (m2, devm)):
        m.train()
        t = FusedVOTrainStep(m, loss_inv_weight=1.0, move_forward_id=1)
        losses.append([float(t.step(o, tg, acts).item()) for _ in range(2)])
    assert np.allclose(losses[0], losses[1], rtol=2e-2), losses


def test_raw_pipeline_fp16_depth_is_exact():
    """8f-2: depth pairs shipped in the dataset's storage type (float16, regression_geo_invariance_iter_dataset.py:229-236)
    and widened on the device == the reference's route (widen on the host, ship fp32): identical top-down maps, identical
    assembled input and outputs, equal running statistics."""
    import copy

    from pointnav_vo_b200.utils import geometry_utils as gu

    model, space, _ = _load_vo("r18_30ch")
    obs = helpers.vo_inputs(3, 27, space, "cuda")
    rgb = obs["rgb"].to(torch.uint8).contiguous()
    d16 = obs["depth"].half().contiguous()
    d32 = d16.float().contiguous()
    gen = gu.NormalizedDepth2TopDownViewHabitatTorch(0.1, 10.0, 192, 341, 70)
    assert torch.equal(gu.gen_top_down_view_pairs(gen, d16), gu.gen_top_down_view_pairs(gen, d32))
    m2 = copy.deepcopy(model)
    m3 = copy.deepcopy(model)
    model.exact_stem = m2.exact_stem = False  # same stem formulation on both routes: the inputs must then be identical
    for train in (False, True):
        model.train(train)
        m2.train(train)
        with torch.no_grad():
            y1 = model({"rgb": rgb, "depth": d32})
            x1 = model._plan_for({"rgb": rgb, "depth": d32}, False, train).x0.clone()
            y2 = m2({"rgb": rgb, "depth": d16})
            x2 = m2._plan_for({"rgb": rgb, "depth": d16}, False, train).x0.clone()
        if not train:
            assert torch.equal(x1, x2) and torch.equal(y1, y2)
        else:
            r1, r2 = model.visual_encoder.running_mean_and_var, m2.visual_encoder.running_mean_and_var
            assert rel(r2._mean, r1._mean) <= 1e-6 and rel(r2._var, r1._var) <= 1e-6
            assert (x1.float() - x2.float()).abs().max().item() <= 2e-3 and rel(y2, y1) <= 1e-4
    # the exact-input stem (the default for uint8 rgb + fp16 depth) against the residual-plane route: same function
    m3.exact_stem = True
    m3.load_state_dict(model.state_dict())  # (the training-mode forward above moved the running statistics)
    model.eval()
    m3.eval()
    with torch.no_grad():
        y1 = model({"rgb": rgb, "depth": d16})
        y3 = m3({"rgb": rgb, "depth": d16})
    assert [p.exact_stem for p in m3._plans.values()] == [True]
    assert rel(y3, y1) <= 1e-4


@pytest.mark.parametrize("case", ["r18_30ch", "r18_8ch"])
def test_exact_input_stem_against_reference_golden(case, golden_dir):
    """The exact-input stem (csrc/stem_exact.cu; the path bench.py times): uint8 rgb + fp16 depth pairs hold the EXACT raw
    values in the fp16 input tensor, the normalisation is folded into the stem weights (value + residual planes) and a
    border-class bias, and the conv1 weight gradient is rebuilt from the exact-tensor gradient and border-class sums of
    dy.  Same bounds against the reference's own outputs / gradients as the default path: forward <= 1e-3, running
    statistics <= 1e-5, gradients <= 2e-2 relative L2 per tensor."""
    g = np.load(os.path.join(golden_dir, f"vo_{case}.npz"))
    m, space, backbone = _load_vo(case)
    obs = helpers.vo_inputs(2, 11, space, "cuda")
    raw = {"rgb": obs["rgb"].to(torch.uint8).contiguous(), "depth": obs["depth"].half().contiguous()}
    assert torch.equal(raw["depth"].float(), obs["depth"])  # the synthetic depth is fp16-representable (as the datasets')
    m.eval()
    with torch.no_grad():
        y = m(raw)
    assert [p.exact_stem for p in m._plans.values()] == [True]
    print(case, "exact stem eval max|d|/rms", rel(y, torch.from_numpy(g["eval_out"])))
    assert rel(y, torch.from_numpy(g["eval_out"])) <= SPLIT_TOL
    m.train()
    target = torch.from_numpy(g["target"]).cuda()
    y = m(raw)
    loss = sum(vo.vo_losses(y, target))
    loss.backward()
    print(case, "exact stem train max|d|/rms", rel(y, torch.from_numpy(g["train_out"])))
    assert rel(y, torch.from_numpy(g["train_out"])) <= SPLIT_TOL
    sd = m.state_dict()
    assert rel(sd["visual_encoder.running_mean_and_var._mean"], torch.from_numpy(g["train_mean"])) <= 1e-5
    assert rel(sd["visual_encoder.running_mean_and_var._var"], torch.from_numpy(g["train_var"])) <= 1e-5
    P = dict(m.named_parameters())
    worst = 0.0
    for k in g.files:
        if k.startswith("grad/") and g[k].size > 64:
            e = rel_l2(P[k[5:]].grad, torch.from_numpy(g[k]))
            worst = max(worst, e)
            assert e <= GRAD_TOL_SPLIT, (k, e)
    k1 = "visual_encoder.backbone.conv1.0.weight"
    print(case, "exact stem worst gradient rel-L2", worst, "conv1:", rel_l2(P[k1].grad, torch.from_numpy(g["grad/" + k1])))
    norms = dict(zip([str(k) for k in g["grad_keys"]], g["grad_norms"]))
    for k, n in norms.items():
        assert abs(P[k].grad.norm().item() - n) <= GRAD_TOL_SPLIT * n + 1e-7, (k, P[k].grad.norm().item(), n)


def test_geo_inversion_loss_and_gradient():
    """a6: geometric-inversion loss on the device against the oracle (pinned to the reference), value + gradient;
    the known-answer case (ground-truth inverse poses) must give ~0 (the reference's train_debug check)."""
    from pointnav_vo_b200 import lib as L

    torch.manual_seed(3)
    B = 64
    pred = (torch.randn(B, 3) * 0.3).cuda()
    acts = torch.randint(1, 4, (B,)).cuda()
    acts[1::2] = acts[0::2]
    p = pred.clone().requires_grad_(True)
    ref = vo.geo_invariance_inverse_loss(p, acts)
    ref.backward()
    dout = torch.full((B, 3), 0.25, device="cuda")
    loss3 = torch.tensor([1.5, 0.0, 0.0], device="cuda")
    L.run_ops([L.op_geo_inv_loss(pred, acts, dout, loss3, B, 3, 2.0, 0.5)])
    assert abs(loss3[0].item() - (1.5 + 2.0 * ref.item())) <= 1e-5 * (1 + abs(ref.item()))
    assert abs((loss3[1] + loss3[2]).item() - ref.item()) <= 1e-5 * (1 + abs(ref.item()))
    assert rel(dout - 0.25, p.grad * 2.0 * 0.5) <= 1e-5
    # known answer: b = inverse of a  =>  loss = 0
    a = torch.randn(B // 2, 3) * 0.2
    c, s_ = torch.cos(-a[:, 2]), torch.sin(-a[:, 2])
    b = torch.stack((-(c * a[:, 0] + s_ * a[:, 1]), -(-s_ * a[:, 0] + c * a[:, 1]), -a[:, 2]), 1)
    both = torch.stack((a, b), 1).reshape(B, 3).cuda()
    loss3.zero_()
    L.run_ops([L.op_geo_inv_loss(both, acts, None, loss3, B, 3, 1.0, 1.0)])
    assert loss3[0].item() < 1e-10


def test_vo_backward_block_by_block():
    """Backward logic without accumulated fp16 noise: each residual block of the oracle is fed the plan's own
    activations and upstream gradient; input gradients must agree to 2 % (mean abs / rms; measured 0.5 - 1.3 % from run
    to run: ReLU masks flip where a pre-activation is within fp16 rounding of zero)."""
    case = "r18_8ch"
    m, space, backbone = _load_vo(case)
    obs = helpers.vo_inputs(2, 11, space, "cuda")
    m.train()
    y = m(obs)
    y.square().sum().backward()
    plan = [p for p in m._plans.values() if p.training][0]
    P = dict(m.named_parameters())
    ng = m.visual_encoder.ngroups

    def nchw(t):
        return t.float().permute(0, 3, 1, 2)

    for bi in range(len(plan.blocks) - 1, -1, -1):
        blk = plan.blocks[bi]
        sd = {k: v.detach().clone().requires_grad_(True) for k, v in P.items() if blk["name"] in k}
        xin = nchw(blk["x_in"]).clone().requires_grad_(True)
        yb = vo._basic_block(xin, sd, blk["name"], ng, blk["convs"][0].stride, blk["down"] is not None)
        assert rel(nchw(blk["y"]), yb.detach()) <= 8e-3
        yb.backward(nchw(blk["g_y"]))
        gx = nchw(plan.blocks[bi - 1]["g_y"] if bi > 0 else plan.g_pool)
        err = (gx.cpu() - xin.grad.cpu()).abs().mean() / xin.grad.pow(2).mean().sqrt().cpu()
        assert err.item() <= 2e-2, (blk["name"], err.item())
        for k in sd:
            assert rel_l2(P[k].grad, sd[k].grad) <= 0.08, k


@pytest.mark.parametrize("case", ["r18_30ch", "r50_8ch"])
def test_vo_layer_taps_against_oracle(case):
    """Per-layer taps (promoted from the bring-up harness tests/selftest_gpu.py): every stage of the default-precision
    plan against the fp32 oracle run on the same inputs -- forward activations (value + residual planes) within the
    north-star 1e-3 of each tap's rms, and the gradient buffers of a training step within 3e-2 relative L2.  Separates
    "which layer" from "how much" when a whole-network bound fails."""
    m, space, backbone = _load_vo(case)
    obs = helpers.vo_inputs(2, 11, space, "cuda")
    sd = {k: v.cuda() for k, v in helpers.vo_state_dict(case).items()}
    m.eval()
    with torch.no_grad():
        m(obs)
        taps = {}
        vo.vo_forward(obs, sd, space, backbone, training=False, taps=taps)
    plan = [p for p in m._plans.values() if not p.training][0]
    full = lambda t: t.float() + plan.lo(t).float()  # noqa: E731
    C = m.visual_encoder.input_channels
    fwd = {"input": full(plan.x0)[:, :, 3:3 + plan.W, :C] if plan.x0_pitch else full(plan.x0)[..., :C],
           "conv1_raw": full(plan.raw1), "pool": full(plan.pool)}
    li = 0
    for bi, blk in enumerate(plan.blocks):
        nxt = plan.blocks[bi + 1]["name"] if bi + 1 < len(plan.blocks) else None
        if nxt is None or nxt.split(".")[-2] != blk["name"].split(".")[-2]:
            li += 1
            fwd[f"layer{li}"] = full(blk["y"])
    cc = m.visual_encoder.output_shape[0]
    fwd["compression"] = full(plan.feat)[..., :cc]
    for k, t in fwd.items():
        e = rel(t.permute(0, 3, 1, 2), taps[k])
        assert e <= SPLIT_TOL, (case, k, e)
    # training step: gradient taps against fp32 autograd through the oracle
    m.train()
    target = torch.from_numpy(np.random.default_rng(5).normal(0, 0.1, size=(2, 3)).astype(np.float32)).cuda()
    y = m(obs)
    sum(vo.vo_losses(y, target)).backward()
    sdg = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
    taps = {}
    yo, _ = vo.vo_forward(obs, sdg, space, backbone, training=True, taps=taps)
    for t in taps.values():
        if t.requires_grad:
            t.retain_grad()
    sum(vo.vo_losses(yo, target)).backward()
    plan = [p for p in m._plans.values() if p.training][0]
    bwd = {"compression": plan.g_feat[..., :cc], "pool": plan.g_pool, "conv1_raw": plan.dx1}
    li = 0
    for bi, blk in enumerate(plan.blocks):
        nxt = plan.blocks[bi + 1]["name"] if bi + 1 < len(plan.blocks) else None
        if nxt is None or nxt.split(".")[-2] != blk["name"].split(".")[-2]:
            li += 1
            bwd[f"layer{li}"] = blk["g_y"]
    worst = 0.0
    for k, t in bwd.items():
        # relative L2 per tap: a ReLU / max-pool routing decision that flips at a near-zero pre-activation moves one
        # gradient element by O(1) of the rms, which a max-norm would report as a layer-wide failure
        e = rel_l2(t.permute(0, 3, 1, 2), taps[k].grad)
        worst = max(worst, e)
        print(case, "gradient tap", k, "rel-L2", e)
        assert e <= 3e-2, (case, "grad tap", k, e)
    print(case, "worst gradient tap rel-L2", worst)


def test_vo_forward_is_reproducible_run_to_run():
    """GroupNorm statistics are accumulated with fp64 atomics, so their order no longer reaches the fp32 mean / rstd:
    two eval-mode forwards of the same batch are bit-identical (with fp32 atomics they differed at the 4e-3 level)."""
    m, space, _ = _load_vo("r18_30ch")
    obs = helpers.vo_inputs(4, 13, space, "cuda")
    m.eval()
    with torch.no_grad():
        outs = [m(obs).clone() for _ in range(4)]
    for o in outs[1:]:
        assert torch.equal(o, outs[0])


def test_vo_eval_is_batch_independent_at_full_size():
    """Size-independent property at the benchmark batch: in eval mode sample i of a batch-256 forward equals the
    same sample run in a batch of 8.  GroupNorm is per-sample, so the only coupling is the ORDER of the fp32
    atomics that accumulate its statistics; a last-bit change there flips individual fp16 roundings downstream,
    so two runs agree to the fp16-storage noise level (same bound as against the reference), not bitwise."""
    from bench import DevicePreproc, build_model, synth_batch

    dev = torch.device("cuda", 0)
    model = build_model(dev).eval()
    rgb, dep, _ = synth_batch(256, 5)
    pre = DevicePreproc(256, dev)
    obs = pre(torch.from_numpy(rgb).to(dev), torch.from_numpy(dep).to(dev))
    # preprocessing invariants at full size (base_trainer_with_vo.py:162-163, geometry_utils.py:541-554)
    assert float(obs["discretized_depth"].sum()) == 256 * 192 * 341 * 2
    td = obs["top_down_view"]
    assert float(td.max()) == 1.0 and float(td.min()) == 0.0
    with torch.no_grad():
        y_full = model(obs)
        sub = {k: v[40:48].contiguous() for k, v in obs.items()}
        y_sub = model(sub)
    assert torch.isfinite(y_full).all()
    assert rel(y_full[40:48], y_sub) <= 6e-3


def test_fused_train_step_matches_autograd_plus_adam():
    from pointnav_vo_b200.vo.engine.train_step import FusedVOTrainStep

    case = "r18_8ch"
    m1, space, _ = _load_vo(case)
    m2, _, _ = _load_vo(case)
    obs = helpers.vo_inputs(2, 11, space, "cuda")
    target = torch.randn(2, 3, device="cuda") * 0.1
    m1.train()
    m2.train()
    opt = torch.optim.Adam(m1.parameters(), lr=2.5e-4, eps=1e-8)
    y = m1(obs)
    loss1 = sum(vo.vo_losses(y, target))
    loss1.backward()
    opt.step()
    step = FusedVOTrainStep(m2, lr=2.5e-4, eps=1e-8)
    loss2 = step.step(obs, target)
    assert abs(loss1.item() - loss2.item()) <= 1e-2 * abs(loss1.item())  # two runs differ at the fp16-noise level
    P1 = dict(m1.named_parameters())
    for k, g2 in step._plan.grads.items():
        assert rel_l2(g2, P1[k].grad) <= 0.15, k  # same kernels; atomics order -> fp16 rounding / ReLU flips
    for (k, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
        # Adam's first step moves every weight by ~lr * sign(g); a gradient within rounding of zero may flip sign
        d = (a - b).abs()
        assert d.max().item() <= 2 * 2.5e-4 + 1e-7, k
        if d.numel() >= 256:
            assert d.mean().item() <= 0.2 * 2.5e-4, k
        else:  # small tensors (GroupNorm affine, biases): bound the FRACTION of sign flips instead of their mean
            assert (d > 2.5e-4).float().mean().item() <= 0.25, k


def test_policy_against_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "policy_r18_depth.npz"))
    pol = helpers.policy_state_dict(device="cuda").eval()
    dep = torch.from_numpy(synth.depth_frames(3, seed=21)[..., None]).cuda()
    obs = {"depth": dep, "pointgoal_with_gps_compass": torch.from_numpy(g["goal"]).cuda()}
    hid, prev_a, masks = (torch.from_numpy(g[k]).cuda() for k in ("hidden", "prev_actions", "masks"))
    with torch.no_grad():
        value, action, logp, new_hid = pol.act(obs, hid, prev_a, masks, deterministic=True)
        plan = list(pol.net._plans.values())[0]
        enc = (plan.feat.float() + plan.lo(plan.feat).float())[..., :114].permute(0, 3, 1, 2)
    # default precision = split (value + residual planes): the north-star bound
    assert rel(enc, torch.from_numpy(g["encoder_out"])) <= SPLIT_TOL
    assert rel(value, torch.from_numpy(g["value"])) <= SPLIT_TOL
    assert rel(new_hid, torch.from_numpy(g["new_hidden"])) <= SPLIT_TOL
    assert np.array_equal(action.cpu().numpy(), g["action"])
    assert np.allclose(logp.cpu().numpy(), g["logp"], atol=1e-4)
    pol.net.set_precision("fp16")  # the single-pass throughput mode: 21 fp16-rounded layers
    with torch.no_grad():
        value16, action16, _, hid16 = pol.act(obs, hid, prev_a, masks, deterministic=True)
    assert rel(value16, torch.from_numpy(g["value"])) <= 2e-2 and rel(hid16, torch.from_numpy(g["new_hidden"])) <= 2e-2


def test_policy_rgbd_with_input_normalisation(golden_dir):
    """a11 with rgb + depth and normalize_visual_inputs=True (resnet_policy.py:61-174): RunningMeanAndVar acts on the
    2x2-pooled input; eval-mode act() and a training-mode forward (statistics update) against the reference golden."""
    g = np.load(os.path.join(golden_dir, "policy_r18_rgbd_norm.npz"))
    pol = helpers.policy_state_dict(device="cuda", vis_types=("rgb", "depth"), normalize=True).eval()
    assert [str(k) for k in g["keys"]] == list(pol.state_dict().keys())
    obs = {"rgb": torch.from_numpy(synth.rgb_frames(3, seed=32).astype(np.float32)).cuda(),
           "depth": torch.from_numpy(synth.depth_frames(3, seed=31)[..., None]).cuda(),
           "pointgoal_with_gps_compass": torch.from_numpy(g["goal"]).cuda()}
    hid, prev_a, masks = (torch.from_numpy(g[k]).cuda() for k in ("hidden", "prev_actions", "masks"))
    with torch.no_grad():
        value, action, logp, new_hid = pol.act(obs, hid, prev_a, masks, deterministic=True)
        plan = list(pol.net._plans.values())[0]
        C = g["encoder_out"].shape[1]
        enc = (plan.feat.float() + plan.lo(plan.feat).float())[..., :C].permute(0, 3, 1, 2)
    print("policy rgbd+norm:", rel(enc, torch.from_numpy(g["encoder_out"])), rel(value, torch.from_numpy(g["value"])))
    assert rel(enc, torch.from_numpy(g["encoder_out"])) <= SPLIT_TOL
    assert rel(value, torch.from_numpy(g["value"])) <= SPLIT_TOL
    assert rel(new_hid, torch.from_numpy(g["new_hidden"])) <= SPLIT_TOL
    assert np.array_equal(action.cpu().numpy(), g["action"])
    pol.train()
    with torch.no_grad():
        pol.net.visual_features(obs)
        plan = [p for p in pol.net._plans.values()][-1]
        enc_tr = (plan.feat.float() + plan.lo(plan.feat).float())[..., :C].permute(0, 3, 1, 2)
    rm = pol.net.visual_encoder.running_mean_and_var
    assert rel(rm._mean, torch.from_numpy(g["train_mean"])) <= 1e-5 and rel(rm._var, torch.from_numpy(g["train_var"])) <= 1e-5
    assert float(rm._count) == float(g["train_count"])
    assert rel(enc_tr, torch.from_numpy(g["train_encoder_out"])) <= SPLIT_TOL


def test_policy_split_precision_within_north_star_tolerance(golden_dir):
    """a11/a12 in split precision (no-grad rollout forwards): encoder features, value and recurrent state against the
    reference's own fp32 outputs within 1e-3."""
    g = np.load(os.path.join(golden_dir, "policy_r18_depth.npz"))
    pol = helpers.policy_state_dict(device="cuda").eval()
    pol.net.set_precision("split")
    dep = torch.from_numpy(synth.depth_frames(3, seed=21)[..., None]).cuda()
    obs = {"depth": dep, "pointgoal_with_gps_compass": torch.from_numpy(g["goal"]).cuda()}
    hid, prev_a, masks = (torch.from_numpy(g[k]).cuda() for k in ("hidden", "prev_actions", "masks"))
    with torch.no_grad():
        value, action, logp, new_hid = pol.act(obs, hid, prev_a, masks, deterministic=True)
        plan = [p for p in pol.net._plans.values() if p.split][0]
        enc = (plan.feat.float() + plan.lo(plan.feat).float())[..., :114].permute(0, 3, 1, 2)
    print("policy split:", rel(enc, torch.from_numpy(g["encoder_out"])), rel(value, torch.from_numpy(g["value"])),
          rel(new_hid, torch.from_numpy(g["new_hidden"])))
    assert rel(enc, torch.from_numpy(g["encoder_out"])) <= SPLIT_TOL
    assert rel(value, torch.from_numpy(g["value"])) <= SPLIT_TOL
    assert rel(new_hid, torch.from_numpy(g["new_hidden"])) <= SPLIT_TOL
    assert np.array_equal(action.cpu().numpy(), g["action"])
    assert np.allclose(logp.cpu().numpy(), g["logp"], atol=1e-4)


def test_dropout_training_mode():
    """nn.Dropout(p) in front of both Linear layers (vo_cnn.py:218,224): identity in eval mode; in training mode a
    fresh mask per forward, kept fraction ~ 1-p, survivors scaled by 1/(1-p), and the backward pass sees the same mask."""
    from pointnav_vo_b200.vo.models import vo_cnn

    name, space, backbone, kw = helpers.VO_CASES["r18_8ch"]
    m = vo_cnn.baseline_registry.get_vo_model(name)(observation_space=space, observation_size=(341, 192),
                                                    hidden_size=512, backbone=backbone, normalize_visual_inputs=True,
                                                    output_dim=3, dropout_p=0.2, **kw)
    m.load_state_dict(helpers.vo_state_dict("r18_8ch"))
    m = m.cuda()
    obs = helpers.vo_inputs(4, 11, space, "cuda")
    m.eval()
    with torch.no_grad():
        y_eval = m(obs)
        y_eval2 = m(obs)
    assert rel(y_eval, y_eval2) <= 6e-3
    m.train()
    y1 = m(obs)
    plan = [p for p in m._plans.values() if p.training and p.dropout_p > 0][0]
    h1 = plan.h32.clone()
    kept_h = (h1 != 0).float().mean().item()
    y1.square().sum().backward()
    dz = plan.dz16.float()
    assert bool(((h1 == 0) <= (dz == 0)).all())  # no gradient flows through dropped / inactive hidden units
    y2 = m(obs)
    assert not torch.equal(plan.h32 == 0, h1 == 0)  # fresh mask
    assert rel(y1, y2) > 1e-2
    with torch.no_grad():
        m.eval()
        h_eval_nonzero = None
        m(obs)
        plan_e = [p for p in m._plans.values() if not p.training][0]
        h_eval_nonzero = (plan_e.h32 != 0).float().mean().item()
    assert abs(kept_h - 0.8 * h_eval_nonzero) <= 0.04


def test_ppo_update_on_device():
    """a13 + a14: rollout storage on the device (GAE through pnvo_gae_scan, bit-exact against the oracle) and one
    PPO update through the B200 actor-critic: finite losses, parameters move, returns equal the oracle's."""
    from pointnav_vo_b200.rl.common.rollout_storage import RolloutStorage
    from pointnav_vo_b200.rl.ppo.ppo import PPO

    pol = helpers.policy_state_dict(device="cuda")
    obs_space, act_space = helpers.policy_spaces()
    T, N = 4, 4
    rs = RolloutStorage(T, N, obs_space, act_space, 512, num_recurrent_layers=pol.net.num_recurrent_layers)
    rs.to("cuda")
    torch.manual_seed(5)
    obs = {"depth": torch.rand(N, 192, 341, 1, device="cuda"), "pointgoal_with_gps_compass": torch.rand(N, 2, device="cuda")}
    rs.observations["depth"][0].copy_(obs["depth"])
    rs.observations["pointgoal_with_gps_compass"][0].copy_(obs["pointgoal_with_gps_compass"])
    rs.masks[0].fill_(1.0)
    pol.eval()
    for t in range(T):
        with torch.no_grad():
            step_obs = {k: v[t] for k, v in rs.observations.items()}
            v, a, lp, h = pol.act(step_obs, rs.recurrent_hidden_states[t], rs.prev_actions[t], rs.masks[t])
        nxt = {"depth": torch.rand(N, 192, 341, 1, device="cuda"),
               "pointgoal_with_gps_compass": torch.rand(N, 2, device="cuda")}
        rs.insert(nxt, h, a, lp, v, torch.randn(N, 1, device="cuda"), (torch.rand(N, 1, device="cuda") > 0.1).float())
    with torch.no_grad():
        nv = pol.get_value({k: v[T] for k, v in rs.observations.items()}, rs.recurrent_hidden_states[T],
                           rs.prev_actions[T], rs.masks[T])
    vp_before = rs.value_preds.clone()
    rs.compute_returns(nv, True, 0.99, 0.95)
    want, _ = po.gae_returns(rs.rewards.cpu().numpy(), vp_before.cpu().numpy(), rs.masks.cpu().numpy(),
                             nv.cpu().numpy(), True, 0.99, 0.95)
    assert np.array_equal(rs.returns.cpu().numpy(), want)
    agent = PPO(pol, clip_param=0.2, ppo_epoch=1, num_mini_batch=2, value_loss_coef=0.5, entropy_coef=0.01, lr=2.5e-4,
                eps=1e-5, max_grad_norm=0.2, use_normalized_advantage=False)
    before = [p.detach().clone() for p in pol.parameters()]
    pol.train()
    vl, al, ent = agent.update(rs)
    assert all(np.isfinite(x) for x in (vl, al, ent)) and ent > 0
    moved = sum(int(not torch.equal(a, b)) for a, b in zip(before, pol.parameters()))
    assert moved >= len(before) - 2
    rs.after_update()


def test_ppo_update_against_reference_golden(golden_dir):
    """a14 / 8f-3: PPO.update through the B200 actor-critic and the fused loss kernel (pnvo_ppo_loss) against the
    UNMODIFIED reference's PPO.update on the same seeded rollout (tests/golden/make_golden.py: gen_ppo; 2 epochs x 2
    minibatches with the reference's recorded env permutations): GAE returns bit-exact, first-minibatch loss terms and
    total <= 1e-3, first-minibatch gradient norms <= 2e-2 per tensor, reported losses, and every parameter after the
    four Adam steps."""
    from pointnav_vo_b200.rl.common.rollout_storage import RolloutStorage
    from pointnav_vo_b200.rl.ppo.ppo import PPO

    g = np.load(os.path.join(golden_dir, "ppo_update.npz"))
    T, N = 4, 4
    pol = helpers.policy_state_dict(device="cuda")
    assert [str(k) for k in g["keys"]] == list(pol.state_dict().keys())
    before = {k: v.detach().clone() for k, v in pol.state_dict().items()}
    obs_space, act_space = helpers.policy_spaces()
    rs = RolloutStorage(T, N, obs_space, act_space, 512, num_recurrent_layers=4)
    rs.to("cuda")
    dep = synth.depth_frames((T + 1) * N, seed=42).reshape(T + 1, N, 192, 341, 1)
    rs.observations["depth"].copy_(torch.from_numpy(dep))
    rs.observations["pointgoal_with_gps_compass"].copy_(torch.from_numpy(g["in/goal"]))
    rs.recurrent_hidden_states[0].copy_(torch.from_numpy(g["in/hidden0"]))
    for k in ("actions", "prev_actions", "masks", "rewards", "value_preds", "action_log_probs"):
        getattr(rs, k).copy_(torch.from_numpy(g["in/" + k]))
    rs.step = T
    rs.compute_returns(torch.from_numpy(g["in/next_value"]).cuda(), True, 0.99, 0.95)
    assert np.array_equal(rs.returns.cpu().numpy(), g["returns"])
    rs.fixed_perms = [torch.from_numpy(p).cuda() for p in g["perms"]]
    agent = PPO(pol, clip_param=0.2, ppo_epoch=2, num_mini_batch=2, value_loss_coef=0.5, entropy_coef=0.01, lr=2.5e-4,
                eps=1e-5, max_grad_norm=0.2, use_normalized_advantage=False)
    first = {}
    real_before_step, real_losses = agent.before_step, agent._losses

    def before_step():
        if "grads" not in first:
            first["grads"] = {k: p.grad.norm().item() for k, p in pol.named_parameters() if p.grad is not None}
        real_before_step()

    def losses(sample):
        out = real_losses(sample)
        first.setdefault("terms", tuple(float(x.detach()) for x in out))
        return out

    agent.before_step, agent._losses = before_step, losses
    pol.train()
    vl, al, ent = agent.update(rs)
    # first minibatch, before any parameter moved: total, value_loss, action_loss, entropy
    t = {k[4:]: torch.from_numpy(np.asarray(g[k])) for k in g.files if k.startswith("mb0/")}
    vl0, al0, ent0 = vo.ppo_losses(t["values"], t["log_probs"], t["entropy"], t["value_preds"], t["returns"],
                                   t["old_log_probs"], t["adv"], 0.2, True)
    total, v0, a0, e0 = first["terms"]
    print("ppo first minibatch:", first["terms"], "reference:", float(g["mb0/total"]), vl0.item(), al0.item(), ent0.item())
    assert abs(total - float(g["mb0/total"])) <= 1e-3 * abs(float(g["mb0/total"]))
    assert abs(v0 - vl0.item()) <= 1e-3 * abs(vl0.item()) and abs(a0 - al0.item()) <= 1e-3 * abs(al0.item())
    assert abs(e0 - ent0.item()) <= 1e-3 * abs(ent0.item())
    worst = 0.0
    for k, n in zip([str(k) for k in g["grad_keys"]], g["grad_norms"]):
        if n > 1e-6:
            worst = max(worst, abs(first["grads"][k] - n) / n)
            assert abs(first["grads"][k] - n) <= 2e-2 * n, (k, first["grads"][k], n)
    print("ppo first-minibatch gradient norms: worst relative deviation", worst)
    ref_losses = g["losses"]
    print("ppo update losses:", (vl, al, ent), "reference:", ref_losses)
    for got, want in zip((vl, al, ent), ref_losses):
        assert abs(got - want) <= 5e-3 * abs(want)
    # parameters after 4 Adam steps: the update of every tensor has the reference's size, and the small tensors
    # (biases, GroupNorm affine parameters: stored in full) move the same way
    after = pol.state_dict()
    for k, dn in zip([str(k) for k in g["keys"]], g["delta_norms"]):
        d = (after[k] - before[k]).float()
        if dn > 0:
            assert abs(d.norm().item() - dn) <= 0.1 * dn, (k, d.norm().item(), dn)
        if "delta/" + k in g.files and dn > 0:
            assert rel_l2(d, torch.from_numpy(g["delta/" + k])) <= 0.25, k


def test_losses_against_reference_golden(golden_dir):
    """a6: PNVO_OP_MSE_LOSS / PNVO_OP_GEO_INV_LOSS against the reference's loss composition (_process_one_batch,
    vo_cnn_regression_geo_invariance_engine.py:676-792; golden from the reference's own functions): plain batches,
    per-data-type means, and the inversion loss restricted to the TURN rows with unpaired MOVE_FORWARD rows in between.
    Value and d(loss)/d(pred) <= 1e-5."""
    from pointnav_vo_b200 import lib as L

    g = np.load(os.path.join(golden_dir, "vo_losses.npz"))
    w = tuple(float(x) for x in g["loss_weights"])
    for name in ("plain", "types_only", "joint"):
        pred = torch.from_numpy(g[f"{name}/pred"]).cuda()
        tgt = torch.from_numpy(g[f"{name}/target"]).cuda()
        acts = torch.from_numpy(g[f"{name}/actions"]).cuda()
        dzm = torch.from_numpy(g[f"{name}/dz_mask"]).cuda()
        types = torch.from_numpy(g[f"{name}/data_types"]).cuda() if f"{name}/data_types" in g.files else None
        B = pred.shape[0]
        dout = torch.zeros(B, 3, device="cuda")
        loss = torch.zeros(3, device="cuda")
        err = torch.zeros(1, dtype=torch.int32, device="cuda")
        ops = [L.op_mse_loss(pred, tgt, dzm, dout, loss, B, 3, w, 1.0, data_types=types)]
        inv_w = float(g[f"{name}/inv_w"])
        if inv_w > 0:
            ops.append(L.op_geo_inv_loss(pred, acts, dout, loss, B, 3, inv_w, 1.0, 1, data_types=types, err=err))
        L.run_ops(ops)
        assert int(err.item()) == 0
        assert abs(loss[0].item() - float(g[f"{name}/loss"])) <= 1e-5 * abs(float(g[f"{name}/loss"])), name
        assert rel(dout, torch.from_numpy(g[f"{name}/grad"])) <= 1e-5, name
    # misaligned TURN rows: the device flags them (the reference asserts, :373-374) and the host check raises
    from pointnav_vo_b200.vo.engine.train_step import FusedVOTrainStep

    bad_a, bad_t = torch.tensor([2, 2, 3, 3]).cuda(), torch.tensor([0, 0, 1, 1]).cuda()
    pred = torch.zeros(4, 3, device="cuda")
    L.run_ops([L.op_geo_inv_loss(pred, bad_a, None, loss, 4, 3, 1.0, 1.0, 1, data_types=bad_t, err=err)])
    assert int(err.item()) == 1
    with pytest.raises(L.PnvoError):
        FusedVOTrainStep._check_pairing(bad_a.cpu().numpy(), bad_t.cpu().numpy())


def test_fused_train_step_matches_reference_loss_composition():
    """The whole fused step on an inverse_joint_train batch (pair map with unpaired MOVE_FORWARD rows between the TURN
    pairs): the loss it reports and the gradient it feeds to the backward program equal the oracle of the reference's
    _process_one_batch evaluated on the step's own predictions."""
    from pointnav_vo_b200.vo.dataset import geo_invariance as gi
    from pointnav_vo_b200.vo.engine.train_step import FusedVOTrainStep

    model, space, _ = _load_vo("r18_8ch")
    obs = helpers.vo_inputs(4, 23, space, "cuda")
    rgb, dep = obs["rgb"].to(torch.uint8).contiguous(), obs["depth"].contiguous()
    pm = gi.make_pair_map([2, 1, 3, 2], act_type=2, geo_invariance_types=("inverse_joint_train",))
    assert list(pm["actions"]) == [2, 3, 1, 3, 2, 2, 3] and list(pm["data_types"]) == [0, 1, 0, 0, 1, 0, 1]
    rng = np.random.default_rng(5)
    tg = torch.from_numpy(gi.expand_targets(rng.normal(0, 0.2, (4, 3)).astype(np.float32), pm)).cuda()
    batch = {"rgb": rgb, "depth": dep, "pair_map": torch.from_numpy(pm["pair_map"]).cuda()}
    model.train()
    tr = FusedVOTrainStep(model, loss_weights=(1.0, 2.0, 0.5), loss_inv_weight=0.7)
    loss = tr.step(batch, tg, actions=pm["actions"], data_types=pm["data_types"])
    plan = tr._plan
    pred = plan.out.detach().float().cpu().requires_grad_(True)
    want = vo.vo_total_loss(pred, tg.cpu(), torch.from_numpy(pm["actions"]), torch.from_numpy(pm["data_types"]),
                            (1.0, 2.0, 0.5), None, 0.7)
    want.backward()
    assert abs(loss.item() - want.item()) <= 1e-5 * abs(want.item())
    assert rel(plan.dout, pred.grad) <= 1e-5
    tr.check()
    with pytest.raises(Exception):  # rows out of order: TURN rows no longer alternate
        tr.step(batch, tg, actions=pm["actions"][::-1].copy(), data_types=pm["data_types"])


def test_vo_inference_in_the_rl_loop():
    """a9: `_compute_local_delta_states_from_vo` (reference signature, numpy observations in, python lists out) against the
    oracle chain discretise -> top-down -> VO net, and the batched variant (raw uint8 path) against the per-env calls."""
    import types

    from pointnav_vo_b200.rl.common.base_trainer_with_vo import VOInferenceMixin

    case = "r18_30ch"
    name, space, backbone, kw = helpers.VO_CASES[case]
    sd = helpers.vo_state_dict(case)
    models = {}
    for act in ("forward", "left", "right"):
        m, _, _ = _load_vo(case)
        models[act] = m.eval()
    ns = types.SimpleNamespace
    trainer = VOInferenceMixin()
    trainer.device = torch.device("cuda")
    trainer.vo_model = models
    trainer.config = ns(VO=ns(VO_TYPE="REGRESS", VIS_SIZE_H=192, VIS_SIZE_W=341,
                              REGRESS_MODEL=ns(name=name, discretize_depth="hard", discretized_depth_channels=10,
                                               regress_type="sep_act", mode="det", rnd_mode_n=2)),
                        TASK_CONFIG=ns(SIMULATOR=ns(DEPTH_SENSOR=ns(MIN_DEPTH=0.1, MAX_DEPTH=10.0, HFOV=70))))
    trainer._setup_vo_preproc()
    rgb = synth.rgb_frames(6, seed=31)
    dep = synth.depth_frames(6, seed=32)[..., None]
    prev = [{"rgb": rgb[2 * i], "depth": dep[2 * i]} for i in range(3)]
    cur = [{"rgb": rgb[2 * i + 1], "depth": dep[2 * i + 1]} for i in range(3)]
    acts = [1, 2, 3]
    orc = po.TopDownOracle()
    singles = []
    for p, c, a in zip(prev, cur, acts):
        deltas, std, extra = trainer._compute_local_delta_states_from_vo(p, c, a)
        assert isinstance(deltas, list) and len(deltas) == 3 and std == [0, 0, 0] and extra == {}
        d2 = np.stack([p["depth"][..., 0], c["depth"][..., 0]])
        oh = po.discretize_depth_onehot(d2)
        obs = {"rgb": torch.from_numpy(np.concatenate([p["rgb"], c["rgb"]], -1)[None].astype(np.float32)),
               "depth": torch.from_numpy(np.stack([d2[0], d2[1]], -1)[None]),
               "discretized_depth": torch.from_numpy(np.concatenate([oh[0], oh[1]], -1)[None]),
               "top_down_view": torch.from_numpy(np.stack([orc.gen_top_down_view(d2[j])[..., 0] for j in range(2)], -1)[None])}
        ref, _ = vo.vo_forward(obs, sd, space, backbone, training=False)
        # the mixin runs the VO nets in split precision by default: inside the north-star 1e-3 of the fp32 oracle
        assert rel(torch.tensor(deltas)[None], ref) <= 1e-3
        singles.append(deltas)
    batched = trainer.compute_local_delta_states_batched(prev, cur, acts)
    assert batched.shape == (3, 3)
    assert rel(batched, torch.tensor(np.array(singles))) <= 1e-3
    trainer.vo_precision = "fp16"  # the single-pass mode stays available
    fast = trainer.compute_local_delta_states_batched(prev, cur, acts)
    assert rel(fast, torch.tensor(np.array(singles))) <= 8e-3


def test_train_step_with_prefetched_input_pipeline():
    """Double-buffered input staging: running the NEXT batch's input pipeline on a side stream (alternate staging buffer,
    pointer-patched forward / backward programs) must not change the training trajectory."""
    import copy

    from pointnav_vo_b200.vo.engine.train_step import FusedVOTrainStep

    m1, space, _ = _load_vo("r18_30ch")
    m2 = copy.deepcopy(m1)
    m1.train()
    m2.train()
    count0 = float(m1.visual_encoder.running_mean_and_var._count)
    batches = []
    for k in range(4):
        o = helpers.vo_inputs(2, 40 + k, space, "cuda")
        batches.append(({"rgb": o["rgb"].to(torch.uint8).contiguous(), "depth": o["depth"].contiguous()},
                        torch.randn(2, 3, device="cuda") * 0.1))
    s1, s2 = FusedVOTrainStep(m1), FusedVOTrainStep(m2)
    l1, l2 = [], []
    for k, (obs, tgt) in enumerate(batches):
        l1.append(s1.step(obs, tgt).item())
        nxt = batches[k + 1][0] if k + 1 < len(batches) else None
        l2.append(s2.step(obs, tgt, prefetch=nxt).item())
        # dW is flushed with fp32 atomics (last-bit differences) and Adam's first steps amplify them (measured: 0, 5e-5,
        # 6e-4, 5e-2 relative loss difference over four free-running steps), so the second trainer is put back on the
        # first one's weights / moments after every step: each step's forward must then agree to rounding
        torch.cuda.synchronize()
        drift = ((s2._flat - s1._flat).abs().max() / s1._flat.abs().max()).item()
        assert drift <= 2 * s1.lr + 1e-6, drift  # one Adam step moves a weight by at most ~lr
        s2._flat.copy_(s1._flat)
        s2._m.copy_(s1._m)
        s2._v.copy_(s1._v)
        m2._packed_version = None
    torch.cuda.synchronize()
    for a, b in zip(l1, l2):
        assert abs(a - b) <= 1e-5 * abs(a) + 1e-7, (l1, l2)
    r1, r2 = m1.visual_encoder.running_mean_and_var, m2.visual_encoder.running_mean_and_var
    assert float(r1._count) == float(r2._count) == count0 + 8.0
    assert rel(r2._mean, r1._mean) <= 1e-6 and rel(r2._var, r1._var) <= 1e-6


def test_fused_train_step_state_dict_resumes():
    """The fused step's optimiser state in torch.optim.Adam's format (a run trained on the fused path must
    resume with its moments and bias-correction step): 2 steps + save + load into a fresh trainer + 1 step == 3 steps."""
    import copy

    from pointnav_vo_b200.vo.engine.train_step import FusedVOTrainStep

    case = "r18_8ch"
    m1, space, _ = _load_vo(case)
    obs = [helpers.vo_inputs(2, 31 + i, space, "cuda") for i in range(3)]
    tgt = [torch.randn(2, 3, device="cuda", generator=torch.Generator("cuda").manual_seed(i)) * 0.1 for i in range(3)]
    m1.train()
    t1 = FusedVOTrainStep(m1, lr=2.5e-4, eps=1e-8)
    for i in range(2):
        t1.step(obs[i], tgt[i])
    sd_opt = t1.state_dict()
    m2, _, _ = _load_vo(case)     # (a module that has run holds ctypes programs: build a fresh one instead of deepcopy)
    m2.load_state_dict(copy.deepcopy(m1.state_dict()))
    m2.train()
    t2 = FusedVOTrainStep(m2, lr=1.0)      # hyper-parameters come from the file
    t2.load_state_dict(sd_opt, example_obs=obs[2])
    assert t2.step_count == 2 and t2.lr == 2.5e-4
    t1.step(obs[2], tgt[2])
    t2.step(obs[2], tgt[2])
    for (k, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
        d = (a - b).abs()
        assert d.max().item() <= 2 * 2.5e-4 + 1e-7, k        # same bound as the autograd comparison: fp16-noise sign flips
        assert (d > 1e-5).float().mean().item() <= 0.05, k
    # the file loads into a plain torch.optim.Adam over the same parameters
    opt = torch.optim.Adam(m1.parameters(), lr=1.0)
    opt.load_state_dict(sd_opt)
    assert opt.param_groups[0]["lr"] == 2.5e-4


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs on one node")
def test_peer_memory_gradient_exchange_matches_nccl_world2():
    """8e: the fused reduce-scatter + Adam + all-gather kernel over NVLink peer memory against NCCL all-reduce + Adam."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(root, "tests", "multigpu", "peer_adam_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PEER_ADAM_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
