import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from tests import helpers
from tests.test_gpu_parity import _load_vo, rel, rel_l2
from oracle import vo_oracle as vo
for case in ("r18_30ch", "r18_8ch"):
    g = np.load(f"tests/golden/vo_{case}.npz")
    for exact in (True, False):
        m, space, backbone = _load_vo(case)
        m.exact_stem = exact
        obs = helpers.vo_inputs(2, 11, space, "cuda")
        raw = {"rgb": obs["rgb"].to(torch.uint8).contiguous(), "depth": obs["depth"].half().contiguous()}
        m.train()
        y = m(raw)
        sum(vo.vo_losses(y, torch.from_numpy(g["target"]).cuda())).backward()
        k1 = "visual_encoder.backbone.conv1.0.weight"
        got = m.visual_encoder.backbone.conv1[0].weight.grad.cpu()
        ref = torch.from_numpy(g["grad/" + k1])
        C = got.shape[1]; cf = C // 2
        print(case, "exact" if exact else "lo-plane", "fwd", rel(y, torch.from_numpy(g["train_out"])), "conv1 grad rel-L2", rel_l2(got, ref))
        for c in range(cf):
            cc = [c, c + cf]
            print("   ch %2d: rel-L2 %.4f  |ref| %.3e" % (c, rel_l2(got[:, cc], ref[:, cc]), ref[:, cc].norm().item()))
