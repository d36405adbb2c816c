#!/usr/bin/env python
"""Effect of the residual-weight product of the exact-input stem (engine.STEM_WLO) on the network output: max|d| / rms
against the reference's fp32 golden outputs, eval and training mode, with and without the second stem launch, plus seeded
random batches against the fp32 oracle (tolerance: BASELINE.json north star 1e-3)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vo_oracle as vo  # noqa: E402
from pointnav_vo_b200 import engine  # noqa: E402
from pointnav_vo_b200.vo.models import vo_cnn  # noqa: E402
from tests import helpers  # noqa: E402


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.pow(2).mean().sqrt()).item()


def load(case):
    name, space, backbone, kw = helpers.VO_CASES[case]
    m = vo_cnn.baseline_registry.get_vo_model(name)(
        observation_space=space, observation_size=(341, 192), hidden_size=512, backbone=backbone,
        normalize_visual_inputs=True, output_dim=3, dropout_p=0.0, **kw)
    sd = helpers.vo_state_dict(case)
    m.load_state_dict(sd)
    return m.cuda(), space, backbone, sd


for wlo in (True, False):
    engine.STEM_WLO = wlo
    for case in ("r18_30ch", "r18_8ch"):
        g = np.load(os.path.join(ROOT, "tests", "golden", f"vo_{case}.npz"))
        m, space, backbone, sd = load(case)
        obs = helpers.vo_inputs(2, 11, space, "cuda")
        raw = {"rgb": obs["rgb"].to(torch.uint8).contiguous(), "depth": obs["depth"].half().contiguous()}
        m.eval()
        with torch.no_grad():
            y = m(raw)
        e_eval = rel(y, torch.from_numpy(g["eval_out"]))
        m.train()
        y = m(raw)
        e_train = rel(y, torch.from_numpy(g["train_out"]))
        # more samples: 5 seeded batches of 8 pairs against the fp32 oracle (eval mode)
        m.eval()
        worst, rms_err = 0.0, []
        for seed in range(5):
            o = helpers.vo_inputs(8, 100 + seed, space, "cuda")
            r = {"rgb": o["rgb"].to(torch.uint8).contiguous(), "depth": o["depth"].half().contiguous()}
            with torch.no_grad():
                yy = m(r)
            ref, _ = vo.vo_forward({k: v.cpu() for k, v in o.items()}, sd, space, backbone, training=False)
            worst = max(worst, rel(yy, ref))
            rms_err.append(((yy.cpu() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item())
        print(f"STEM_WLO={int(wlo)} {case}: golden eval {e_eval:.2e} train {e_train:.2e} | 40 random pairs: max|d|/rms {worst:.2e}, "
              f"rms(d)/rms {np.mean(rms_err):.2e}", flush=True)
