#!/usr/bin/env python
"""One VO training step (B=256 by default) between cudaProfilerStart/Stop, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--fwd-only", action="store_true")
    a = ap.parse_args()
    from pointnav_vo_b200.vo.engine.train_step import FusedVOTrainStep

    dev = torch.device("cuda", 0)
    model = bench.build_model(dev)
    tr = FusedVOTrainStep(model)
    rgb, dep, tgt = bench.synth_batch(a.batch, 1)
    pre = bench.DevicePreproc(a.batch, dev)
    obs = pre(torch.from_numpy(rgb).to(dev), torch.from_numpy(dep).to(dev))
    t = torch.from_numpy(tgt).to(dev)
    for _ in range(a.warmup):
        tr.step(obs, t)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(a.steps):
        if a.fwd_only:
            model._run_forward(tr._plan, obs, True)
        else:
            tr.step(obs, t)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
