"""Per-step VO inference for the RL loop: drop-in for BaseRLTrainerWithVO's VO helpers
(pointnav_vo/rl/common/base_trainer_with_vo.py:37-314), running on libpnvo.

`VOInferenceMixin` provides, with the reference's names / signatures / return values:
    _setup_vo_preproc()                       the relevant part of _setup_vo_model (:101-128)
    _discretize_depth_func(raw_depth)         (:135-167)
    _compute_local_delta_states_from_vo(prev_obs, cur_obs, act, vis_video=False)   (:169-314)
and one addition:
    compute_local_delta_states_batched(prev_obs_list, cur_obs_list, acts)  -> [N, 3] tensor
which groups the environments by action and runs discretisation, top-down projection and the VO net
once per group instead of once per environment.

A maintainer switches the reference over by mixing this class in front of BaseRLTrainerWithVO (see
INTEGRATION.md); `self.config`, `self.device` and `self.vo_model` are the reference's own attributes.
"""
import numpy as np
import torch

from ...utils.geometry_utils import (NormalizedDepth2TopDownViewHabitatTorch, discretize_depth, discretize_end_vals)
from ...vo.common.common_vars import ACT_IDX2NAME


class VOInferenceMixin:
    _vo_obs_transformer = None
    # Precision of the per-step VO forward (VisualOdometryCNNBase.set_precision).  "split" = value + residual fp16
    # planes, three tensor-core products per convolution: the deltas agree with the reference's fp32 network to ~1e-4
    # relative, and at rollout batch sizes (one to a few dozen pairs) the step is launch-bound either way.
    vo_precision = "split"

    # ---- base_trainer_with_vo.py:101-128 ------------------------------------------------------
    def _setup_vo_preproc(self):
        cfg = self.config.VO.REGRESS_MODEL
        if "discretize_depth" in cfg.name or "dd" in cfg.name:
            if cfg.discretize_depth in ["hard"]:
                self._discretized_depth_end_vals = discretize_end_vals(cfg.discretized_depth_channels)
            else:
                raise NotImplementedError
        if "top_down" in cfg.name:
            sensor = self.config.TASK_CONFIG.SIMULATOR.DEPTH_SENSOR
            self._top_down_view_generator = NormalizedDepth2TopDownViewHabitatTorch(
                min_depth=sensor.MIN_DEPTH, max_depth=sensor.MAX_DEPTH, vis_size_h=self.config.VO.VIS_SIZE_H,
                vis_size_w=self.config.VO.VIS_SIZE_W, hfov_rad=sensor.HFOV)  # HFOV passed verbatim, as the reference does

    # ---- base_trainer_with_vo.py:135-167 ------------------------------------------------------
    def _discretize_depth_func(self, raw_depth):
        cfg = self.config.VO.REGRESS_MODEL
        if cfg.discretize_depth != "hard":
            raise NotImplementedError
        # the kernel counts out-of-range pixels; the assert mirrors the reference's two asserts
        return discretize_depth(raw_depth, cfg.discretized_depth_channels, self._discretized_depth_end_vals, check=True)

    # ---- shared ---------------------------------------------------------------------------------
    def _vo_inputs(self, rgb_pair, depth_pair):
        """rgb_pair [N,H,W,6], depth_pair [N,H,W,2] fp32 CUDA -> the VO model's observation dict."""
        cfg = self.config.VO.REGRESS_MODEL
        obs_pairs = {"rgb": rgb_pair, "depth": depth_pair}
        n, h, w = depth_pair.shape[:3]
        if "discretize_depth" in cfg.name or "dd" in cfg.name:
            assert depth_pair.size(-1) == 2
            c = cfg.discretized_depth_channels
            dd = torch.empty((n, h, w, 2 * c), dtype=torch.float32, device=depth_pair.device)
            for k in range(2):
                d = depth_pair[..., k].contiguous()
                discretize_depth(d, c, self._discretized_depth_end_vals, check=True, out=dd[..., k * c:(k + 1) * c])
            obs_pairs["discretized_depth"] = dd
        if "top_down" in cfg.name:
            td = torch.empty((n, h, w, 2), dtype=torch.float32, device=depth_pair.device)
            for k in range(2):
                self._top_down_view_generator.gen_top_down_view(depth_pair[..., k].contiguous(), out=td[..., k:k + 1])
            obs_pairs["top_down_view"] = td
        return obs_pairs

    def _vo_key(self, act):
        return "all" if self.config.VO.REGRESS_MODEL.regress_type == "unified_act" else ACT_IDX2NAME[act]

    # ---- base_trainer_with_vo.py:169-314 ------------------------------------------------------
    def _compute_local_delta_states_from_vo(self, prev_obs, cur_obs, act, vis_video=False):
        dev = self.device
        rgb_pair = torch.cat([torch.as_tensor(np.asarray(prev_obs["rgb"]), dtype=torch.float32, device=dev),
                              torch.as_tensor(np.asarray(cur_obs["rgb"]), dtype=torch.float32, device=dev)],
                             dim=2).unsqueeze(0)
        depth_pair = torch.cat([torch.as_tensor(np.asarray(prev_obs["depth"]), dtype=torch.float32, device=dev),
                                torch.as_tensor(np.asarray(cur_obs["depth"]), dtype=torch.float32, device=dev)],
                               dim=2).unsqueeze(0)
        if self._vo_obs_transformer is not None:
            raise NotImplementedError("VO.OBS_TRANSFORM other than 'none' is not supported on the B200 path "
                                      "(the shipped configs use none: configs/rl/ddppo_pointnav.yaml:43,99)")
        if self.config.VO.VO_TYPE != "REGRESS":
            raise NotImplementedError
        obs_pairs = self._vo_inputs(rgb_pair.contiguous(), depth_pair.contiguous())
        extra_infos = {}
        if vis_video:
            extra_infos["ego_top_down_map"] = obs_pairs["top_down_view"][0, :, :, 1:2]
        cfg = self.config.VO.REGRESS_MODEL
        model = self.vo_model[self._vo_key(act)]
        model.set_precision(self.vo_precision)
        local_delta_states, local_delta_states_std = [], []
        with torch.no_grad():
            if cfg.mode == "det":
                model.eval()
                if "act_embed" in cfg.name:
                    actions = torch.tensor([act], dtype=torch.long, device=dev)
                    tmp_deltas = model(obs_pairs, actions)
                else:
                    tmp_deltas = model(obs_pairs)
                local_delta_states = list(tmp_deltas.cpu().numpy()[0, :])
                local_delta_states_std = [0, 0, 0]
            elif cfg.mode == "rnd":
                model.train()
                samples = [model(obs_pairs).cpu().numpy()[0, :] for _ in range(cfg.rnd_mode_n)]
                local_delta_states = list(np.mean(np.array(samples), axis=0))
                local_delta_states_std = list(np.std(samples, axis=0))
        return local_delta_states, local_delta_states_std, extra_infos

    # ---- batched variant (new) --------------------------------------------------------------------
    def compute_local_delta_states_batched(self, prev_obs_list, cur_obs_list, acts):
        """All environments of one rollout step at once.  Returns a CUDA fp32 tensor [N, 3] of (dx, dz, dyaw);
        row i equals what _compute_local_delta_states_from_vo(prev_obs_list[i], cur_obs_list[i], acts[i])
        returns in 'det' mode."""
        dev = self.device
        n = len(acts)
        rgb_np = np.stack([np.concatenate([np.asarray(p["rgb"]), np.asarray(c["rgb"])], axis=2)
                           for p, c in zip(prev_obs_list, cur_obs_list)])
        dep = torch.as_tensor(np.stack([np.concatenate([np.asarray(p["depth"]), np.asarray(c["depth"])], axis=2)
                                        for p, c in zip(prev_obs_list, cur_obs_list)])).to(dev).float().contiguous()
        cfg = self.config.VO.REGRESS_MODEL
        if rgb_np.dtype == np.uint8 and cfg.discretize_depth == "hard":
            # raw pairs: the simulator's uint8 rgb goes to the device as is (6 B/pixel instead of 24) and the model
            # derives the discretised-depth / top-down channels itself (csrc/raw_input.cu), with this trainer's geometry
            obs = {"rgb": torch.as_tensor(rgb_np).to(dev).contiguous(), "depth": dep}
            for m in self.vo_model.values():
                m.set_raw_input_config(getattr(self, "_top_down_view_generator", None),
                                       getattr(self, "_discretized_depth_end_vals", None))
        else:
            obs = self._vo_inputs(torch.as_tensor(rgb_np).to(dev).float().contiguous(), dep)
        out = torch.empty((n, 3), dtype=torch.float32, device=dev)
        with torch.no_grad():
            for key in sorted({self._vo_key(int(a)) for a in acts}):
                idx = torch.tensor([i for i, a in enumerate(acts) if self._vo_key(int(a)) == key], device=dev)
                model = self.vo_model[key]
                model.set_precision(self.vo_precision)
                model.eval()
                sub = {k: v.index_select(0, idx).contiguous() for k, v in obs.items()}
                out.index_copy_(0, idx, model(sub)[:, :3])
        return out
