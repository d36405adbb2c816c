"""Host mirror of the reference's depth preprocessing interface, running on libpnvo's CUDA kernels.

  NormalizedDepth2TopDownViewHabitatTorch   pointnav_vo/utils/geometry_utils.py:491-721
  discretize_depth                          pointnav_vo/rl/common/base_trainer_with_vo.py:135-167
  compute_goal_pos_batched                  pointnav_vo/utils/geometry_utils.py:115-144 (batched, on device)

Same constructor arguments / call signatures as the reference; additionally every entry point accepts a
leading batch dimension so that a whole step's frames go through one kernel launch.
"""
import ctypes
import math

import numpy as np
import torch

from .. import lib as _lib


def discretize_end_vals(n_channels):
    """base_trainer_with_vo.py:107-117: python-double edges i/n plus the closing 1.0."""
    return [i * 1.0 / n_channels for i in range(n_channels)] + [1.0]


_edge_cache = {}


def _edges(end_vals, device):
    key = (tuple(end_vals), str(device))
    if key not in _edge_cache:
        # torch compares an fp32 tensor with a python double in fp32: round each edge once
        _edge_cache[key] = torch.tensor(np.asarray(end_vals, dtype=np.float32), device=device)
    return _edge_cache[key]


def discretize_depth(raw_depth, n_channels=10, end_vals=None, check=True, out=None):
    """raw_depth: CUDA fp32 tensor [...] in [0, 1] -> fp32 one-hot [..., n_channels]
    (base_trainer_with_vo.py:135-167; `check` reproduces its asserts with one device counter)."""
    if end_vals is None:
        end_vals = discretize_end_vals(n_channels)
    assert len(end_vals) == n_channels + 1
    d = raw_depth.contiguous().float()
    stride = n_channels
    if out is None:
        out = torch.empty((*d.shape, n_channels), dtype=torch.float32, device=d.device)
    else:
        # a channel slice of a wider NHWC tensor: rows of n_channels floats, `stride` floats apart
        assert out.dtype == torch.float32 and out.shape == (*d.shape, n_channels) and out.stride(-1) == 1
        stride = out.stride(-2)
        exp = stride
        for k in range(out.dim() - 2, -1, -1):
            assert out.stride(k) == exp, "out must be a uniform-stride channel slice"
            exp *= out.shape[k]
    err = torch.zeros(1, dtype=torch.int32, device=d.device) if check else None
    lib = _lib.load()
    _lib.check(lib.pnvo_discretize_depth(_lib.ptr(d), d.numel(), _lib.ptr(_edges(end_vals, d.device)), n_channels,
                                         _lib.ptr(out), stride, None, _lib.ptr(err), _lib.stream_ptr(d.device)))
    if check:
        assert int(err.item()) == 0, "depth outside [0, 1]"  # :136-137
    return out


def discretize_depth_index(raw_depth, n_channels=10, end_vals=None):
    """uint8 bin index per pixel (255 = outside [0,1]); the compact form of the one-hot map."""
    if end_vals is None:
        end_vals = discretize_end_vals(n_channels)
    d = raw_depth.contiguous().float()
    idx = torch.empty(d.shape, dtype=torch.uint8, device=d.device)
    lib = _lib.load()
    _lib.check(lib.pnvo_discretize_depth(_lib.ptr(d), d.numel(), _lib.ptr(_edges(end_vals, d.device)), n_channels,
                                         None, 0, _lib.ptr(idx), None, _lib.stream_ptr(d.device)))
    return idx


class NormalizedDepth2TopDownViewHabitatTorch:
    """geometry_utils.py:491-721.  hfov_rad is used verbatim (the reference's callers pass degrees)."""

    def __init__(self, min_depth, max_depth, vis_size_h, vis_size_w, hfov_rad, ksize=3, rows_around_center=50,
                 flag_center_crop=True):
        if ksize != 3:
            raise NotImplementedError("only the reference's 3x3 blur is implemented")
        self._epsilon = 0.01
        self._min_depth, self._max_depth = min_depth, max_depth
        self._vis_size_h, self._vis_size_w = vis_size_h, vis_size_w
        self._hfov_rad = hfov_rad
        self._ksize = ksize
        self._rows_around_center = rows_around_center
        self._flag_center_crop = flag_center_crop
        self._get_intrinsic_mat()
        self._dev_ray = {}

    def _get_intrinsic_mat(self):
        """geometry_utils.py:562-583,648-650,678-682 in explicit fp32 steps.
        K = [[f,0,W/2],[0,f,H/2],[0,0,1]] (fp32); Kinv row 0 of this triangular matrix is
        [1/f, 0, -(W/2)/f]; ray[u] = fl(fl(Kinv00*(u+.5)) + Kinv02)."""
        F = np.float32
        W, H = self._vis_size_w, self._vis_size_h
        f = F((W / 2) / math.tan(self._hfov_rad / 2))
        u0 = F(W / 2)
        k00 = F(1.0) / f
        k02 = -(u0 / f)
        xr = F(k00 * F(W - 0.5)) + k02
        max_x = F(xr * F(self._max_depth))
        self._min_x = -max_x
        self._x_range = F(max_x - self._min_x)
        self._consts = _lib.TopdownConsts(
            min_x=float(self._min_x), x_den=float(F(self._x_range * F(1 + self._epsilon))),
            z_den=float(F((self._max_depth - self._min_depth) * (1 + self._epsilon))),
            depth_scale=float(F(self._max_depth - self._min_depth)), depth_off=float(F(self._min_depth)),
            rows_around_center=int(self._rows_around_center), center_crop=int(bool(self._flag_center_crop)))
        u = np.arange(W, dtype=F) + F(0.5)
        self._ray = ((k00 * u).astype(F) + k02).astype(F)

    def _ray_on(self, device):
        key = str(device)
        if key not in self._dev_ray:
            self._dev_ray[key] = torch.from_numpy(self._ray).to(device)
        return self._dev_ray[key]

    def gen_top_down_view(self, normalized_depth, out=None, return_counts=False):
        """normalized_depth: CUDA fp32 [H, W, 1] (reference signature) or batched [N, H, W, 1] / [N, H, W]
        -> [H, W, 1] or [N, H, W, 1] fp32 in [0, 1]."""
        H, W = self._vis_size_h, self._vis_size_w
        d = normalized_depth
        single = d.dim() == 3 and d.shape[-1] == 1
        d = d.reshape(-1, H, W).contiguous().float()
        n = d.shape[0]
        if out is None:
            out = torch.empty((n, H, W, 1), dtype=torch.float32, device=d.device)
            frame_stride, pix_stride = H * W, 1
        else:
            # out: [n, H, W, C] view of a channel of a wider NHWC tensor
            frame_stride, pix_stride = out.stride(0), out.stride(2)
        cnt = torch.empty((n, H, W), dtype=torch.int32, device=d.device) if return_counts else None
        lib = _lib.load()
        _lib.check(lib.pnvo_topdown_project(_lib.ptr(d), H * W, n, H, W, _lib.ptr(self._ray_on(d.device)),
                                            ctypes.byref(self._consts), _lib.ptr(out), frame_stride, pix_stride,
                                            _lib.ptr(cnt), _lib.stream_ptr(d.device)))
        res = out[0] if single else out
        return (res, cnt) if return_counts else res


def gen_top_down_view_pairs(generator, depth_pairs, out=None):
    """Top-down maps of both frames of every depth pair in ONE launch, reading the interleaved [B, H, W, 2]
    tensor in place (no per-frame copies) and writing [B, H, W, 2] (channel 0 = prev, 1 = cur): what
    base_trainer_with_vo.py:232-269 / regression_geo_invariance_iter_dataset.py:251-267 compute frame by frame."""
    H, W = generator._vis_size_h, generator._vis_size_w
    d = depth_pairs
    # fp16 pairs (the dataset's storage type) are widened exactly inside the kernel
    assert d.is_cuda and d.dtype in (torch.float32, torch.float16) and d.is_contiguous() and d.shape[1:] == (H, W, 2)
    B = d.shape[0]
    if out is None:
        out = torch.empty((B, H, W, 2), dtype=torch.float32, device=d.device)
    assert out.is_contiguous() and out.shape == (B, H, W, 2) and out.dtype == torch.float32
    lib = _lib.load()
    fn = lib.pnvo_topdown_project_strided_f16 if d.dtype == torch.float16 else lib.pnvo_topdown_project_strided
    _lib.check(fn(_lib.ptr(d), 2 * H * W, 2, 2 * B, H, W, _lib.ptr(generator._ray_on(d.device)),
                  ctypes.byref(generator._consts), _lib.ptr(out), 2 * H * W, 2, None, _lib.stream_ptr(d.device)))
    return out


def compute_goal_pos_batched(prev_goal_xyz, local_delta_states):
    """geometry_utils.py:115-144 for n agents at once.  prev_goal_xyz: CUDA fp64 [n,3] (updated in place),
    local_delta_states: CUDA fp32 [n,3] = (dx, dz, dyaw).  Returns {"cartesian": [n,3] f64, "polar": [n,2] f32}."""
    g = prev_goal_xyz
    assert g.dtype == torch.float64 and g.is_contiguous()
    d = local_delta_states.contiguous().float()
    polar = torch.empty((g.shape[0], 2), dtype=torch.float32, device=g.device)
    lib = _lib.load()
    _lib.check(lib.pnvo_goal_update(_lib.ptr(g), _lib.ptr(d), _lib.ptr(polar), g.shape[0], _lib.stream_ptr(g.device)))
    return {"cartesian": g, "polar": polar}
