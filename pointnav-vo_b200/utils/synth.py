"""Seeded synthetic inputs for parity tests and bench.py.

numpy only (default_rng is bit-reproducible across hosts); the same generator feeds the oracle, the
golden-vector script and the CUDA path.  There is no dataset in this environment (the reference's
HDF5 sets need the Habitat simulator, TRAIN.md:21), so every benchmark says data = "synthetic".
"""
import numpy as np

H, W = 192, 341  # configs/vo/vo_pointnav.yaml:28-29 (VIS_SIZE_H, VIS_SIZE_W)


def _upsample_bilinear(grid, h, w):
    gh, gw = grid.shape
    ys = np.linspace(0.0, gh - 1.0, h)
    xs = np.linspace(0.0, gw - 1.0, w)
    y0 = np.minimum(ys.astype(np.int64), gh - 2)
    x0 = np.minimum(xs.astype(np.int64), gw - 2)
    fy = (ys - y0)[:, None]
    fx = (xs - x0)[None, :]
    a = grid[y0][:, x0]
    b = grid[y0][:, x0 + 1]
    c = grid[y0 + 1][:, x0]
    d = grid[y0 + 1][:, x0 + 1]
    return (a * (1 - fy) * (1 - fx) + b * (1 - fy) * fx + c * fy * (1 - fx) + d * fy * fx)


def depth_frames(n, seed=1, h=H, w=W, borders=True):
    """[n, h, w] fp32 depth in [0,1]: smooth field rounded through fp16 (the dataset stores fp16,
    generate_datasets.py:272), ~25 % of frames with zero left/top/right/bottom borders and ~2 % zero holes."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, h, w), dtype=np.float32)
    for i in range(n):
        g = rng.random((6, 11))
        d = np.clip(_upsample_bilinear(g, h, w) * 1.1 - 0.05, 0.0, 1.0)
        d = d.astype(np.float16).astype(np.float32)
        if borders:
            holes = rng.random((h, w)) < 0.02
            d[holes] = 0.0
            if rng.random() < 0.25:
                t, l, b, r = rng.integers(0, 24, size=4)
                d[:t, :] = 0
                d[:, :l] = 0
                if b:
                    d[h - b:, :] = 0
                if r:
                    d[:, w - r:] = 0
        out[i] = d
    return out


def rgb_frames(n, seed=2, h=H, w=W):
    """[n, h, w, 3] uint8 uniform."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)


def gae_inputs(T=128, N=128, seed=3):
    """rewards~N(0,1) [T,N,1], values~N(0,1) [T+1,N,1], masks~Bernoulli(.98) [T+1,N,1], next_value [N,1]."""
    rng = np.random.default_rng(seed)
    rewards = rng.standard_normal((T, N, 1)).astype(np.float32)
    values = rng.standard_normal((T + 1, N, 1)).astype(np.float32)
    masks = (rng.random((T + 1, N, 1)) < 0.98).astype(np.float32)
    next_value = rng.standard_normal((N, 1)).astype(np.float32)
    return rewards, values, masks, next_value


def fill_state_dict(state_dict, seed=0):
    """Deterministic weights for any module: same numbers on every host / torch version.
    Conv/Linear weights ~ U(-b, b) with b = 1/sqrt(fan_in) (the reference's default init scale),
    GroupNorm weight ~ U(.5, 1.5), biases ~ U(-.1, .1); running stats get plausible non-trivial values.
    Returns {key: np.ndarray} in the state_dict's key order."""
    rng = np.random.default_rng(seed)
    out = {}
    for k, v in state_dict.items():
        shape = tuple(v.shape)
        if k.endswith("_count"):
            a = np.asarray(1000.0, dtype=np.float32)
        elif k.endswith("_mean"):
            a = rng.uniform(0.1, 0.6, size=shape).astype(np.float32)
        elif k.endswith("_var"):
            a = rng.uniform(0.005, 0.15, size=shape).astype(np.float32)  # some below the 1e-2 clamp
        elif len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            b = 1.0 / np.sqrt(fan_in)
            a = rng.uniform(-b, b, size=shape).astype(np.float32)
        elif k.endswith("weight"):
            a = rng.uniform(0.5, 1.5, size=shape).astype(np.float32)
        else:
            a = rng.uniform(-0.1, 0.1, size=shape).astype(np.float32)
        out[k] = a
    return out
