"""CPU ORACLE (test infrastructure, NOT product code) for the HBM-bound pieces of the PointNav-VO
hot path: depth discretisation, egocentric top-down projection, the PPO GAE/return scan and the
SE(2) goal update.  Plain numpy with every fp32 rounding made explicit so that the result does not
depend on BLAS / torch / cv2 builds.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product path (pointnav_vo_b200.*) never does.

Parity status: PINNED.  The reference has no tests or golden vectors of its own,
so every function here is checked bit-for-bit against the unmodified reference code executed in the
build container (tests/golden/make_golden.py -> tests/golden/*.npz, tests/test_oracle_golden.py).

Reference (all paths relative to /root/reference/pointnav_vo):
  discretize_depth      rl/common/base_trainer_with_vo.py:135-167 (torch fp32 semantics)
  TopDownOracle         utils/geometry_utils.py:491-721 (NormalizedDepth2TopDownViewHabitatTorch)
  gae_returns           rl/common/rollout_storage.py:102-120
  compute_goal_pos      utils/geometry_utils.py:115-144
"""
import math

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------------
# a7: depth discretisation
# --------------------------------------------------------------------------------------------
def discretize_end_vals(n_channels):
    """base_trainer_with_vo.py:107-117: python-double edges i*1.0/n, last edge 1.0."""
    return [i * 1.0 / n_channels for i in range(n_channels)] + [1.0]


def discretize_depth_index(depth, n_channels=10):
    """Bin index per pixel (uint8).  depth: fp32 array in [0, 1], any shape.

    base_trainer_with_vo.py:143-154: bin i holds d >= e_i and d < e_{i+1}; the last bin is closed on
    the right.  torch compares an fp32 tensor with a python double by casting the double to fp32, so the
    thresholds are float32(i/n) -- NOT floor(d*n).
    """
    depth = np.asarray(depth, dtype=F32)
    if depth.size:
        assert depth.max() <= 1.0 and depth.min() >= 0.0  # :136-137
    edges = np.array(discretize_end_vals(n_channels), dtype=F32)
    idx = np.full(depth.shape, 255, dtype=np.uint8)
    for i in range(n_channels):
        if i == n_channels - 1:
            m = (depth >= edges[i]) & (depth <= edges[i + 1])
        else:
            m = (depth >= edges[i]) & (depth < edges[i + 1])
        idx[m] = i
    assert not (idx == 255).any()  # :162-163 one-hot completeness
    return idx


def discretize_depth_onehot(depth, n_channels=10):
    """fp32 one-hot [..., n_channels] exactly as the reference returns it (:139-159)."""
    idx = discretize_depth_index(depth, n_channels)
    out = np.zeros(idx.shape + (n_channels,), dtype=F32)
    np.put_along_axis(out, idx[..., None].astype(np.int64), 1.0, axis=-1)
    return out


# --------------------------------------------------------------------------------------------
# a8: top-down projection (Torch fp32 variant)
# --------------------------------------------------------------------------------------------
class TopDownOracle:
    """Restatement of NormalizedDepth2TopDownViewHabitatTorch (geometry_utils.py:491-721).

    hfov_rad is used verbatim (the reference passes degrees)."""

    def __init__(self, min_depth=0.1, max_depth=10.0, vis_size_h=192, vis_size_w=341, hfov_rad=70,
                 ksize=3, rows_around_center=50, flag_center_crop=True):
        assert ksize == 3
        self.min_depth, self.max_depth = min_depth, max_depth
        self.H, self.W = vis_size_h, vis_size_w
        self.rows_around_center = rows_around_center
        self.flag_center_crop = flag_center_crop
        # :565-572  K = FloatTensor([[f,0,u0],[0,f,v0],[0,0,1]]); f is a python double rounded once.
        f = F32((vis_size_w / 2) / math.tan(hfov_rad / 2))
        u0 = F32(vis_size_w / 2)
        # torch.inverse(K) (:577,:648) for this upper-triangular K gives exactly
        #   Kinv[0,0] = fl(1/f), Kinv[0,1] = 0, Kinv[0,2] = -fl(u0/f)  (checked vs torch in
        #   make_golden.py; both roundings are what LAPACK getri produces for this matrix)
        self.k00 = F32(1.0) / f
        self.k02 = -(u0 / f)
        # :573-583 _get_x_range(max_depth): (Kinv @ [W-0.5, 0, 1]) * depth
        xr = F32(self.k00 * F32(vis_size_w - 0.5)) + self.k02
        self.max_x = F32(xr * F32(max_depth))
        self.min_x = -self.max_x
        self.x_range = F32(self.max_x - self.min_x)
        eps = 0.01
        self.x_den = F32(self.x_range * F32(1 + eps))  # :679
        self.z_den = F32((max_depth - min_depth) * (1 + eps))  # :680-682 python-double product, one rounding
        self.depth_scale = F32(max_depth - min_depth)  # :559
        self.depth_off = F32(min_depth)
        # ray[u] = (Kinv @ [u+.5, v+.5, 1])[0]; sgemm result = fl(fl(k00*(u+.5)) + k02), independent of v
        u = np.arange(vis_size_w, dtype=F32) + F32(0.5)
        self.ray = (self.k00 * u).astype(F32) + self.k02

    @staticmethod
    def bbox(depth):
        """:585-606 first/last row and column whose fp32 sum is > 0 (depth >= 0 so: any element > 0)."""
        nz = depth > 0
        rows = np.flatnonzero(nz.any(axis=1))
        cols = np.flatnonzero(nz.any(axis=0))
        if rows.size == 0:
            return None
        return int(rows[0]), int(rows[-1]), int(cols[0]), int(cols[-1])

    @staticmethod
    def blur3(x):
        """cv2.GaussianBlur(x, (3,3), 0, 0, BORDER_ISOLATED|BORDER_CONSTANT(0)) on fp32 (:529-535).
        Separable [.25,.5,.25]; cv2's fp32 row/column filters evaluate (a + c)*.25 + b*.5 without FMA,
        rows (horizontal) first, zero padding at the crop edge."""
        x = x.astype(F32)
        p = np.pad(x, ((0, 0), (1, 1)))
        h = ((p[:, :-2] + p[:, 2:]).astype(F32) * F32(0.25)).astype(F32) + (p[:, 1:-1] * F32(0.5)).astype(F32)
        h = h.astype(F32)
        p = np.pad(h, ((1, 1), (0, 0)))
        v = ((p[:-2, :] + p[2:, :]).astype(F32) * F32(0.25)).astype(F32) + (p[1:-1, :] * F32(0.5)).astype(F32)
        return v.astype(F32)

    def center_rows(self, h):
        """:608-620 rows of the crop that are projected."""
        if self.flag_center_crop:
            c = int(math.ceil(h / 2))
            return max(0, c - self.rows_around_center), min(h, c + self.rows_around_center)
        return 0, min(self.rows_around_center * 2, h)

    def pixel_coords(self, d, col0):
        """d: blurred crop rows [n, w] fp32; col0: first non-zero column.  Returns (row, col) int64."""
        z = (d * self.depth_scale).astype(F32) + self.depth_off  # :558-560
        z = z.astype(F32)
        ray = self.ray[col0:col0 + d.shape[1]][None, :]
        x = (ray * z).astype(F32)  # :656 coords_3d *= true_depth
        nx = ((x - self.min_x).astype(F32) / self.x_den).astype(F32)  # :679
        nz = ((z - self.depth_off).astype(F32) / self.z_den).astype(F32)  # :680
        row = F32(self.H) - np.ceil((F32(self.H) * nz).astype(F32))  # :689-691
        col = np.floor((F32(self.W) * nx).astype(F32))  # :692
        return row.astype(np.int64), col.astype(np.int64)

    def count_map(self, depth):
        """int32 [H, W] point counts (before normalisation); the bit-exact index map."""
        depth = np.asarray(depth, dtype=F32).reshape(self.H, self.W)
        cnt = np.zeros((self.H, self.W), dtype=np.int32)
        bb = self.bbox(depth)
        if bb is None:  # :519-525
            return cnt
        r0, r1, c0, c1 = bb
        blurred = self.blur3(depth[r0:r1 + 1, c0:c1 + 1])
        a, b = self.center_rows(blurred.shape[0])
        row, col = self.pixel_coords(blurred[a:b], c0)
        ok = (row >= 0) & (row < self.H) & (col >= 0) & (col < self.W)  # :706-711
        np.add.at(cnt, (row[ok], col[ok]), 1)
        return cnt

    def gen_top_down_view(self, depth):
        """[H, W, 1] fp32 = count / max(count) (:541-554)."""
        cnt = self.count_map(depth).astype(F32)
        m = cnt.max()
        if m == 0:
            return np.zeros((self.H, self.W, 1), dtype=F32)
        out = (cnt / F32(m)).astype(F32)
        out[out > 1.0] = 1.0
        return out[..., None]


# --------------------------------------------------------------------------------------------
# a13: GAE / discounted returns
# --------------------------------------------------------------------------------------------
def gae_returns(rewards, value_preds, masks, next_value, use_gae=True, gamma=0.99, tau=0.95):
    """rollout_storage.py:102-120.  rewards [T,N,1]; value_preds, masks [T+1,N,1]; next_value [N,1].
    Returns returns[T+1,N,1] fp32 (row T is left 0 in GAE mode, = next_value otherwise), and the
    value_preds with row T overwritten (as the reference does in place).

    torch multiplies fp32 tensors by python doubles in fp32 (scalar cast to fp32); gamma * tau is a
    python-double product rounded once."""
    rewards = np.asarray(rewards, dtype=F32)
    value_preds = np.array(value_preds, dtype=F32, copy=True)
    masks = np.asarray(masks, dtype=F32)
    T = rewards.shape[0]
    returns = np.zeros_like(value_preds)
    g, gt = F32(gamma), F32(gamma * tau)
    if use_gae:
        value_preds[T] = next_value
        gae = np.zeros_like(value_preds[0])
        for t in reversed(range(T)):
            # delta = r + gamma * V[t+1] * m[t+1] - V[t]   (left-to-right fp32 evaluation)
            delta = (rewards[t] + ((g * value_preds[t + 1]).astype(F32) * masks[t + 1]).astype(F32)).astype(F32)
            delta = (delta - value_preds[t]).astype(F32)
            # gae = delta + gamma*tau * m[t+1] * gae
            gae = (delta + ((gt * masks[t + 1]).astype(F32) * gae).astype(F32)).astype(F32)
            returns[t] = (gae + value_preds[t]).astype(F32)
    else:
        returns[T] = next_value
        for t in reversed(range(T)):
            returns[t] = (((returns[t + 1] * g).astype(F32) * masks[t + 1]).astype(F32) + rewards[t]).astype(F32)
    return returns, value_preds


# --------------------------------------------------------------------------------------------
# a10: goal update (fp64, CPU in the reference)
# --------------------------------------------------------------------------------------------
def compute_goal_pos(prev_goal_pos, local_delta_state):
    """geometry_utils.py:115-144 with the habitat helpers written out:
    q = angle-axis(dyaw, +y); g' = q^-1 (g - [dx,0,dz]) q; polar: rho = |(-g'_z, g'_x)|,
    phi = atan2(g'_x, -g'_z); returns cartesian g' (fp64) and polar [rho, -phi] (fp32)."""
    dx, dz, dyaw = (float(v) for v in local_delta_state)
    v = np.asarray(prev_goal_pos, dtype=np.float64) - np.array([dx, 0.0, dz])
    c, s = math.cos(dyaw), math.sin(dyaw)
    # rotation by -dyaw about +y:  [x', z'] = [c x - s z, s x + c z]
    cur = np.array([c * v[0] - s * v[2], v[1], s * v[0] + c * v[2]])
    rho = math.hypot(-cur[2], cur[0])
    phi = math.atan2(cur[0], -cur[2])
    return {"cartesian": cur, "polar": np.array([rho, -phi], dtype=np.float32)}
