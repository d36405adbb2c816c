"""Geometric-invariance augmentation without a second copy of the frames.

The reference's dataset (vo/dataset/regression_geo_invariance_iter_dataset.py:290-420) appends, for every turn sample
it wants inverted, a SECOND sample with prev / cur swapped on the channel axis, the opposite turn action and the pose of
prev relative to cur as regression target; both copies then cross PCIe and both get their own top-down projection.
Here the host ships every frame pair once, plus a `pair_map` with one int32 per network row

    pair_map[row] = 2 * source_pair + swap

and the input kernels (csrc/raw_input.cu) read the source pair and swap prev / cur on the fly.  `make_pair_map`
reproduces the dataset's selection rules; `inverse_delta_states` gives the targets of the swapped rows.
"""
import numpy as np

from ..common.common_vars import CUR_REL_TO_PREV, MOVE_FORWARD, PREV_REL_TO_CUR, TURN_LEFT, TURN_RIGHT


def make_pair_map(actions, act_type=-1, geo_invariance_types=("inverse_joint_train",)):
    """actions: per source pair action ids.  Returns dict(pair_map int32 [R], actions int64 [R], data_types int64 [R],
    src int64 [R]) in the dataset's row order: [pair i, (swapped pair i)?, pair i+1, ...]
    (regression_geo_invariance_iter_dataset.py:290-366)."""
    actions = np.asarray(actions).reshape(-1)
    inv_joint = "inverse_joint_train" in geo_invariance_types
    inv_aug = "inverse_data_augment_only" in geo_invariance_types
    rows, acts, types, src = [], [], [], []
    for i, a in enumerate(actions):
        a = int(a)
        if act_type == -1 or a == act_type or inv_joint:
            rows.append(2 * i)
            acts.append(a)
            types.append(CUR_REL_TO_PREV)
            src.append(i)
        flag1 = act_type != -1 and inv_aug and a != MOVE_FORWARD and a != act_type
        flag2 = act_type != -1 and a != MOVE_FORWARD and inv_joint
        if flag1 or flag2:
            rows.append(2 * i + 1)
            acts.append(TURN_LEFT if a == TURN_RIGHT else TURN_RIGHT)
            types.append(PREV_REL_TO_CUR)
            src.append(i)
    return {"pair_map": np.asarray(rows, dtype=np.int32), "actions": np.asarray(acts, dtype=np.int64),
            "data_types": np.asarray(types, dtype=np.int64), "src": np.asarray(src, dtype=np.int64)}


def inverse_delta_states(deltas):
    """[N, 3] (dx, dz, dyaw) of cur relative to prev -> the same for prev relative to cur, for planar motion (rotation
    about the up axis only): dyaw' = -dyaw, pos' = -R(dyaw') pos with the left-handed rotation the reference's inversion
    loss uses (vo_cnn_regression_geo_invariance_engine.py:398-424), so that loss(delta, inverse_delta_states(delta)) = 0.
    The dataset derives the same quantity from the global poses (agent_state_target2ref(cur_state, prev_state),
    regression_geo_invariance_iter_dataset.py:389-413)."""
    d = np.asarray(deltas, dtype=np.float64)
    yaw = -d[:, 2]
    c, s = np.cos(yaw), np.sin(yaw)
    dx = -(c * d[:, 0] + s * d[:, 1])
    dz = -(-s * d[:, 0] + c * d[:, 1])
    return np.stack([dx, dz, yaw], axis=1).astype(np.asarray(deltas).dtype if np.asarray(deltas).dtype.kind == "f" else np.float32)


def expand_targets(deltas, pm):
    """Regression targets of every network row of `pm = make_pair_map(...)`: the source delta for plain rows, its
    inverse for swapped rows."""
    deltas = np.asarray(deltas)
    out = deltas[pm["src"]].copy()
    sw = (pm["pair_map"] & 1) == 1
    if sw.any():
        out[sw] = inverse_delta_states(deltas[pm["src"][sw]])
    return out
