"""Constants of the VO path (values follow pointnav_vo/vo/common/common_vars.py:1-57; they are the
action ids / channel counts the reference's engines, trainers and checkpoints are keyed on)."""
import math

N_ACTS = 4
UNIFIED, STOP, MOVE_FORWARD, TURN_LEFT, TURN_RIGHT = -1, 0, 1, 2, 3
ACT_IDX2NAME = {UNIFIED: "unified", MOVE_FORWARD: "forward", TURN_LEFT: "left", TURN_RIGHT: "right"}
ACT_NAME2IDX = {"forward": MOVE_FORWARD, "left": TURN_LEFT, "right": TURN_RIGHT, "all": -1}
CUR_REL_TO_PREV, PREV_REL_TO_CUR = 0, 1
NO_NOISE_DELTAS = {MOVE_FORWARD: [0.0, -0.25, 0.0], TURN_LEFT: [0.0, 0.0, math.radians(10)],
                   TURN_RIGHT: [0.0, 0.0, -math.radians(10)]}
DEFAULT_LOSS_WEIGHTS = {"dx": 1.0, "dz": 1.0, "dyaw": 1.0}
DEFAULT_DELTA_TYPES = ["dx", "dz", "dyaw"]
EMBED_DIM = 32
RGB_PAIR_CHANNEL = 6
DEPTH_PAIR_CHANNEL = 2
TOP_DOWN_VIEW_PAIR_CHANNEL = 2
DEFAULT_DELTA_STATE_SIZE = 4
