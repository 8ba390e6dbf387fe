"""Motion clip files (host side, load time only).

Same data model as the reference (add_gym/anim/motion.py:7-75): a clip is ``frames[F,36]`` float64
(root xyz, root quaternion **xyzw**, 29 hinge angles in file order) + ``fps`` (30) + ``loop_mode``
(CLAMP).  Accepted containers: the reference's CSV ``.motion`` text and its ``.pkl`` dict, plus a
plain ``.npy`` of the frame matrix (binary clips avoid the per-value ``float()`` parse).
Unlike the reference, loading a ``.motion`` does NOT write a ``.pkl`` next to it (motion.py:40-42
does; the asset tree may be read-only).
"""
import enum
import pickle

import numpy as np


class LoopMode(enum.Enum):
    CLAMP = 0
    WRAP = 1


class Motion:
    def __init__(self, loop_mode, fps, frames):
        self.loop_mode = loop_mode
        self.fps = fps
        self.frames = frames

    def get_length(self):
        return float(self.frames.shape[0] - 1) / self.fps


def load_motion(file, loop_mode=LoopMode.CLAMP, fps=30):
    if file.endswith(".motion"):
        rows = []
        with open(file, "r") as f:
            for line in f:
                rows.append([float(v) for v in line.strip().split(",")])
        return Motion(loop_mode, fps, np.array(rows))
    if file.endswith(".npy"):
        return Motion(loop_mode, fps, np.load(file))
    with open(file, "rb") as f:
        d = pickle.load(f)
    return Motion(LoopMode(d["loop_mode"]), d["fps"], d["frames"])
