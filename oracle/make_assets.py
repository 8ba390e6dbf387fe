"""Derive the small data files the repo ships from the reference's read-only assets.

TEST/BENCH INFRASTRUCTURE.  Runs only in the build container (needs /root/reference):
  * assets/g1_29_kinematics.json  <- assets/g1_description/g1_29.xml  (BFS body list, hinge axes, limits)
  * assets/*.npy                  <- assets/motions/*.motion           ([F,36] rows parsed exactly as
                                     reference add_gym/anim/motion.py:26-31 does: float(text) -> float64)
The `.npy` clips keep float64 so that the later float32 rounding is the reference's own
(motion_lib.py:108-110).  Two extra clips are truncated to keep the repository small; they are only
used by the multi-clip parity tests.
"""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from add_gym_b200.kinematics import parse_mjcf_bodies  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(REPO, "add_gym_b200", "assets")


def parse_motion(path, max_frames=None):
    rows = []
    with open(path, "r") as f:
        for line in f:
            rows.append([float(v) for v in line.strip().split(",")])
            if max_frames is not None and len(rows) >= max_frames:
                break
    return np.array(rows)


def main():
    os.makedirs(OUT, exist_ok=True)
    bodies = parse_mjcf_bodies(os.path.join(REF, "assets/g1_description/g1_29.xml"))
    with open(os.path.join(OUT, "g1_29_kinematics.json"), "w") as f:
        json.dump({"source": "assets/g1_description/g1_29.xml", "bodies": bodies}, f, indent=1)
    clips = [
        ("walk1_subject1_trimmed", None),
        ("run2_subject4_trimmed", 700),
        ("fallAndGetUp3_subject1", 450),
    ]
    for name, mx in clips:
        fr = parse_motion(os.path.join(REF, "assets/motions", name + ".motion"), mx)
        np.save(os.path.join(OUT, name + ".npy"), fr)
        print(name, fr.shape, fr.dtype)
    with open(os.path.join(OUT, "three_clips.yaml"), "w") as f:
        f.write("motions:\n")
        for (name, _), w in zip(clips, (1.0, 0.5, 0.25)):
            f.write("  - file: \"{}.npy\"\n    weight: {}\n".format(name, w))


if __name__ == "__main__":
    main()
