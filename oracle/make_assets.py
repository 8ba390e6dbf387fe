"""Derive the small data files the repo ships from the reference's read-only assets.

TEST/BENCH INFRASTRUCTURE.  Runs only in the build container (needs /root/reference):
  * assets/g1_29_kinematics.json  <- assets/g1_description/g1_29.xml  (BFS body list, hinge axes, limits)
  * assets/*.npy                  <- assets/motions/*.motion           ([F,36] rows parsed exactly as
                                     reference add_gym/anim/motion.py:26-31 does: float(text) -> float64)
  * assets/motions_all.addkc      <- ALL 42 assets/motions/*.motion, lossless (add_gym_b200/motion_io.py:
                                     save_clip_pack; BASELINE configs[2], "G1 multi-clip motion library (all
                                     assets/motions)"), verified bit for bit against the parsed text here
  * assets/seven_clips.yaml       a 7-clip library of truncated pack clips (golden case seven_clips_n14)
The `.npy` clips keep float64 so that the later float32 rounding is the reference's own
(motion_lib.py:108-110).  Two extra clips are truncated to keep the repository small; they are only
used by the multi-clip parity tests.
"""
import glob
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from add_gym_b200 import motion_io  # noqa: E402
from add_gym_b200.kinematics import parse_mjcf_bodies  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(REPO, "add_gym_b200", "assets")
# (clip, weight, frames kept): clip ids 0..6 pin the Q2 start-index quirk well beyond clip 2, with CLAMP clips that end
# inside the rollout; the same list drives tests/golden/make_golden.py
SEVEN_CLIPS = [("dance1_subject1", 1.0, 400), ("walk1_subject1_trimmed", 1.0, 600), ("run2_subject4_trimmed", 0.5, 700),
               ("fallAndGetUp3_subject1", 0.25, 450), ("jumps1_subject1", 0.75, 500), ("fight1_subject2", 0.5, 350),
               ("sprint1_subject2", 1.0, 300)]


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
    # ---- the whole library as one lossless pack
    every = {}
    for path in sorted(glob.glob(os.path.join(REF, "assets/motions", "*.motion"))):
        every[os.path.splitext(os.path.basename(path))[0]] = parse_motion(path)
    pack = os.path.join(OUT, "motions_all.addkc")
    motion_io.save_clip_pack(pack, every)
    motion_io._pack_cache.clear()
    back = motion_io.load_clip_pack(pack)
    assert sorted(back) == sorted(every)
    for k, v in every.items():
        assert np.array_equal(v.view(np.int64), back[k].view(np.int64)), k          # bit for bit, signed zeros included
    print("motions_all.addkc: %d clips, %d frames, %.1f MB" % (len(every), sum(v.shape[0] for v in every.values()),
                                                             os.path.getsize(pack) / 1e6))
    with open(os.path.join(OUT, "seven_clips.yaml"), "w") as f:
        f.write("motions:\n")
        for name, w, cut in SEVEN_CLIPS:
            f.write("  - file: \"motions_all.addkc#{}@{}\"\n    weight: {}\n".format(name, cut, w))


if __name__ == "__main__":
    main()
