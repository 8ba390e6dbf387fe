"""Builders shared by the tests and bench.py's CPU-baseline leg.  TEST / BASELINE INFRASTRUCTURE.

`make_oracle_agent` assembles the CPU oracle (oracle/add_oracle.py) over a CPU SyntheticEngine using only
files that ship in this repository, so it also runs on the GPU box (where /root/reference is absent).
"""
import os
import sys

import numpy as np
import torch
import yaml

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from add_gym_b200 import config as b200_config  # noqa: E402
from add_gym_b200 import kinematics, motion_io  # noqa: E402
from add_gym_b200.env import ImitationEnvironment  # noqa: E402
from oracle import add_oracle  # noqa: E402


def load_clips(motion_file):
    files, weights = motion_io.fetch_motion_files(motion_file)
    clips = []
    for f in files:
        m = motion_io.load_motion(f)
        clips.append((m.frames, m.fps, m.loop_mode.value))
    return clips, weights


def make_oracle_lib(cfg, fix_start_idx=False, jrot_override=None):
    kin = kinematics.KinCharModel()
    kin.load_char_file(cfg["robot"]["urdf_path"])
    clips, weights = load_clips(cfg["task"]["motion_file"])
    return add_oracle.OracleMotionLib(clips, weights, kin.dof_axes(),
                                      kin.motion_column_of_dof(list(cfg["task"]["motion_joint_order"])),
                                      cfg["engine"]["ctrl_dt"], fix_start_idx=fix_start_idx, jrot_override=jrot_override)


def make_cpu_env(cfg, engine_seed=1234, fall_prob=0.002, device="cpu"):
    c = {**cfg, "engine": {**cfg["engine"], "seed": engine_seed, "noise_device": "cpu" if device == "cpu" else "device",
                           "fall_prob": fall_prob,
                           "_target_": cfg["engine"].get("_target_", "add_gym_b200.engine.SyntheticEngine")}}
    return ImitationEnvironment(c, device)


def make_oracle_agent(num_envs, seed=0, engine_seed=1234, cfg=None, rng=None, mimic_reference_rng=True,
                      fall_prob=0.002, lib=None, device="cpu"):
    """device != "cpu": the GPU-eager comparator of bench.py -- the same port with every tensor on the device (the motion
    table is built on the CPU and moved); the caller runs it inside `with torch.device(device):`."""
    cfg = cfg or b200_config.default_config(num_envs=num_envs)
    cfg["engine"]["num_envs"] = num_envs
    lib = lib or make_oracle_lib(cfg, cfg["task"].get("fix_start_idx", False))
    if device != "cpu":
        lib = lib.to(device)
        with torch.device(device):
            env = make_cpu_env(cfg, engine_seed, fall_prob, device=device)
            torch.manual_seed(seed)
            return add_oracle.OracleAgent(cfg, env, lib, rng=rng, mimic_reference_rng=mimic_reference_rng)
    env = make_cpu_env(cfg, engine_seed, fall_prob)
    torch.manual_seed(seed)
    agent = add_oracle.OracleAgent(cfg, env, lib, rng=rng, mimic_reference_rng=mimic_reference_rng)
    return agent
