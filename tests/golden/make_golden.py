"""Generate tests/golden/*.npz by EXECUTING the unmodified reference (needs /root/reference).

    python tests/golden/make_golden.py

For each case the real `ADDAgent` (reference add_gym/learning/add/add_agent.py) is constructed on CPU over
a SyntheticEngine with `torch.manual_seed(0)`, all envs are reset, `_init_train()`'s buffer clear is applied
and one `_train_iter()` (32-step rollout + build_train_data + 40 optimizer steps + normalizer update) runs.
Stored: a strided sample of the motion step table, the initial observation, every experience buffer, the
per-iteration diagnostics, post-update parameter digests, normalizer and sampler state.  These are the
pins for oracle/add_oracle.py (tests/test_oracle_golden.py) and, through it, for the CUDA path.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from oracle import ref_harness  # noqa: E402

from oracle.make_assets import SEVEN_CLIPS  # noqa: E402

THREE = [("walk1_subject1_trimmed.motion", 1.0, None), ("run2_subject4_trimmed.motion", 0.5, 700),
         ("fallAndGetUp3_subject1.motion", 0.25, 450)]
CASES = {
    "walk_n12": dict(num_envs=12, motion_file=None, fall_prob=0.01),
    "three_clips_n10": dict(num_envs=10, fall_prob=0.01, motion_file=THREE),
    # non-default observation switches the reference keeps: local frame (heading-relative), velocity and phase features
    "local_vel_phase_n9": dict(num_envs=9, fall_prob=0.01, motion_file=THREE,
                               task_overrides={"global_obs": False, "enable_vel_obs": True, "enable_phase_obs": True}),
    # seven clips (ids 0..6): the Q2 start-index quirk beyond clip 2, short CLAMP clips that end inside the rollout
    "seven_clips_n14": dict(num_envs=14, fall_prob=0.01,
                            motion_file=[(n + ".motion", w, cut) for n, w, cut in SEVEN_CLIPS]),
}
BUF_KEYS = ["obs", "next_obs", "action", "reward", "done", "a_logp", "tar_val", "adv", "rand_action_mask", "disc_obs",
            "disc_obs_demo", "motion_ids", "motion_times"]


def run_case(name, num_envs, motion_file, fall_prob, task_overrides=None):
    agent, cfg = ref_harness.make_reference_agent(num_envs, seed=0, engine_seed=1234, motion_file=motion_file,
                                                  fall_prob=fall_prob, task_overrides=task_overrides)
    out = {}
    ml = agent._add_motion.motion_lib
    S = ml._step_root_pos.shape[0]
    rows = np.arange(0, S, 97)
    tab = torch.cat([ml._step_root_pos, ml._step_root_rot, ml._step_dof_pos, ml._step_root_vel, ml._step_root_ang_vel,
                     ml._step_dof_vel], dim=-1)
    out["table_rows"] = rows
    out["table_sample"] = tab[rows].numpy()
    out["table_shape"] = np.array(tab.shape)
    out["table_colsum"] = tab.double().sum(0).numpy()
    out["start_idx"] = ml._motion_start_idx.numpy()
    out["lengths"] = ml._motion_lengths.numpy()
    names = [k for k, p in agent._model.named_parameters() if p.requires_grad]
    for k, p in agent._model.named_parameters():
        if p.requires_grad:
            out["p0/" + k] = p.detach().flatten()[:64].numpy().copy()
            out["p0sum/" + k] = np.array(p.detach().double().sum().item())
    agent._curr_obs, agent._curr_info = agent._reset_envs()
    agent._init_train_done = True
    agent._exp_buffer.clear()
    agent._train_return_tracker.reset()
    out["obs0"] = agent._curr_obs.numpy().copy()
    out["ids0"] = agent._add_obs._motion_ids.numpy().copy()
    out["off0"] = agent._add_obs._motion_time_offsets.numpy().copy()
    info = agent._train_iter()
    for k in BUF_KEYS:
        out["buf/" + k] = agent._exp_buffer.get_data(k).numpy().copy()
    for k, v in info.items():
        out["info/" + k] = np.array(float(v))
    for k, p in agent._model.named_parameters():
        if p.requires_grad:
            out["p1/" + k] = p.detach().flatten()[:64].numpy().copy()
            out["p1sum/" + k] = np.array(p.detach().double().sum().item())
    out["obs_mean"] = agent._obs_norm._mean.numpy().copy()
    out["obs_std"] = agent._obs_norm._std.numpy().copy()
    out["diff_mean_abs"] = agent._disc_obs_norm._mean_abs.numpy().copy()
    out["sampler_errors"] = agent._add_motion.sampler.errors.numpy().copy()
    out["param_names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "written;", {k: float(v) for k, v in info.items() if k in ("loss", "mean_return", "num_eps")})


def run_motion_frame_queries(name="motion_frame_queries"):
    """MotionLib.calc_motion_frame of the executed reference at arbitrary (clip, time) queries on the seven-clip library
    with two clips switched to WRAP: the runtime interpolation API (root lerp, root / joint slerp, twist angle, loop offset)."""
    ref_harness._install_stubs()
    import add_gym.anim.motion as ref_motion
    orig = ref_motion.load_motion

    def load_with_wrap(file):
        m = orig(file)
        if "run2" in file or "jumps1" in file:
            m.loop_mode = ref_motion.LoopMode.WRAP
        return m
    agent, cfg = ref_harness.make_reference_agent(2, seed=0, engine_seed=1234,
                                                  motion_file=[(n + ".motion", w, cut) for n, w, cut in SEVEN_CLIPS])
    import add_gym.anim.motion_lib as ref_ml
    ref_ml.motion.load_motion = load_with_wrap
    try:
        ml = ref_ml.MotionLib(motion_file=cfg["task"]["motion_file"], motion_order=list(cfg["task"]["motion_joint_order"]),
                              kin_char_model=agent._env.robot._kin_char_model, dt=0.01, device="cpu")
    finally:
        ref_ml.motion.load_motion = orig
    g = torch.Generator().manual_seed(21)
    n = 4000
    ids = torch.randint(0, 7, (n,), generator=g)
    times = torch.rand(n, generator=g) * 30.0 - 2.0          # before the start, inside, beyond the end (CLAMP) / several loops (WRAP)
    times[:14] = torch.tensor([0.0, 1.0 / 30.0, 13.3, 13.3 + 1e-4, 1e-6, -1.0, 100.0] * 2)
    ids[:14] = torch.tensor([0] * 7 + [2] * 7)
    out = ml.calc_motion_frame(ids, times)
    res = {"ids": ids.numpy(), "times": times.numpy(), "loop_modes": ml._motion_loop_modes.numpy()}
    for k, v in zip(("root_pos", "root_rot", "root_vel", "root_ang_vel", "joint_rot", "dof_pos", "dof_vel"), out):
        res[k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **res)
    print(name, "written;", {k: v.shape for k, v in res.items()})


if __name__ == "__main__":
    only = sys.argv[1:]
    if not only or "motion_frame_queries" in only:
        run_motion_frame_queries()
    for name, kw in CASES.items():
        if not only or name in only:
            run_case(name, **kw)
