"""CPU: the oracle port (oracle/add_oracle.py) against the golden vectors produced by the EXECUTED
reference (tests/golden/make_golden.py).  Same seed, same SyntheticEngine stream -> the port must
reproduce the reference bit for bit (it is the same torch-on-CPU arithmetic in the same order)."""
import os

import numpy as np
import pytest
import torch

from add_gym_b200 import config as b200_config
from oracle import add_oracle, harness

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BUF_KEYS = ["obs", "next_obs", "action", "reward", "done", "a_logp", "tar_val", "adv", "rand_action_mask", "disc_obs",
            "disc_obs_demo", "motion_ids", "motion_times"]


def _run(case, num_envs, motion_file, task_overrides=None):
    cfg = b200_config.default_config(num_envs=num_envs, motion_file=motion_file)
    cfg["task"].update(task_overrides or {})
    agent = harness.make_oracle_agent(num_envs, seed=0, engine_seed=1234, cfg=cfg, fall_prob=0.01)
    g = np.load(os.path.join(GOLD, case + ".npz"))
    return agent, g


LOCAL = {"global_obs": False, "enable_vel_obs": True, "enable_phase_obs": True}


@pytest.mark.parametrize("case,num_envs,motion,overrides", [
    ("walk_n12", 12, None, None),
    ("three_clips_n10", 10, os.path.join(b200_config.ASSET_DIR, "three_clips.yaml"), None),
    ("local_vel_phase_n9", 9, os.path.join(b200_config.ASSET_DIR, "three_clips.yaml"), LOCAL),
    ("seven_clips_n14", 14, os.path.join(b200_config.ASSET_DIR, "seven_clips.yaml"), None),
])
def test_oracle_reproduces_reference_iteration(case, num_envs, motion, overrides):
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    agent, g = _run(case, num_envs, motion, overrides)
    lib = agent.lib
    # step table (slerp/lerp resampling) -- sampled rows, shape, column checksums
    assert list(g["table_shape"]) == list(lib.table.shape)
    np.testing.assert_array_equal(lib.table[torch.from_numpy(g["table_rows"])].numpy(), g["table_sample"])
    np.testing.assert_allclose(lib.table.double().sum(0).numpy(), g["table_colsum"], rtol=1e-12)
    np.testing.assert_array_equal(lib.start_idx.numpy(), g["start_idx"])       # quirk Q2: 30 fps cumsum
    np.testing.assert_array_equal(lib.lengths.numpy(), g["lengths"])
    for k in agent.names:
        np.testing.assert_array_equal(agent.params[k].detach().flatten()[:64].numpy(), g["p0/_" + k[1:]] if False else g["p0/" + k])
    agent.start()
    np.testing.assert_array_equal(agent.curr_obs.numpy(), g["obs0"])
    np.testing.assert_array_equal(agent.motion_ids.numpy(), g["ids0"])
    np.testing.assert_array_equal(agent.offsets.numpy(), g["off0"])
    info = agent.train_iter()
    for k in BUF_KEYS:
        np.testing.assert_array_equal(agent.buf[k].numpy(), g["buf/" + k], err_msg=k)
    for k in agent.names:
        np.testing.assert_array_equal(agent.params[k].detach().flatten()[:64].numpy(), g["p1/" + k], err_msg=k)
        assert abs(agent.params[k].detach().double().sum().item() - float(g["p1sum/" + k])) <= 1e-9 * max(1.0, abs(float(g["p1sum/" + k])))
    np.testing.assert_array_equal(agent.obs_mean.numpy(), g["obs_mean"])
    np.testing.assert_array_equal(agent.obs_std.numpy(), g["obs_std"])
    np.testing.assert_array_equal(agent.diff_mean_abs.numpy(), g["diff_mean_abs"])
    np.testing.assert_array_equal(agent.errors.numpy(), g["sampler_errors"])
    for k in ("loss", "critic_loss", "actor_loss", "clip_frac", "imp_ratio", "disc_loss", "disc_grad_penalty",
              "disc_logit_loss", "disc_neg_logit", "adv_mean", "adv_std", "disc_reward_mean", "disc_reward_std"):
        assert abs(float(info[k]) - float(g["info/" + k])) <= 1e-6 * max(1.0, abs(float(g["info/" + k]))), k


def test_arange_time_formula_matches_torch():
    """The build kernel and the oracle sample the clips at torch.arange(0, len, dt)'s exact fp32 values."""
    for length in (124.16667, 23.3, 14.966667, 4.0, 0.37, 300.0, 77.123, 9.99, 10.0, 10.01):
        L = torch.tensor(length, dtype=torch.float32)
        ref = torch.arange(0, L, 0.01)
        got = add_oracle.arange_times(ref.numel(), 0.01)
        assert torch.equal(ref, got), length


def test_adamw_restatement_matches_torch():
    torch.manual_seed(3)
    p = torch.randn(257, 33)
    q = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([q], 1e-4, weight_decay=0.0)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 6):
        g = torch.randn_like(p) * 10 ** (-step)
        q.grad = g.clone()
        opt.step()
        add_oracle.adamw_step(p, g, m, v, step, 1e-4)
        assert torch.equal(p, q.data), step


def test_td_lambda_properties():
    """Linearity in (r, next_vals) and the lambda cut at done steps."""
    torch.manual_seed(0)
    T, N = 16, 9
    r1, r2, v1, v2 = (torch.randn(T, N).double() for _ in range(4))
    done = (torch.rand(T, N) < 0.2).int() * 3
    f = lambda r, v: add_oracle.td_lambda_return(r, v, done, 0.99, 0.95)
    assert torch.allclose(f(r1 + 2 * r2, v1 + 2 * v2), f(r1, v1) + 2 * f(r2, v2), atol=1e-12)
    ret = f(r1, v1)
    cut = done[:-1] != 0
    assert torch.allclose(ret[:-1][cut], (r1[:-1] + 0.99 * v1[:-1])[cut], atol=1e-12)


def test_oracle_calc_motion_frame_reproduces_reference():
    """MotionLib.calc_motion_frame at arbitrary (clip, time) queries (motion_lib.py:61-88): the oracle's restatement against
    the executed reference on the seven-clip library with two WRAP clips -- bit for bit."""
    g = np.load(os.path.join(GOLD, "motion_frame_queries.npz"))
    cfg = b200_config.default_config(num_envs=2, motion_file=os.path.join(b200_config.ASSET_DIR, "seven_clips.yaml"))
    clips, weights = harness.load_clips(cfg["task"]["motion_file"])
    clips = [(f, fps, int(lm)) for (f, fps, _), lm in zip(clips, g["loop_modes"])]
    from add_gym_b200 import kinematics
    kin = kinematics.KinCharModel()
    kin.load_char_file(cfg["robot"]["urdf_path"])
    lib = add_oracle.OracleMotionLib(clips, weights, kin.dof_axes(), kin.motion_column_of_dof(list(cfg["task"]["motion_joint_order"])),
                                     cfg["engine"]["ctrl_dt"])
    out = lib.calc_motion_frame(torch.from_numpy(g["ids"]), torch.from_numpy(g["times"]))
    for k, v in zip(("root_pos", "root_rot", "root_vel", "root_ang_vel", "joint_rot", "dof_pos", "dof_vel"), out):
        np.testing.assert_array_equal(v.numpy(), g[k], err_msg=k)


def _ref_checkpoint():
    """tests/golden/ref_checkpoint.zip: written by the executed reference's own save() (make_ref_checkpoint.py)"""
    import io
    import zipfile
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    with zipfile.ZipFile(os.path.join(here, "ref_checkpoint.zip")) as z:
        return torch.load(io.BytesIO(z.read("model.pt")), map_location="cpu")


def test_reference_checkpoint_fixture_has_the_layout_the_loader_expects():
    """The checkpoint the reference wrote (base_agent.py:148-155): dict layout, state-dict key names, one optimizer state
    entry per trainable tensor in `named_parameters()` order, the per-tensor patterns of make_ref_checkpoint.py."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_ref_checkpoint import pattern
    ck = _ref_checkpoint()
    assert set(ck.keys()) == {"model", "optimizer", "iter", "sample_count"}
    assert ck["iter"] == 7 and ck["sample_count"] == 7 * 128
    trainable = [k for k in ck["model"] if k.startswith("_model.") and not k.endswith("_logstd_net")]
    assert len(trainable) == 22 and len(ck["optimizer"]["state"]) == 22
    assert trainable[0] == "_model._actor_layers.0.weight" and trainable[-1] == "_model._disc_logits.bias"
    for i, k in enumerate(trainable):
        p, st = ck["model"][k], ck["optimizer"]["state"][i]
        assert st["exp_avg"].shape == p.shape and float(st["step"]) == 40.0
        assert torch.equal(p.flatten(), pattern(p.numel(), 0, i)), k
        assert torch.equal(st["exp_avg"].flatten(), pattern(p.numel(), 1, i)), k
        assert torch.equal(st["exp_avg_sq"].flatten(), pattern(p.numel(), 2, i)), k
    g = ck["optimizer"]["param_groups"][0]
    assert g["lr"] == 1e-4 and tuple(g["betas"]) == (0.9, 0.999) and g["weight_decay"] == 0.0
    assert int(ck["model"]["_obs_norm._count"]) == 128        # one iteration of 4 envs x 32 steps went into the normalizers
