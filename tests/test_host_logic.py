"""Host-side logic of the drop-in classes that needs no GPU: the kinematic digest (BFS dof order), the
default configuration (SURVEY section 8 constants) and -- where /root/reference exists -- both against the reference's own
MJCF / YAML files."""
import os

import numpy as np
import pytest

from add_gym_b200 import config as b200_config
from add_gym_b200.kinematics import JOINT_HINGE, KinCharModel, parse_mjcf_bodies

REF = "/root/reference/add_gym"

# SURVEY Q7 (kin_char_model.py:116-161, motion_lib.py:102-111): breadth-first over the MJCF bodies
BFS_DOF_ORDER = [
    "left_hip_pitch", "right_hip_pitch", "waist_yaw", "left_hip_roll", "right_hip_roll", "waist_roll",
    "left_hip_yaw", "right_hip_yaw", "waist_pitch", "left_knee", "right_knee", "left_shoulder_pitch",
    "right_shoulder_pitch", "left_ankle_pitch", "right_ankle_pitch", "left_shoulder_roll", "right_shoulder_roll",
    "left_ankle_roll", "right_ankle_roll", "left_shoulder_yaw", "right_shoulder_yaw", "left_elbow", "right_elbow",
    "left_wrist_roll", "right_wrist_roll", "left_wrist_pitch", "right_wrist_pitch", "left_wrist_yaw", "right_wrist_yaw",
]


def _model():
    m = KinCharModel()
    m.load_char_file(b200_config.default_config()["robot"]["urdf_path"])
    return m


def test_dof_order_is_breadth_first_over_the_bodies():
    m = _model()
    assert m.get_dof_size() == 29
    hinge = [j for j in m._joints[1:] if j.joint_type == JOINT_HINGE]
    assert [j.name for j in hinge] == [n + "_joint" for n in BFS_DOF_ORDER]
    assert [j.dof_idx for j in hinge] == list(range(29)), "dof indices follow body order"
    # parents come before children (what a breadth-first enumeration guarantees) and body 0 is the floating root
    assert m._parent_indices[0] == -1 and all(0 <= p < i for i, p in enumerate(m._parent_indices) if i > 0)


def test_motion_column_permutation_and_axes():
    m = _model()
    order = b200_config.MOTION_JOINT_ORDER
    col = m.motion_column_of_dof(order)
    assert sorted(col.tolist()) == list(range(29)), "a permutation of the 29 file columns"
    assert [order[c] for c in col] == [n + "_joint" for n in BFS_DOF_ORDER]
    ax = m.dof_axes()
    assert ax.shape == (29, 3) and ax.dtype == np.float32
    assert np.allclose(np.linalg.norm(ax, axis=1), 1.0), "unit hinge axes"
    assert np.all(np.sum(ax != 0, axis=1) == 1), "G1 hinges are axis-aligned"
    for name, want in (("left_hip_pitch", 1), ("waist_yaw", 2), ("left_hip_roll", 0), ("left_knee", 1)):
        assert int(np.argmax(np.abs(ax[BFS_DOF_ORDER.index(name)]))) == want, name
    lim = m.dof_limits()
    assert lim.shape == (29, 2) and np.all(lim[:, 0] < lim[:, 1])


def test_default_config_restates_the_reference_constants():
    cfg = b200_config.default_config(num_envs=64)
    a, t, e = cfg["agent"], cfg["task"], cfg["engine"]
    assert e["num_envs"] == 64 and e["ctrl_dt"] == 0.01
    assert (a["steps_per_iter"], a["update_epochs"], a["batch_size"]) == (32, 5, 4)
    assert (a["discount"], a["td_lambda"], a["ppo_clip_ratio"], a["norm_adv_clip"]) == (0.99, 0.95, 0.2, 4.0)
    assert (a["task_reward_weight"], a["disc_reward_weight"], a["disc_reward_scale"]) == (0.0, 1.0, 2)
    assert (a["disc_grad_penalty"], a["disc_logit_reg"], a["disc_weight_decay"], a["disc_loss_weight"]) == (20, 0.01, 0.0001, 0.5)
    assert a["optimizer"] == {"type": "Adam", "learning_rate": 1e-4} and "grad_clip" not in a["optimizer"]   # Q4
    assert a["model"]["mlp_precision"] == "f16x3" and a["model"]["action_std"] == 0.05
    assert t["tar_obs_steps"] == [1, 2, 3, 4, 5, 6] and t["num_disc_obs_steps"] == 3
    assert t["global_obs"] and t["root_height_obs"] and t["enable_tar_obs"] and not t["enable_vel_obs"] and not t["enable_phase_obs"]
    assert (t["reward_pose_w"], t["reward_vel_w"], t["reward_root_pose_w"], t["reward_root_vel_w"]) == (0.5, 0.1, 0.15, 0.1)
    assert (t["reward_pose_scale"], t["reward_vel_scale"], t["reward_root_pose_scale"], t["reward_root_vel_scale"]) == (0.25, 0.01, 5.0, 1.0)
    assert t["max_episode_length"] == 20 and t["pose_termination"] and t["pose_termination_dist"] == 1.0
    assert len(t["motion_joint_order"]) == 29 and os.path.exists(t["motion_file"])
    # a fresh deep copy per call: editing one must not leak into the next
    cfg["task"]["tar_obs_steps"].append(7)
    assert b200_config.default_config()["task"]["tar_obs_steps"] == [1, 2, 3, 4, 5, 6]
    assert b200_config.default_config(mlp_precision="bf16")["agent"]["model"]["mlp_precision"] == "bf16"


@pytest.mark.reference
def test_kinematic_digest_matches_the_reference_mjcf():
    bodies = parse_mjcf_bodies(os.path.join(os.path.dirname(REF), "assets", "g1_description", "g1_29.xml"))
    ref = KinCharModel()
    ref._init_from_bodies(bodies)
    m = _model()
    assert m.get_body_names() == ref.get_body_names() and m._parent_indices == ref._parent_indices
    assert m.get_joint_order() == ref.get_joint_order()
    assert np.array_equal(m.dof_axes(), ref.dof_axes()) and np.array_equal(m.dof_limits(), ref.dof_limits())


@pytest.mark.reference
def test_default_config_matches_the_reference_yaml():
    import yaml
    ref = {}
    for group, name in (("agent", "add_g1"), ("task", "pose"), ("robot", "g1")):
        with open(os.path.join(REF, "configs", group, name + ".yaml")) as f:
            ref[group] = yaml.safe_load(f)
    cfg = b200_config.default_config()
    skip = {"motion_file", "urdf_path", "mlp_precision"}           # paths / B200-only keys
    def walk(ours, theirs, path):
        for k, v in theirs.items():
            if k in skip or (path == "robot" and k not in ours):   # robot.yaml also holds keys the path never reads
                continue
            assert k in ours, "missing key %s.%s" % (path, k)
            if isinstance(v, dict):
                walk(ours[k], v, path + "." + k)
            else:
                if isinstance(v, str) and not isinstance(ours[k], str):
                    v = float(v)                                  # PyYAML reads "1e-4" (no dot) as a string
                assert ours[k] == v, "%s.%s: %r != reference %r" % (path, k, ours[k], v)
    for group in ref:
        walk(cfg[group], ref[group], group)


# ---- reference hooks of the agent whose work is fused into kernels -----------------------------------
class _StubBuffer:
    def __init__(self, T, N):
        import torch
        shapes = {"obs": (264,), "action": (29,), "a_logp": (), "rand_action_mask": (), "next_obs": (264,), "reward": (),
                  "done": (), "disc_obs": (114,), "disc_obs_demo": (114,)}
        self.data = {k: torch.arange(T * N * max(1, int(np.prod(s))), dtype=torch.float32).reshape((T, N) + s)
                     for k, s in shapes.items()}

    def get_data(self, name):
        return self.data[name]


def test_agent_record_hooks_are_called_only_when_overridden_and_compute_loss_refuses():
    import torch
    from add_gym_b200 import _lib
    from add_gym_b200.add_agent import ADDAgent

    base = ADDAgent.__new__(ADDAgent)
    assert not base._record_hooks_overridden()
    assert base._record_data_pre_step(None, None, None, None) is None and base._record_data_post_step(None, None, None, None) is None
    with pytest.raises(_lib.AddkError, match="fused into addk_update_minibatch"):
        base._compute_loss({})

    class Recording(ADDAgent):
        def _record_data_post_step(self, next_obs, r, done, next_info):
            self.seen.append(("post", next_obs, r, done, sorted(next_info)))

        def _record_data_pre_step(self, obs, info, action, action_info):
            self.seen.append(("pre", obs, action, sorted(action_info), info))

    sub = Recording.__new__(Recording)
    assert sub._record_hooks_overridden()
    buf = _StubBuffer(4, 3)
    info = {"disc_obs": torch.zeros(3, 114), "disc_obs_demo": torch.zeros(3, 114)}
    for k, v in (("seen", []), ("_exp_buffer", buf), ("_add_obs", type("O", (), {"info": info})())):
        object.__setattr__(sub, k, v)
    sub._call_record_hooks(2)
    (pre, obs, action, a_keys, got_info), (post, next_obs, r, done, n_keys) = sub.seen
    assert (pre, post) == ("pre", "post") and got_info is info
    assert torch.equal(obs, buf.data["obs"][2]) and torch.equal(action, buf.data["action"][2]) and a_keys == ["a_logp", "rand_action_mask"]
    assert torch.equal(next_obs, buf.data["next_obs"][2]) and torch.equal(r, buf.data["reward"][2]) and torch.equal(done, buf.data["done"][2])
    assert n_keys == ["disc_obs", "disc_obs_demo"]

    class OwnLoss(ADDAgent):
        def _compute_loss(self, batch):
            return {"loss": 0.0}

    with pytest.raises(_lib.AddkError, match="overrides _compute_loss"):
        OwnLoss(b200_config.default_config(num_envs=4))
